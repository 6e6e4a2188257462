"""
ctypes wrapper of the C/OpenMP oracle (oracle/sv_oracle.c).  Test / baseline infrastructure
only -- see the header of sv_oracle.c.  ``build()`` compiles it with gcc; nothing here is
imported by the product package.
"""

import ctypes as ct
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liborc.so")
_lib = None
_ENT = {"cx": 0, "cz": 1, "cp": 2}


def build() -> str:
    res = subprocess.run(["make", "-C", _HERE, "liborc.so"], capture_output=True, text=True, check=False)
    if res.returncode != 0:
        raise RuntimeError("building oracle/liborc.so failed:\n" + res.stdout + res.stderr)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        if not os.path.isfile(_LIB):
            build()
        lib = ct.CDLL(_LIB)
        i32p, dp = ct.POINTER(ct.c_int32), ct.c_void_p
        lib.orc_apply.argtypes = [ct.c_int] * 5 + [i32p, i32p, dp, dp, ct.c_int]
        lib.orc_grad.argtypes = [ct.c_int] * 5 + [i32p, i32p, dp, dp, dp, dp]
        lib.orc_num_threads.restype = ct.c_int
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(_load().orc_num_threads())


def use_all_cores() -> int:
    """All host cores for the OpenMP loops, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    lib = _load()
    lib.orc_set_num_threads(int(os.cpu_count() or 1))
    return num_threads()


def _desc(circ, as_generic):
    trot = 0
    if hasattr(circ, "is_second_order") and not as_generic:
        trot = 2 if circ.is_second_order else 1
    blocks = np.ascontiguousarray(circ.blocks, dtype=np.int32)
    p = ct.POINTER(ct.c_int32)
    return trot, blocks, blocks[0].ctypes.data_as(p), blocks[1].ctypes.data_as(p)


def apply_v(circ, thetas, state, dagger=False, log2_cols=0):
    """V @ state or V^H @ state (flat complex128 array of 2^(n + log2_cols) entries)."""
    lib = _load()
    trot, blocks, pc, pt = _desc(circ, log2_cols > 0)
    th = np.ascontiguousarray(thetas, dtype=np.float64)
    out = np.array(state, dtype=np.complex128).ravel().copy()
    lib.orc_apply(circ.num_qubits, log2_cols, _ENT[circ.entangler], trot, circ.num_blocks, pc, pt,
                  th.ctypes.data, out.ctypes.data, int(dagger))
    return out.reshape(np.shape(state))


def grad_sweep(circ, thetas, x, z0, log2_cols=0, inplace=False):
    """Complex gradient of <V x|y> given z0 = V^H y; returns (grad, w_final, z_final)."""
    lib = _load()
    trot, blocks, pc, pt = _desc(circ, log2_cols > 0)
    th = np.ascontiguousarray(thetas, dtype=np.float64)
    w = x if inplace else np.array(x, dtype=np.complex128).ravel().copy()
    z = z0 if inplace else np.array(z0, dtype=np.complex128).ravel().copy()
    grad = np.zeros(th.size, dtype=np.complex128)
    lib.orc_grad(circ.num_qubits, log2_cols, _ENT[circ.entangler], trot, circ.num_blocks, pc, pt,
                 th.ctypes.data, w.ctypes.data, z.ctypes.data, grad.ctypes.data)
    return grad, w, z
