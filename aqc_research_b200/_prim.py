"""
Host side of the single-gate primitives (csrc/aqc_prim.cu): a short list of 2x2 gates, each acting
on the index bit given by its stride, applied on the GPU to a host array in place; and
``np.vdot(G w, z)`` for one such gate.  Used by the gate-by-gate drop-ins of ``core_operations`` and
``core_op_matrix``; there is no CPU path.
"""

import ctypes as ct
from typing import Sequence, Tuple
import numpy as np
from . import _lib

# (target stride, control stride, mode, 2x2 gate); mode 0 plain, 1 controlled, 2 controlled with zero branch
Op = Tuple[int, int, int, np.ndarray]


def _pack(ops: Sequence[Op]):
    strides = np.array([[int(o[0]), int(o[1])] for o in ops], dtype=np.int64)
    modes = np.array([int(o[2]) for o in ops], dtype=np.int32)
    gates = np.ascontiguousarray([np.asarray(o[3], dtype=np.complex128).reshape(2, 2) for o in ops])
    return strides, modes, gates


def apply_gates(arr: np.ndarray, ops: Sequence[Op], device: int = 0) -> np.ndarray:
    """Applies ``ops`` in order to the C-contiguous complex128 array ``arr`` (any shape), in place."""
    assert isinstance(arr, np.ndarray) and arr.dtype == np.complex128 and arr.flags.c_contiguous
    strides, modes, gates = _pack(ops)
    lib = _lib.load()
    _lib.check(
        lib.aqc_prim_apply(
            device, arr.ctypes.data_as(ct.c_void_p), arr.size, len(ops),
            strides.ctypes.data_as(_lib.c_int64_p), modes.ctypes.data_as(_lib.c_int32_p),
            gates.ctypes.data_as(ct.c_void_p),
        )
    )
    return arr


def gate_vdot(w: np.ndarray, z: np.ndarray, op: Op, device: int = 0) -> complex:
    """``np.vdot(G w, z)`` for the gate ``op``; the inputs are not modified."""
    for a in (w, z):
        assert isinstance(a, np.ndarray) and a.dtype == np.complex128 and a.flags.c_contiguous
    assert w.size == z.size
    strides, modes, gates = _pack([op])
    out = np.zeros(1, dtype=np.complex128)
    lib = _lib.load()
    _lib.check(
        lib.aqc_prim_dot(
            device, w.ctypes.data_as(ct.c_void_p), z.ctypes.data_as(ct.c_void_p), w.size,
            strides.ctypes.data_as(_lib.c_int64_p), modes.ctypes.data_as(_lib.c_int32_p),
            gates.ctypes.data_as(ct.c_void_p), out.ctypes.data_as(ct.c_void_p),
        )
    )
    return complex(out[0])
