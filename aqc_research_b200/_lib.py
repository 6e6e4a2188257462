"""
ctypes binding of ``libaqc_b200.so`` (C-ABI declared in ``include/aqc_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` /
``make -C aqc_research_b200/csrc``.  Loading never falls back to anything else: if the
shared object is missing, ``load()`` raises.
"""

import ctypes as ct
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaqc_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

_lock = threading.Lock()
_lib = None

c_double_p = ct.POINTER(ct.c_double)
c_int32_p = ct.POINTER(ct.c_int32)
c_int64_p = ct.POINTER(ct.c_int64)

# name -> (restype, argtypes); mirrors include/aqc_b200.h one to one
SIGNATURES = {
    "aqc_last_error": (ct.c_char_p, []),
    "aqc_version": (ct.c_int, []),
    "aqc_device_count": (ct.c_int, []),
    "aqc_circuit_create": (
        ct.c_int,
        [ct.c_int, ct.c_int, c_int32_p, ct.c_int, ct.c_int, ct.POINTER(ct.c_void_p)],
    ),
    "aqc_circuit_destroy": (None, [ct.c_void_p]),
    "aqc_circuit_num_thetas": (ct.c_int, [ct.c_void_p]),
    "aqc_sv_create": (
        ct.c_int,
        [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.POINTER(ct.c_void_p)],
    ),
    "aqc_sv_destroy": (None, [ct.c_void_p]),
    "aqc_sv_state_size": (ct.c_int64, [ct.c_void_p]),
    "aqc_sv_upload": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, ct.c_void_p, ct.c_int64]),
    "aqc_sv_download": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, ct.c_void_p, ct.c_int64]),
    "aqc_sv_set_basis": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int64]),
    "aqc_sv_set_sparse": (ct.c_int, [ct.c_void_p, ct.c_int, c_int64_p, ct.c_void_p, ct.c_int]),
    "aqc_sv_set_identity": (ct.c_int, [ct.c_void_p, ct.c_int]),
    "aqc_sv_fill_random": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_uint64]),
    "aqc_sv_gather": (ct.c_int, [ct.c_void_p, ct.c_int, c_int64_p, ct.c_int, ct.c_void_p]),
    "aqc_sv_vdot": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, ct.c_void_p]),
    "aqc_sv_apply": (ct.c_int, [ct.c_void_p, c_double_p, ct.c_int, ct.c_int, ct.c_int]),
    "aqc_sv_grad": (
        ct.c_int,
        [ct.c_void_p, c_double_p, ct.c_int, ct.c_int64, ct.c_int, ct.c_int, ct.c_int, ct.c_void_p],
    ),
    "aqc_sv_grad_begin": (
        ct.c_int,
        [ct.c_void_p, c_double_p, ct.c_int, ct.c_int64, ct.c_int, ct.c_int, ct.c_int],
    ),
    "aqc_sv_grad_end": (ct.c_int, [ct.c_void_p, ct.c_void_p]),
    "aqc_sv_objective": (
        ct.c_int,
        [ct.c_void_p, c_double_p, ct.c_int, ct.c_int, c_int64_p, ct.c_int, ct.c_void_p],
    ),
    "aqc_sv_eval_begin": (
        ct.c_int,
        [ct.c_void_p, c_double_p, ct.c_int, ct.c_int, c_int64_p, ct.c_int, ct.c_int64, ct.c_int, ct.c_int],
    ),
    "aqc_sv_eval_hs": (ct.c_int, [ct.c_void_p, ct.c_void_p]),
    "aqc_sv_can_eval": (ct.c_int, [ct.c_void_p]),
    "aqc_sv_eval_times": (ct.c_int, [ct.c_void_p, ct.POINTER(ct.c_float), ct.POINTER(ct.c_float)]),
    "aqc_sv_last_kernel_ms": (ct.c_float, [ct.c_void_p]),
    "aqc_sv_last_num_launches": (ct.c_int, [ct.c_void_p]),
    "aqc_sv_timer_start": (ct.c_int, [ct.c_void_p]),
    "aqc_sv_timer_stop": (ct.c_int, [ct.c_void_p, ct.POINTER(ct.c_float)]),
    "aqc_sv_coord_descent": (
        ct.c_int,
        [ct.c_void_p, ct.POINTER(ct.c_double), ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.POINTER(ct.c_double)],
    ),
    "aqc_sv_set_dense_target": (ct.c_int, [ct.c_void_p, ct.c_void_p]),
    "aqc_sv_target_matmul": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int]),
    "aqc_sv_orthonormalize": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int]),
    "aqc_sv_sub": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int]),
    "aqc_sv_gather_target_columns": (ct.c_int, [ct.c_void_p, c_int64_p, ct.c_int, ct.c_int, ct.c_int]),
    "aqc_sv_num_passes": (ct.c_int, [ct.c_void_p, ct.c_int]),
    "aqc_sv_num_stages": (ct.c_int, [ct.c_void_p, ct.c_int]),
    "aqc_debug_program": (
        ct.c_int,
        [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, c_int32_p, ct.c_int64, c_int64_p],
    ),
    "aqc_debug_dense_program": (
        ct.c_int,
        [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, c_int32_p, ct.c_int64, c_int64_p],
    ),
    "aqc_debug_dense_program_fused": (
        ct.c_int,
        [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, c_int32_p, ct.c_int64, c_int64_p],
    ),
    "aqc_sv_slot_ptr": (ct.c_void_p, [ct.c_void_p, ct.c_int]),
    "aqc_sv_create_sharded": (
        ct.c_int,
        [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.POINTER(ct.c_void_p)],
    ),
    "aqc_sv_num_epochs": (ct.c_int, [ct.c_void_p, ct.c_int]),
    "aqc_sv_epoch_layout": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int]),
    "aqc_sv_begin": (ct.c_int, [ct.c_void_p, c_double_p, ct.c_int]),
    "aqc_sv_run_epoch": (
        ct.c_int,
        [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int64, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int],
    ),
    "aqc_sv_can_push": (ct.c_int, [ct.c_void_p]),
    "aqc_sv_ipc_close": (ct.c_int, [ct.c_void_p]),
    "aqc_sv_grad_finish": (ct.c_int, [ct.c_void_p, ct.c_void_p]),
    "aqc_sv_ipc_export": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_void_p]),
    "aqc_sv_ipc_import": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, ct.c_void_p]),
    "aqc_sv_peer_attach": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, ct.c_void_p]),
    "aqc_sv_exchange": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int]),
    "aqc_sv_fill_random_logical": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_uint64, c_double_p]),
    "aqc_sv_scale": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_double]),
    "aqc_debug_program_sharded": (
        ct.c_int,
        [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, c_int32_p, ct.c_int64, c_int64_p],
    ),
    "aqc_mps_create": (
        ct.c_int,
        [ct.c_void_p, ct.c_int, ct.c_int, ct.c_double, ct.c_int, ct.POINTER(ct.c_void_p)],
    ),
    "aqc_mps_destroy": (None, [ct.c_void_p]),
    "aqc_mps_bond_capacity": (ct.c_int, [ct.c_void_p]),
    "aqc_mps_upload": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_void_p, ct.c_void_p, c_int32_p]),
    "aqc_mps_download": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_void_p, ct.c_void_p, c_int32_p]),
    "aqc_prim_apply": (ct.c_int, [ct.c_int, ct.c_void_p, ct.c_int64, ct.c_int, c_int64_p, c_int32_p, ct.c_void_p]),
    "aqc_prim_dot": (ct.c_int, [ct.c_int, ct.c_void_p, ct.c_void_p, ct.c_int64, c_int64_p, c_int32_p, ct.c_void_p,
                                ct.c_void_p]),
    "aqc_mps_set_product": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int64]),
    "aqc_mps_set_product_site": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int64, ct.c_int, ct.c_void_p]),
    "aqc_mps_apply": (ct.c_int, [ct.c_void_p, c_double_p, ct.c_int, ct.c_int, ct.c_int]),
    "aqc_mps_amplitudes": (ct.c_int, [ct.c_void_p, ct.c_int, c_int64_p, ct.c_int, ct.c_void_p]),
    "aqc_mps_objective": (
        ct.c_int,
        [ct.c_void_p, c_double_p, ct.c_int, ct.c_int, c_int64_p, ct.c_int, ct.c_void_p],
    ),
    "aqc_mps_dot": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, ct.c_void_p]),
    "aqc_mps_grad": (
        ct.c_int,
        [ct.c_void_p, c_double_p, ct.c_int, ct.c_int64, ct.c_int, ct.c_int, ct.c_int, ct.c_void_p],
    ),
    "aqc_mps_debug_sweeps": (ct.c_int, [ct.c_void_p, c_int32_p, ct.c_int]),
    "aqc_mps_last_kernel_ms": (ct.c_float, [ct.c_void_p]),
    "aqc_mps_last_num_launches": (ct.c_int, [ct.c_void_p]),
    "aqc_mps_truncation_stats": (ct.c_int, [ct.c_void_p, c_double_p]),
    "aqc_sv_stream": (ct.c_void_p, [ct.c_void_p]),
}


class AqcError(RuntimeError):
    """Error reported by the CUDA library."""


def build(verbose: bool = False) -> str:
    """Compiles the shared library in-tree with nvcc for sm_100a. Returns its path."""
    res = subprocess.run(["make", "-j4", "-C", CSRC_DIR], capture_output=True, text=True, check=False)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0 or not os.path.isfile(LIB_PATH):
        raise AqcError("building libaqc_b200.so failed (see output above)")
    return LIB_PATH


def load() -> ct.CDLL:
    """Loads the library (once) and declares every signature. Raises if it is missing."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise AqcError(
                f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C aqc_research_b200/csrc`. There is no CPU fallback."
            )
        lib = ct.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int) -> None:
    """Raises AqcError with the library's message if ``rc`` is an error code."""
    if rc != 0:
        msg = load().aqc_last_error()
        text = msg.decode("utf-8", "replace") if msg else "unknown error"
        if rc == -1:
            raise ValueError(text)
        raise AqcError(f"[{rc}] {text}")


def device_count() -> int:
    return int(load().aqc_device_count())


def require_gpu() -> None:
    """Raises unless at least one CUDA device is visible."""
    if device_count() <= 0:
        raise AqcError("no CUDA device visible: aqc_research_b200 has no CPU fallback")
