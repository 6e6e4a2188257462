"""
Thin Python handles over the C-ABI (``include/aqc_b200.h``): a circuit-structure handle and
a GPU workspace of state "slots".  Everything numeric happens inside ``libaqc_b200.so``.
"""

import ctypes as ct
from typing import Optional, Sequence
import numpy as np
from . import _lib
from .parametric_circuit import ParametricCircuit, is_parametric_circuit, is_trotter_ansatz

_ENT_CODE = {"cx": 0, "cz": 1, "cp": 2}


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(ct.c_void_p)


def _thetas_ptr(thetas: np.ndarray, expected: int):
    th = np.ascontiguousarray(thetas, dtype=np.float64)
    if th.size != expected:
        raise ValueError(f"expects {expected} angular parameters, got {th.size}")
    return th, th.ctypes.data_as(_lib.c_double_p)


def circuit_signature(circ, as_generic: bool = False):
    """Hashable description of a circuit structure, computed on the host without a C handle."""
    trotter = 0
    if is_trotter_ansatz(circ) and not as_generic:
        trotter = 2 if circ.is_second_order else 1
    blocks = np.ascontiguousarray(circ.blocks, dtype=np.int32)
    return (int(circ.num_qubits), circ.entangler, trotter, blocks.tobytes())


class CircuitHandle:
    """Owns an ``aqc_circuit`` built from a ParametricCircuit / TrotterAnsatz description."""

    def __init__(self, circ: ParametricCircuit, as_generic: bool = False):
        assert is_parametric_circuit(circ)
        self._lib = _lib.load()
        self.num_qubits = circ.num_qubits
        self.num_thetas = circ.num_thetas
        self.entangler = circ.entangler
        trotter = 0
        if is_trotter_ansatz(circ) and not as_generic:
            trotter = 2 if circ.is_second_order else 1
        self.trotter = trotter
        blocks = np.ascontiguousarray(circ.blocks, dtype=np.int32)
        self.blocks = blocks.copy()
        handle = ct.c_void_p()
        _lib.check(
            self._lib.aqc_circuit_create(
                circ.num_qubits,
                _ENT_CODE[circ.entangler],
                blocks.ctypes.data_as(_lib.c_int32_p),
                int(blocks.shape[1]),
                trotter,
                ct.byref(handle),
            )
        )
        self.handle = handle

    def signature(self):
        """Hashable description of the structure (used to detect structural changes)."""
        return (self.num_qubits, self.entangler, self.trotter, self.blocks.tobytes())

    def debug_program(self, log2_cols: int, tile_bits: int, low_bits: int, reversed_: bool,
                      dense: bool = False, fused: bool = False):
        """Serialised tile-pass program (host-only; see aqc_debug_program / aqc_debug_dense_program)."""
        fn = self._lib.aqc_debug_dense_program if dense else self._lib.aqc_debug_program
        if dense and fused:
            fn = self._lib.aqc_debug_dense_program_fused
        need = ct.c_int64(0)
        _lib.check(
            fn(
                self.handle, log2_cols, tile_bits, low_bits, int(reversed_), None, 0, ct.byref(need)
            )
        )
        buf = np.zeros(need.value, dtype=np.int32)
        _lib.check(
            fn(
                self.handle,
                log2_cols,
                tile_bits,
                low_bits,
                int(reversed_),
                buf.ctypes.data_as(_lib.c_int32_p),
                buf.size,
                ct.byref(need),
            )
        )
        return buf

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self._lib.aqc_circuit_destroy(h)


class SvWorkspace:
    """
    GPU workspace: ``num_slots`` arrays, each ``batch`` states of 2^(n + log2_cols) complex128.
    Raises if no CUDA device is visible (there is no CPU path).
    """

    def __init__(
        self,
        circ: ParametricCircuit,
        num_slots: int,
        *,
        device: int = 0,
        log2_cols: int = 0,
        batch: int = 1,
        as_generic: bool = False,
    ):
        self._lib = _lib.load()
        self.circuit = CircuitHandle(circ, as_generic=as_generic or log2_cols > 0)
        self.num_thetas = self.circuit.num_thetas
        self.batch = int(batch)
        self.log2_cols = int(log2_cols)
        self.num_slots = int(num_slots)
        handle = ct.c_void_p()
        _lib.check(
            self._lib.aqc_sv_create(
                self.circuit.handle, device, log2_cols, batch, num_slots, ct.byref(handle)
            )
        )
        self.handle = handle
        self.size = int(self._lib.aqc_sv_state_size(handle))
        self.can_eval = bool(self._lib.aqc_sv_can_eval(handle))  # one-submission evaluations (eval_begin)

    # -- data movement ---------------------------------------------------------------------
    def upload(self, slot: int, data: np.ndarray, batch_index: int = -1):
        arr = np.ascontiguousarray(data, dtype=np.complex128).ravel()
        if arr.size != self.size:
            raise ValueError(f"expects {self.size} amplitudes, got {arr.size}")
        _lib.check(self._lib.aqc_sv_upload(self.handle, slot, batch_index, _dptr(arr), arr.size))

    def download(self, slot: int, batch_index: int = 0, out: Optional[np.ndarray] = None):
        if out is None:
            out = np.empty(self.size, dtype=np.complex128)
        assert out.dtype == np.complex128 and out.flags.c_contiguous and out.size == self.size
        _lib.check(self._lib.aqc_sv_download(self.handle, slot, batch_index, _dptr(out), out.size))
        return out

    def set_basis(self, slot: int, index: int):
        _lib.check(self._lib.aqc_sv_set_basis(self.handle, slot, int(index)))

    def set_sparse(self, slot: int, indices: Sequence[int], amplitudes: Sequence[complex]):
        """slot = sum_k amplitudes[k] |indices[k]> (at most 8 terms); stream ordered, does not wait."""
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        amp = np.ascontiguousarray(amplitudes, dtype=np.complex128)
        if idx.ndim != 1 or idx.shape != amp.shape:
            raise ValueError("indices and amplitudes must be 1D sequences of equal length")
        _lib.check(
            self._lib.aqc_sv_set_sparse(
                self.handle, slot, idx.ctypes.data_as(_lib.c_int64_p), _dptr(amp), idx.size
            )
        )

    def set_identity(self, slot: int):
        _lib.check(self._lib.aqc_sv_set_identity(self.handle, slot))

    def fill_random(self, slot: int, seed: int):
        _lib.check(self._lib.aqc_sv_fill_random(self.handle, slot, int(seed)))

    def gather(self, slot: int, indices: Sequence[int]) -> np.ndarray:
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        out = np.empty((self.batch, idx.size), dtype=np.complex128)
        _lib.check(
            self._lib.aqc_sv_gather(
                self.handle, slot, idx.ctypes.data_as(_lib.c_int64_p), idx.size, _dptr(out)
            )
        )
        return out

    def vdot(self, slot_a: int, slot_b: int) -> np.ndarray:
        out = np.empty(self.batch, dtype=np.complex128)
        _lib.check(self._lib.aqc_sv_vdot(self.handle, slot_a, slot_b, _dptr(out)))
        return out

    # -- compute -----------------------------------------------------------------------------
    def apply(self, thetas: np.ndarray, src: int, dst: int, dagger: bool = False):
        _, ptr = _thetas_ptr(thetas, self.batch * self.num_thetas)
        _lib.check(self._lib.aqc_sv_apply(self.handle, ptr, int(dagger), src, dst))

    def objective(self, thetas: np.ndarray, target: int, z0: int, indices) -> np.ndarray:
        """z0 = V^H target; returns z0[b][indices] as (batch, len(indices)) complex array."""
        _, ptr = _thetas_ptr(thetas, self.batch * self.num_thetas)
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        out = np.empty((self.batch, idx.size), dtype=np.complex128)
        _lib.check(
            self._lib.aqc_sv_objective(
                self.handle, ptr, target, z0, idx.ctypes.data_as(_lib.c_int64_p), idx.size, _dptr(out)
            )
        )
        return out

    def grad(
        self,
        thetas: np.ndarray,
        *,
        z0: int,
        w: int,
        z: int,
        x_slot: int = -1,
        x_basis: int = 0,
    ) -> np.ndarray:
        """Raw complex gradient (batch, num_thetas) of <V x|y> given slot z0 = V^H y."""
        _, ptr = _thetas_ptr(thetas, self.batch * self.num_thetas)
        out = np.empty((self.batch, self.num_thetas), dtype=np.complex128)
        _lib.check(
            self._lib.aqc_sv_grad(self.handle, ptr, x_slot, int(x_basis), z0, w, z, _dptr(out))
        )
        return out

    def grad_begin(self, thetas: np.ndarray, *, z0: int, w: int, z: int, x_slot: int = -1, x_basis: int = 0):
        """Enqueues the gradient sweep and returns at once; collect the result with ``grad_end``."""
        _, ptr = _thetas_ptr(thetas, self.batch * self.num_thetas)
        _lib.check(self._lib.aqc_sv_grad_begin(self.handle, ptr, x_slot, int(x_basis), z0, w, z))

    def grad_end(self) -> np.ndarray:
        out = np.empty((self.batch, self.num_thetas), dtype=np.complex128)
        _lib.check(self._lib.aqc_sv_grad_end(self.handle, _dptr(out)))
        return out

    def eval_begin(self, thetas: np.ndarray, target: int, z0: int, indices, *, x_basis: int, w: int, z: int) -> np.ndarray:
        """
        Enqueues one whole evaluation -- V^H sweep, gather, gradient sweep from ``|x_basis>`` -- and
        returns hs = z0[indices] (batch, len(indices)) as soon as the gather has finished; the gradient
        sweep keeps running and is collected with ``grad_end`` (or dropped by the next call).
        """
        _, ptr = _thetas_ptr(thetas, self.batch * self.num_thetas)
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        _lib.check(
            self._lib.aqc_sv_eval_begin(
                self.handle, ptr, target, z0, idx.ctypes.data_as(_lib.c_int64_p), idx.size, int(x_basis), w, z
            )
        )
        out = np.empty((self.batch, idx.size), dtype=np.complex128)
        _lib.check(self._lib.aqc_sv_eval_hs(self.handle, _dptr(out)))
        return out

    def eval_times(self):
        """(V^H sweep ms, gradient sweep ms) of the last completed eval_begin .. grad_end pair."""
        a, b = ct.c_float(0), ct.c_float(0)
        _lib.check(self._lib.aqc_sv_eval_times(self.handle, ct.byref(a), ct.byref(b)))
        return float(a.value), float(b.value)

    def coord_descent(self, thetas: np.ndarray, *, target: int, w: int, z: int, num_sweeps: int = 1):
        """
        ``num_sweeps`` coordinate-descent sweeps (coord_descent_single_sweep,
        core_op_matrix.py:765-917) for every start of the batch.  Returns
        (fobj[num_sweeps, batch], new thetas[batch, num_thetas]); ``thetas`` is not modified.
        """
        th = np.array(thetas, dtype=np.float64).reshape(-1).copy()
        if th.size != self.batch * self.num_thetas:
            raise ValueError(f"expects {self.batch * self.num_thetas} angular parameters, got {th.size}")
        fobj = np.empty((int(num_sweeps), self.batch), dtype=np.float64)
        _lib.check(
            self._lib.aqc_sv_coord_descent(
                self.handle, th.ctypes.data_as(_lib.c_double_p), target, w, z, int(num_sweeps),
                fobj.ctypes.data_as(_lib.c_double_p),
            )
        )
        return fobj, th.reshape(self.batch, self.num_thetas)

    # -- sketching generators (dense target, DMMA GEMM, thin QR) --------------------------------
    def set_dense_target(self, target: np.ndarray):
        dim = 1 << self.circuit.num_qubits
        arr = np.ascontiguousarray(target, dtype=np.complex128)
        if arr.shape != (dim, dim):
            raise ValueError(f"expects a {dim} x {dim} target matrix")
        _lib.check(self._lib.aqc_sv_set_dense_target(self.handle, _dptr(arr)))

    def target_matmul(self, src: int, dst: int, conj_transpose: bool = False):
        """dst = U @ src or U^H @ src (U: the dense target)."""
        _lib.check(self._lib.aqc_sv_target_matmul(self.handle, int(conj_transpose), src, dst))

    def orthonormalize(self, slot: int, tmp: int):
        """slot <- orthonormal basis of its column space (thin QR up to a unitary factor)."""
        _lib.check(self._lib.aqc_sv_orthonormalize(self.handle, slot, tmp))

    def sub(self, dst: int, src: int):
        _lib.check(self._lib.aqc_sv_sub(self.handle, dst, src))

    def gather_target_columns(self, indices, x: int, y: int):
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        _lib.check(
            self._lib.aqc_sv_gather_target_columns(
                self.handle, idx.ctypes.data_as(_lib.c_int64_p), idx.size, x, y
            )
        )

    # -- introspection -----------------------------------------------------------------------
    @property
    def last_kernel_ms(self) -> float:
        return float(self._lib.aqc_sv_last_kernel_ms(self.handle))

    @property
    def last_num_launches(self) -> int:
        return int(self._lib.aqc_sv_last_num_launches(self.handle))

    def timer_start(self):
        _lib.check(self._lib.aqc_sv_timer_start(self.handle))

    def timer_stop(self) -> float:
        """Device milliseconds since ``timer_start`` (CUDA events on the workspace stream)."""
        ms = ct.c_float(0)
        _lib.check(self._lib.aqc_sv_timer_stop(self.handle, ct.byref(ms)))
        return float(ms.value)

    def num_passes(self, mode: int) -> int:
        return int(self._lib.aqc_sv_num_passes(self.handle, mode))

    def num_stages(self, mode: int) -> int:
        return int(self._lib.aqc_sv_num_stages(self.handle, mode))

    def slot_ptr(self, slot: int) -> int:
        return int(self._lib.aqc_sv_slot_ptr(self.handle, slot) or 0)

    def close(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self._lib.aqc_sv_destroy(h)

    def __del__(self):
        self.close()
