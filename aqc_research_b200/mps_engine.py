"""
Python handle of the GPU MPS workspace (C-ABI ``aqc_mps_*`` of include/aqc_b200.h) and the
conversion between the reference's ``QiskitMPS`` tuples and the padded device layout.
"""

import ctypes as ct
from typing import List, Sequence, Tuple
import numpy as np
from . import _lib
from .engine import CircuitHandle, _dptr, _thetas_ptr
from .parametric_circuit import ParametricCircuit

QiskitMPS = Tuple[List[Tuple[np.ndarray, np.ndarray]], List[np.ndarray]]


class BondCapacityError(RuntimeError):
    """The bond-dimension cap of the GPU MPS engine truncated a state the caller wanted exact."""


def pack_mps(mps: QiskitMPS, capacity: int):
    """QiskitMPS -> (gam[n][2][C][C], lam[n+1][C], dims[n+1]) padded arrays."""
    gammas, lambdas = mps
    n = len(gammas)
    gam = np.zeros((n, 2, capacity, capacity), dtype=np.complex128)
    lam = np.zeros((n + 1, capacity), dtype=np.float64)
    dims = np.ones(n + 1, dtype=np.int32)
    lam[0, 0] = lam[n, 0] = 1.0
    for k, (g0, g1) in enumerate(gammas):
        g0, g1 = np.asarray(g0), np.asarray(g1)
        rows, cols = g0.shape
        if rows > capacity or cols > capacity:
            raise ValueError(f"bond dimension {max(rows, cols)} exceeds the capacity {capacity}")
        gam[k, 0, :rows, :cols] = g0
        gam[k, 1, :rows, :cols] = g1
        dims[k], dims[k + 1] = rows, cols
    for k, l in enumerate(lambdas):
        l = np.asarray(l, dtype=np.float64).ravel()
        lam[k + 1, : l.size] = l
        if dims[k + 1] != l.size:
            raise ValueError("inconsistent bond dimensions in MPS")
    return gam, lam, dims


def unpack_mps(gam: np.ndarray, lam: np.ndarray, dims: np.ndarray) -> QiskitMPS:
    """Padded arrays -> QiskitMPS (gammas, lambdas)."""
    n = gam.shape[0]
    gammas = [
        (gam[k, 0, : dims[k], : dims[k + 1]].copy(), gam[k, 1, : dims[k], : dims[k + 1]].copy())
        for k in range(n)
    ]
    lambdas = [lam[k + 1, : dims[k + 1]].copy() for k in range(n - 1)]
    return gammas, lambdas


class MpsWorkspace:
    """``num_slots`` MPS states on one GPU plus the scratch of the MPS sweeps."""

    def __init__(self, circ: ParametricCircuit, num_slots: int, *, chi_max: int = 64,
                 trunc_thr: float = 1e-16, device: int = 0):
        self._lib = _lib.load()
        self.circuit = CircuitHandle(circ)
        self.num_qubits = circ.num_qubits
        self.num_thetas = circ.num_thetas
        handle = ct.c_void_p()
        _lib.check(
            self._lib.aqc_mps_create(
                self.circuit.handle, device, int(chi_max), float(trunc_thr), num_slots, ct.byref(handle)
            )
        )
        self.handle = handle
        self.capacity = int(self._lib.aqc_mps_bond_capacity(handle))
        self.chi_max = int(chi_max)
        self.trunc_thr = float(trunc_thr)

    def upload(self, slot: int, mps: QiskitMPS):
        if len(mps[0]) != self.num_qubits:
            raise ValueError("MPS has a wrong number of qubits")
        gam, lam, dims = pack_mps(mps, self.capacity)
        _lib.check(
            self._lib.aqc_mps_upload(self.handle, slot, _dptr(gam), _dptr(lam), dims.ctypes.data_as(_lib.c_int32_p))
        )

    def download(self, slot: int) -> QiskitMPS:
        n, c = self.num_qubits, self.capacity
        gam = np.empty((n, 2, c, c), dtype=np.complex128)
        lam = np.empty((n + 1, c), dtype=np.float64)
        dims = np.empty(n + 1, dtype=np.int32)
        _lib.check(
            self._lib.aqc_mps_download(self.handle, slot, _dptr(gam), _dptr(lam), dims.ctypes.data_as(_lib.c_int32_p))
        )
        return unpack_mps(gam, lam, dims)

    def set_product(self, slot: int, index: int):
        _lib.check(self._lib.aqc_mps_set_product(self.handle, slot, int(index)))

    def set_product_site(self, slot: int, index: int, site: int, amp0: complex, amp1: complex):
        """|index> with qubit ``site`` replaced by amp0 |0> + amp1 |1> (bond dimension 1)."""
        amps = np.array([amp0, amp1], dtype=np.complex128)
        _lib.check(self._lib.aqc_mps_set_product_site(self.handle, slot, int(index), int(site), _dptr(amps)))

    def apply(self, thetas: np.ndarray, src: int, dst: int, dagger: bool = False):
        _, ptr = _thetas_ptr(thetas, self.num_thetas)
        _lib.check(self._lib.aqc_mps_apply(self.handle, ptr, int(dagger), src, dst))

    def amplitudes(self, slot: int, indices: Sequence[int]) -> np.ndarray:
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        out = np.empty(idx.size, dtype=np.complex128)
        _lib.check(
            self._lib.aqc_mps_amplitudes(self.handle, slot, idx.ctypes.data_as(_lib.c_int64_p), idx.size, _dptr(out))
        )
        return out

    def objective(self, thetas: np.ndarray, target: int, z0: int, indices) -> np.ndarray:
        _, ptr = _thetas_ptr(thetas, self.num_thetas)
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        out = np.empty(idx.size, dtype=np.complex128)
        _lib.check(
            self._lib.aqc_mps_objective(
                self.handle, ptr, target, z0, idx.ctypes.data_as(_lib.c_int64_p), idx.size, _dptr(out)
            )
        )
        return out

    def dot(self, slot_a: int, slot_b: int) -> complex:
        out = np.empty(1, dtype=np.complex128)
        _lib.check(self._lib.aqc_mps_dot(self.handle, slot_a, slot_b, _dptr(out)))
        return complex(out[0])

    def grad(self, thetas: np.ndarray, *, z0: int, w: int, z: int, x_slot: int = -1, x_basis: int = 0):
        _, ptr = _thetas_ptr(thetas, self.num_thetas)
        out = np.empty(self.num_thetas, dtype=np.complex128)
        _lib.check(self._lib.aqc_mps_grad(self.handle, ptr, x_slot, int(x_basis), z0, w, z, _dptr(out)))
        return out

    def truncation_stats(self) -> dict:
        """
        What the SVD splits of the most recent apply / objective / grad call discarded:
        ``discarded_weight`` (sum over splits of the dropped squared Schmidt values, relative to each
        split), ``max_discarded``, ``cap_discarded`` (the share removed only because of the ``chi_max``
        cap -- the reference's qiskit-aer run has no cap) and ``cap_hits`` (splits the cap cut).
        """
        out = np.zeros(4, dtype=np.float64)
        _lib.check(self._lib.aqc_mps_truncation_stats(self.handle, out.ctypes.data_as(_lib.c_double_p)))
        return {"discarded_weight": float(out[0]), "max_discarded": float(out[1]),
                "cap_discarded": float(out[2]), "cap_hits": int(out[3])}

    def check_cap(self, what: str):
        """Raises if the chi_max cap (not the trunc_thr rule) removed more weight than trunc_thr allows."""
        st = self.truncation_stats()
        if st["cap_hits"] and st["cap_discarded"] > max(self.trunc_thr, 1e-14):
            raise BondCapacityError(
                f"{what}: the bond-dimension cap chi_max = {self.chi_max} removed a weight of "
                f"{st['cap_discarded']:.3e} in {st['cap_hits']} splits (trunc_thr = {self.trunc_thr:g}); "
                "the reference (qiskit-aer) has no cap -- raise chi_max (<= 64) or trunc_thr"
            )

    @property
    def last_kernel_ms(self) -> float:
        return float(self._lib.aqc_mps_last_kernel_ms(self.handle))

    @property
    def last_num_launches(self) -> int:
        return int(self._lib.aqc_mps_last_num_launches(self.handle))

    def close(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self._lib.aqc_mps_destroy(h)

    def __del__(self):
        self.close()
