"""
Function-level drop-ins of the reference's MPS helpers (aqc_research/mps_operations.py).
``QiskitMPS`` tuples in / out; contraction, gate application and truncation run on the GPU
(``MpsWorkspace``).  ``check_mps``, ``_preprocess_mps`` and ``mps_to_vector`` are host-side
format utilities (pure NumPy, as in the reference :87-189).
"""

from typing import List, Optional
import numpy as np
from .mps_engine import MpsWorkspace, QiskitMPS
from .parametric_circuit import ParametricCircuit

_NO_TRUNCATION_THR = 1e-16
_CHI_MAX = 64


def no_truncation_threshold() -> float:
    return _NO_TRUNCATION_THR


def check_mps(qiskit_mps) -> bool:
    """True if the argument has the structure of a Qiskit MPS (mps_operations.py:87-123)."""
    if not (isinstance(qiskit_mps, tuple) and len(qiskit_mps) == 2):
        return False
    gam, lam = qiskit_mps
    n = len(gam)
    if len(lam) != n - 1:
        return False
    for k in range(n):
        if len(gam[k]) != 2 or gam[k][0].ndim != 2 or gam[k][0].shape != gam[k][1].shape:
            return False
        if k < n - 1:
            l = np.asarray(lam[k])
            if not (l.ndim == 1 or (l.ndim == 2 and min(l.shape) == 1)):
                return False
            l = l.ravel()
            if not np.all(l[:-1] >= l[1:]):
                return False
    return True


def _preprocess_mps(qiskit_mps: QiskitMPS, conjugate: bool = False) -> List[np.ndarray]:
    """Per-site tensors A_k[b] = Gamma_k[b] diag(lambda_k) of shape (2, chi_k, chi_{k+1})."""
    assert check_mps(qiskit_mps)
    gam, lam = qiskit_mps
    out = []
    for k, (g0, g1) in enumerate(gam):
        a = np.stack((g0, g1)).astype(np.complex128)
        if k < len(gam) - 1:
            a = a * np.asarray(lam[k]).ravel()[None, None, :]
        out.append(np.conj(a) if conjugate else a)
    return out


def mps_to_vector(qiskit_mps: QiskitMPS) -> np.ndarray:
    """Dense state of 2^n amplitudes (testing aid; exponential cost)."""
    mats = _preprocess_mps(qiskit_mps)
    psi = mats[0].reshape(2, -1)
    for a in mats[1:]:
        psi = np.einsum("xa,bac->bxc", psi, a).reshape(-1, a.shape[2])
    return psi.reshape(-1)


def _workspace(circ: ParametricCircuit, trunc_thr: float, slots: int = 4, chi_max: int = _CHI_MAX) -> MpsWorkspace:
    return MpsWorkspace(circ, num_slots=slots, chi_max=chi_max, trunc_thr=trunc_thr)


def mps_dot(qiskit_mps1: QiskitMPS, qiskit_mps2: QiskitMPS) -> np.complex128:
    """<mps1|mps2> on the GPU (mps_operations.py:192-213)."""
    from .circuit_structures import make_trotter_like_circuit  # pylint: disable=import-outside-toplevel
    from .parametric_circuit import TrotterAnsatz  # pylint: disable=import-outside-toplevel

    n = len(qiskit_mps1[0])
    assert len(qiskit_mps2[0]) == n
    circ = TrotterAnsatz(n, make_trotter_like_circuit(n, 0), False)  # structure-less carrier
    ws = _workspace(circ, _NO_TRUNCATION_THR, slots=2)
    ws.upload(0, qiskit_mps1)
    ws.upload(1, qiskit_mps2)
    val = ws.dot(0, 1)
    ws.close()
    return np.complex128(val)


def v_mul_mps(circ: ParametricCircuit, thetas: np.ndarray, mps_vec: QiskitMPS, *,
              trunc_thr: Optional[float] = _NO_TRUNCATION_THR, chi_max: int = _CHI_MAX) -> QiskitMPS:
    """
    ``V @ mps_vec`` (mps_operations.py:326-346).  Deviation from the reference: the GPU engine caps the
    bond dimension at ``chi_max`` (<= 64); if the cap -- rather than ``trunc_thr`` -- removes more weight
    than ``trunc_thr``, ``BondCapacityError`` is raised instead of returning a silently truncated state.
    """
    ws = _workspace(circ, trunc_thr, slots=2, chi_max=chi_max)
    try:
        ws.upload(0, mps_vec)
        ws.apply(thetas, 0, 1, dagger=False)
        ws.check_cap("v_mul_mps")
        return ws.download(1)
    finally:
        ws.close()


def v_dagger_mul_mps(circ: ParametricCircuit, thetas: np.ndarray, mps_vec: QiskitMPS, *,
                     trunc_thr: Optional[float] = _NO_TRUNCATION_THR, chi_max: int = _CHI_MAX) -> QiskitMPS:
    """``V^H @ mps_vec`` (mps_operations.py:349-371); the bond cap is reported as in ``v_mul_mps``."""
    ws = _workspace(circ, trunc_thr, slots=2, chi_max=chi_max)
    try:
        ws.upload(0, mps_vec)
        ws.apply(thetas, 0, 1, dagger=True)
        ws.check_cap("v_dagger_mul_mps")
        return ws.download(1)
    finally:
        ws.close()


def rand_mps_vec(num_qubits: int, out_state: Optional[np.ndarray] = None, num_layers: int = 3) -> QiskitMPS:
    """
    Random state in MPS format (mps_operations.py:301-323): a random-angle "spin" ansatz of
    ``num_layers * (num_qubits - 1)`` unit blocks with a randomly chosen entangler applied to
    |0...0>, evaluated by the GPU MPS engine without truncation (the reference builds the Qiskit
    circuit and runs qiskit-aer).  ``out_state``, if given, receives the dense state vector.
    """
    from . import circuit_structures as cs  # pylint: disable=import-outside-toplevel
    from . import utils  # pylint: disable=import-outside-toplevel

    assert isinstance(num_qubits, (int, np.integer)) and num_qubits >= 2
    assert isinstance(num_layers, (int, np.integer)) and num_layers > 0
    blocks = cs.create_ansatz_structure(num_qubits, "spin", "full", num_layers * (num_qubits - 1))
    circ = ParametricCircuit(num_qubits, str(np.random.choice(["cx", "cz", "cp"])), blocks)
    thetas = utils.rand_thetas(circ.num_thetas)
    ws = _workspace(circ, _NO_TRUNCATION_THR, slots=1)
    try:
        ws.set_product(0, 0)
        ws.apply(thetas, 0, 0, dagger=False)
        ws.check_cap("rand_mps_vec")
        mps = ws.download(0)
    finally:
        ws.close()
    if out_state is not None:
        assert out_state.shape == (2**num_qubits,)
        np.copyto(out_state, mps_to_vector(mps))
    return mps
