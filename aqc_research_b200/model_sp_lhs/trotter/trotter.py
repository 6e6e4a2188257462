"""
Trotter helpers the hot path needs on the host: the "perfect" initial angles that make a
TrotterAnsatz equal to the Trotter circuit, and target-state generation with the GPU engine.
Reference: aqc_research/model_sp_lhs/trotter/trotter.py (trotter_alphas :269-283,
neel_init_state :389-398, half_zero_circuit :401-410, fidelity :413-423, slice2q :432-475,
init_ansatz_to_trotter :478-537).  The Qiskit circuit builders of that module (trotter_circuit,
Trotter.as_qcircuit, ...) are callers outside the hot path and are not rebuilt; state-preparation
"circuits" are represented by the list of qubits that carry an X gate.
"""

from typing import List, Optional, Tuple, Union
import numpy as np
from ... import checking as chk
from ...parametric_circuit import ParametricCircuit, TrotterAnsatz, first_layer_included, is_trotter_ansatz


def trotter_alphas(dt: float, delta: float) -> np.ndarray:
    """The three non-trivial angles of one Trotter building block (time step dt)."""
    assert chk.is_float(dt, dt > 0) and chk.is_float(delta, delta > 0)
    return np.array([0.5 * np.pi - 0.5 * delta * dt, 0.5 * dt - 0.5 * np.pi, 0.5 * np.pi - 0.5 * dt])


def neel_init_state(num_qubits: int) -> List[int]:
    """X-gate positions preparing the Neel state |..0101> (bits 0, 2, 4, ... set)."""
    assert chk.is_int(num_qubits, num_qubits >= 2)
    return list(range(0, num_qubits, 2))


def half_zero_circuit(num_qubits: int) -> List[int]:
    """X-gate positions preparing |1..10..0> (upper half of the qubits set)."""
    assert chk.is_int(num_qubits, num_qubits >= 2)
    return list(range(num_qubits // 2, num_qubits))


def basis_index(x_positions) -> int:
    index = 0
    for q in x_positions:
        index |= 1 << int(q)
    return index


def fidelity(state1, state2) -> float:
    """|<state1|state2>|^2 for two vectors or two MPS tuples."""
    if isinstance(state1, np.ndarray) and isinstance(state2, np.ndarray):
        assert chk.complex_1d(state1) and chk.complex_1d(state2)
        return float(np.abs(np.vdot(state1, state2)) ** 2)
    from ...mps_operations import mps_dot  # pylint: disable=import-outside-toplevel

    return float(np.abs(mps_dot(state1, state2)) ** 2)


def state_difference(state1: np.ndarray, state2: np.ndarray) -> float:
    assert chk.complex_1d(state1) and chk.complex_1d(state2)
    return float(np.linalg.norm(state1 - state2))


def slice2q(circ: ParametricCircuit, vec: np.ndarray, *, layer_range: Optional[Tuple[int, int]] = None):
    """View (layers, n-1 triplets, 12 angles) of the block part of a theta-sized vector."""
    if not is_trotter_ansatz(circ):
        raise ValueError("expects Trotterized ansatz")
    assert isinstance(vec, np.ndarray) and vec.shape == (circ.num_thetas,)
    layers = circ.num_layers
    layer_range = (0, layers) if layer_range is None else layer_range
    assert chk.is_tuple(layer_range, len(layer_range) == 2)
    assert 0 <= layer_range[0] < layer_range[1] <= layers
    view = circ.subset2q(vec).reshape(layers, circ.num_qubits - 1, 12)[layer_range[0] : layer_range[1]]
    assert np.shares_memory(view, vec)
    return view, layer_range


def init_ansatz_to_trotter(circ, thetas: np.ndarray, *, evol_time: float, delta: float,
                           layer_range: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """
    Sets (in place) the angles of the layers in ``layer_range`` so that the ansatz reproduces the
    Trotter circuit over ``evol_time``.  Per triplet only three angles are non-zero: entry 5
    (Rz of the middle block's control), entry 0 (Ry of the first block's control) and entry 6
    (Ry of the middle block's target).  For a second-order ansatz the leading half-layer (whose
    angles the implied trailing half-layer re-uses) gets half the time step.
    """
    blocks, layer_range = slice2q(circ, thetas, layer_range=layer_range)
    dt = evol_time / float(layer_range[1] - layer_range[0])
    with_first = first_layer_included(circ, layer_range)
    if with_first:
        circ.subset1q(thetas)[:] = 0
    blocks[:] = 0
    a = trotter_alphas(dt=dt, delta=delta)
    blocks[:, :, 5], blocks[:, :, 0], blocks[:, :, 6] = a[0], a[1], a[2]
    if circ.is_second_order and with_first:
        h = trotter_alphas(dt=0.5 * dt, delta=delta)
        half = circ.half_layer_num_blocks // 3
        blocks[0, :half, 5], blocks[0, :half, 0], blocks[0, :half, 6] = h[0], h[1], h[2]
    return thetas


def trotter_state(num_qubits: int, *, evol_time: float, num_steps: int, delta: float,
                  second_order: bool, ini_state: Union[int, List[int], np.ndarray] = 0) -> np.ndarray:
    """
    Trotter-evolved state as a dense vector, computed on the GPU: a TrotterAnsatz with
    ``num_steps`` layers and the angles of ``init_ansatz_to_trotter`` IS the Trotter circuit
    (what test_trotter_initial_point.py:90-96 of the reference asserts), so the evolution is one
    ``v_mul_vec``.  ``ini_state``: basis index, X positions or
    a dense vector.
    """
    from ...circuit_structures import make_trotter_like_circuit  # pylint: disable=import-outside-toplevel
    from ...core_operations import v_mul_vec  # pylint: disable=import-outside-toplevel

    circ = TrotterAnsatz(num_qubits, make_trotter_like_circuit(num_qubits, num_steps), second_order)
    thetas = init_ansatz_to_trotter(circ, np.zeros(circ.num_thetas), evol_time=evol_time, delta=delta)
    if isinstance(ini_state, np.ndarray) and ini_state.ndim == 1 and ini_state.size == 2**num_qubits:
        vec = ini_state.astype(np.complex128)
    else:
        vec = np.zeros(2**num_qubits, dtype=np.complex128)
        index = int(ini_state) if isinstance(ini_state, (int, np.integer)) else basis_index(ini_state)
        vec[index] = 1
    out = np.empty_like(vec)
    return v_mul_vec(circ, thetas, vec, out)


class Trotter:
    """
    Trotter evolution under the XXZ chain Hamiltonian, first or second order (trotter.py:40-180 of
    the reference, where it wraps a Qiskit circuit).  Here the Trotter circuit is the
    ``TrotterAnsatz`` with the angles of ``init_ansatz_to_trotter`` and runs on the GPU; an initial
    state is a dense vector, a basis index or the list of X-gate positions that the reference's
    preparation circuits (``neel_init_state`` ...) stand for.
    """

    def __init__(self, *, num_qubits: int, evol_time: float, num_steps: int, delta: float = 1.0,
                 second_order: bool):
        assert chk.is_int(num_qubits, num_qubits >= 2)
        assert chk.is_float(evol_time, evol_time > 0)
        assert chk.is_int(num_steps, num_steps >= 1)
        assert chk.is_float(delta, delta > 0)
        assert isinstance(second_order, bool)
        self._num_qubits = int(num_qubits)
        self._evol_time = float(evol_time)
        self._num_trotter_steps = int(num_steps)
        self._delta = float(delta)
        self._dt = float(evol_time) / float(num_steps)
        self._second_order = second_order

    @property
    def evol_time(self) -> float:
        return self._evol_time

    @property
    def time_step(self) -> float:
        return self._dt

    @property
    def num_trotter_steps(self) -> int:
        return self._num_trotter_steps

    def as_vector(self, ini_state: Union[int, List[int], np.ndarray]) -> np.ndarray:
        """``Trotter |ini_state>`` as a dense vector (:97-127); one ``v_mul_vec`` on the GPU."""
        return trotter_state(self._num_qubits, evol_time=self._evol_time, num_steps=self._num_trotter_steps,
                             delta=self._delta, second_order=self._second_order, ini_state=ini_state)

    def as_mps(self, ini_state: Union[int, List[int]], trunc_thr: Optional[float] = None,
               out_state: Optional[np.ndarray] = None, chi_max: int = 64):
        """
        ``Trotter |ini_state>`` in MPS format (:137-163); ``ini_state`` is a basis state.  ``chi_max``
        (<= 64) is the bond-dimension cap of the GPU engine; ``BondCapacityError`` is raised if the cap
        removes more weight than ``trunc_thr`` allows (the reference's qiskit-aer run has no cap).
        """
        from ...circuit_structures import make_trotter_like_circuit  # pylint: disable=import-outside-toplevel
        from ...mps_engine import MpsWorkspace  # pylint: disable=import-outside-toplevel
        from ... import mps_operations as mpsop  # pylint: disable=import-outside-toplevel

        n = self._num_qubits
        circ = TrotterAnsatz(n, make_trotter_like_circuit(n, self._num_trotter_steps), self._second_order)
        thetas = init_ansatz_to_trotter(circ, np.zeros(circ.num_thetas), evol_time=self._evol_time, delta=self._delta)
        index = int(ini_state) if isinstance(ini_state, (int, np.integer)) else basis_index(ini_state)
        thr = mpsop.no_truncation_threshold() if trunc_thr is None else float(trunc_thr)
        ws = MpsWorkspace(circ, num_slots=1, chi_max=chi_max, trunc_thr=thr)
        try:
            ws.set_product(0, index)
            ws.apply(thetas, 0, 0, dagger=False)
            ws.check_cap("Trotter.as_mps")
            mps = ws.download(0)
        finally:
            ws.close()
        if out_state is not None:
            np.copyto(out_state, mpsop.mps_to_vector(mps))
        return mps

    def as_qcircuit(self, ini_state):
        """The reference returns a Qiskit ``QuantumCircuit`` (:129-135); Qiskit is not part of this package."""
        raise NotImplementedError("Trotter.as_qcircuit needs Qiskit; use as_vector / as_mps, or "
                                  "init_ansatz_to_trotter for the angles of the equivalent TrotterAnsatz")


def make_hamiltonian(num_qubits: int, delta: float) -> np.ndarray:
    """
    Dense XXZ chain Hamiltonian  H = -1/4 sum_k (X_k X_{k+1} + Y_k Y_{k+1} + delta Z_k Z_{k+1})
    (half-spin matrices; trotter.py:183-230).  Testing aid, O(4^n) memory.  Built from index
    arithmetic: XX + YY exchanges the two bits of a neighbouring pair when they differ (amplitude 2),
    ZZ is +1 for equal and -1 for different bits; the chain is mirror symmetric, so the bit order is
    immaterial.
    """
    assert chk.is_int(num_qubits, num_qubits >= 2) and chk.is_float(delta)
    dim = 2**num_qubits
    idx = np.arange(dim)
    ham = np.zeros((dim, dim), dtype=np.complex128)
    for k in range(num_qubits - 1):
        b0, b1 = (idx >> k) & 1, (idx >> (k + 1)) & 1
        differ = b0 != b1
        ham[idx, idx] += -0.25 * delta * np.where(differ, -1.0, 1.0)
        flipped = idx ^ ((1 << k) | (1 << (k + 1)))
        ham[flipped[differ], idx[differ]] += -0.25 * 2.0
    return ham


def exact_evolution(hamiltonian: np.ndarray, ini_state: Union[int, List[int], np.ndarray],
                    evol_time: float) -> np.ndarray:
    """``exp(-i t H) |ini_state>`` by dense matrix exponential (trotter.py:233-266); testing aid."""
    from scipy.linalg import expm  # pylint: disable=import-outside-toplevel

    assert chk.complex_2d(hamiltonian) and hamiltonian.shape[0] == hamiltonian.shape[1]
    assert chk.is_float(evol_time, evol_time > 0)
    if isinstance(ini_state, np.ndarray) and ini_state.ndim == 1 and ini_state.size == hamiltonian.shape[0]:
        vec = ini_state.astype(np.complex128)
    else:
        vec = np.zeros(hamiltonian.shape[0], dtype=np.complex128)
        vec[int(ini_state) if isinstance(ini_state, (int, np.integer)) else basis_index(ini_state)] = 1
    return expm((-1.0j * evol_time) * hamiltonian) @ vec


def trotter_global_phase(num_qubits: int, num_steps: int, second_order: bool) -> float:
    """
    Global phase the reference attaches to a Trotter circuit (trotter.py:286-314): pi/4 per block of
    the ``num_steps`` full layers, plus pi/4 times ``num_qubits`` (even) or ``num_qubits - 1`` (odd)
    for the second order -- the reference's own count, kept as it is.  For the first order
    ``e^{i phase} Trotter|psi>`` approaches ``exp(-i t H)|psi>`` as the steps get finer
    (tests/test_trotter_cpu.py).
    """
    assert chk.is_int(num_qubits, num_qubits >= 2) and chk.is_int(num_steps, num_steps >= 1)
    assert isinstance(second_order, bool)
    blocks = (num_qubits - 1) * num_steps
    if second_order:
        blocks += num_qubits if num_qubits % 2 == 0 else num_qubits - 1
    return 0.25 * np.pi * blocks
