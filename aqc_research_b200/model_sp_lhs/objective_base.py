"""
GPU-resident base of the state-preparation objectives.
Reference: aqc_research/model_sp_lhs/objective_base.py (state handlers :42-429, SpService
:437-621, SpLHSObjectiveBase :630-833).  Public surface and call protocol are the reference's;
what changed is WHERE the data live: target, V^H target and the sweep vectors w, z stay in HBM
inside an ``SvWorkspace``; only thetas go down and (hs, gradient) come back per call.
"""

import itertools
from abc import ABC, abstractmethod
from typing import Callable, List, Optional, Tuple
import numpy as np
from .. import checking as chk
from ..engine import SvWorkspace, circuit_signature
from ..parametric_circuit import ParametricCircuit, is_parametric_circuit

_SQRT_EPS = float(np.sqrt(np.finfo(np.float64).eps))  # theta-cache tolerance (objective_base.py:728-730)

# workspace slot roles
SLOT_TARGET, SLOT_VH_TARGET, SLOT_W, SLOT_Z, SLOT_STATE = 0, 1, 2, 3, 4


def flip_combinations(num_qubits: int, max_flips: int) -> Tuple[List[List[Tuple]], int]:
    """All subsets of 1..max_flips qubits, grouped by size; and the state count incl. |0>."""
    combos = [list(itertools.combinations(range(num_qubits), k)) for k in range(1, max_flips + 1)]
    return combos, 1 + sum(len(c) for c in combos)


class BasisStateHandler:
    """
    States S|0>, S X_i|0>, S X_i X_j|0>, ... where S is a product of X gates (possibly empty):
    every state is a computational-basis vector, so <state|v> is a single gathered amplitude
    (ThinStateHandler, objective_base.py:42-256; little-endian: qubit q <-> bit q, :86-87).
    ``init_index`` is the basis index of S|0> (0 for the reference's ThinStateHandler,
    bits 0,2,4,.. set for the Neel state of trotter.py:389-398).
    """

    def __init__(self, num_qubits: int, max_flips: int, init_index: int = 0, verbose: bool = False):
        assert chk.is_int(num_qubits, num_qubits >= 2)
        assert chk.is_int(max_flips, 0 <= max_flips <= num_qubits)
        assert 0 <= init_index < 2**num_qubits
        self._num_qubits = num_qubits
        self._combos, num_states = flip_combinations(num_qubits, max_flips)
        idx = [int(init_index)]
        for group in self._combos:
            for subset in group:
                mask = 0
                for q in subset:
                    mask |= 1 << q
                idx.append(int(init_index) ^ mask)
        self._state_idx = np.asarray(idx, dtype=np.int64)
        assert self._state_idx.size == num_states

    @property
    def num_states(self) -> int:
        return int(self._state_idx.size)

    @property
    def state_indices(self) -> np.ndarray:
        return self._state_idx

    @property
    def flip_qubit_positions(self):
        return self._combos

    def init_state(self, state_no: int) -> np.ndarray:
        """Dense host copy of a state (for callers that want the vector itself)."""
        vec = np.zeros(2**self._num_qubits, dtype=np.complex128)
        vec[self._state_idx[state_no]] = 1
        return vec

    @property
    def state0(self) -> np.ndarray:
        return self.init_state(0)

    def state_dot_vector(self, state_no: int, vec: np.ndarray) -> np.complex128:
        return vec[self._state_idx[state_no]]


class DenseStateHandler:
    """
    n+1 explicit dense states S|0>, S X_i|0> (GenericStateHandler, objective_base.py:258-342).
    The states are supplied as an (n+1, 2^n) array and uploaded to the GPU on demand.
    """

    def __init__(self, states: np.ndarray):
        assert chk.complex_2d(states)
        self._states = np.ascontiguousarray(states)

    @property
    def num_states(self) -> int:
        return int(self._states.shape[0])

    def init_state(self, state_no: int) -> np.ndarray:
        return self._states[state_no]

    @property
    def state0(self) -> np.ndarray:
        return self._states[0]

    def state_dot_vector(self, state_no: int, vec: np.ndarray) -> np.complex128:
        return np.complex128(np.vdot(self._states[state_no], vec))


def make_state_handler(num_qubits: int, max_flips: int, state_prep_func: Optional[Callable]):
    """
    ``state_prep_func(num_qubits)`` may return: an int (basis index of S|0>), a sequence of
    qubits carrying an X gate, an (n+1, 2^n) array of dense states, or -- when Qiskit is
    installed -- a QuantumCircuit as in the reference (objective_base.py:297-303).
    """
    if state_prep_func is None:
        return BasisStateHandler(num_qubits, max_flips, 0)
    prep = state_prep_func(num_qubits)
    if isinstance(prep, (int, np.integer)):
        return BasisStateHandler(num_qubits, max_flips, int(prep))
    if isinstance(prep, np.ndarray) and prep.ndim == 2:
        return DenseStateHandler(prep.astype(np.complex128))
    if isinstance(prep, (list, tuple, np.ndarray)):
        index = 0
        for q in prep:
            index ^= 1 << int(q)
        return BasisStateHandler(num_qubits, max_flips, index)
    # Qiskit QuantumCircuit (optional dependency)
    try:
        from qiskit import QuantumCircuit  # pylint: disable=import-outside-toplevel
        from qiskit.quantum_info import Statevector  # pylint: disable=import-outside-toplevel
    except ImportError as ex:
        raise TypeError("unsupported return type of state_prep_func (Qiskit is not installed)") from ex
    if max_flips > 1:
        raise ValueError("expects 'max_flips <= 1' to save memory")
    states = np.zeros((num_qubits + 1, 2**num_qubits), dtype=np.complex128)
    for i in range(num_qubits + 1):
        qc = QuantumCircuit(num_qubits)
        if i > 0:
            qc.x(i - 1)
        states[i] = Statevector(qc.compose(prep)).data
    return DenseStateHandler(states)


class SpService:
    """
    Counters, optional statistics and the early-termination hooks
    (objective_base.py:437-621).  ``TimeoutChecker`` / ``EarlyStopper`` objects are used through
    their ``check`` methods only; the exceptions they raise (StopIteration / TimeoutError) are the
    reference's early-stop protocol and propagate unchanged through ``gradient()``.
    """

    def __init__(self, user_parameters: dict, circuit: ParametricCircuit, num_states: int, verbose: bool = False):
        assert chk.is_dict(user_parameters) and is_parametric_circuit(circuit)
        assert chk.is_int(num_states, num_states >= 1)
        self._params = user_parameters
        self._circuit = circuit
        self._num_states = num_states
        self._verbose = bool(verbose)
        self._num_fun_ev = 0
        self._num_grad_ev = 0
        self._timeout_checker = None
        self._early_stopper = None
        self._stats = {}
        if user_parameters.get("enable_optim_stats", False):
            self._stats = {
                "hs2": np.empty((0, num_states), dtype=np.float16),
                "weight": np.empty(0, dtype=np.float16),
                "fobj": np.empty(0, dtype=np.float32),
                "grad": np.empty(0, dtype=np.float32),
                "num_fun_ev": 0,
                "num_grad_ev": 0,
            }

    def set_status_trackers(self, timeout=None, stopper=None):
        self._timeout_checker = timeout
        self._early_stopper = stopper

    @property
    def statistics(self) -> dict:
        return self._stats

    def _on_stop(self, fobj: float, thetas: np.ndarray) -> dict:
        return {
            "cost": fobj,
            "num_fun_ev": self._num_fun_ev,
            "num_grad_ev": self._num_grad_ev,
            "num_iters": self._num_grad_ev,
            "thetas": thetas.copy(),
            "blocks": self._circuit.blocks.copy(),
        }

    def on_begin_gradient(self, fobj: float, thetas: np.ndarray, fidelity: Optional[float] = None):
        if self._timeout_checker:
            self._timeout_checker.check(fobj, thetas, self._on_stop)
        if self._early_stopper:
            self._early_stopper.check(
                fobj=fobj, fidelity=fidelity, thetas=thetas, iter_no=self._num_grad_ev, on_stop=self._on_stop
            )

    def on_end_objective(self):
        self._num_fun_ev += 1

    def on_end_gradient(self, fobj: float, fidelity: float, grad: np.ndarray, hs2: np.ndarray, weight: float):
        assert chk.float_1d(grad) and chk.float_1d(hs2, hs2.size == self._num_states)
        self._num_grad_ev += 1
        if self._stats:
            sts = self._stats
            sts["hs2"] = np.concatenate((sts["hs2"], hs2.astype(np.float16)[None, :]), axis=0)
            sts["weight"] = np.append(sts["weight"], np.float16(weight))
            sts["fobj"] = np.append(sts["fobj"], np.float32(fobj))
            sts["grad"] = np.append(sts["grad"], np.float32(np.linalg.norm(grad)))
            sts["num_fun_ev"] = self._num_fun_ev
            sts["num_grad_ev"] = self._num_grad_ev
            sts["num_iters"] = self._num_grad_ev
        if self._params.get("verbose", 0) and self._num_grad_ev % max(1, self._params.get("maxiter", 50) // 50) == 0:
            print(".", end="", flush=True)

    def on_epoch_end(self):
        if self._stats:
            sts = self._stats
            gap = np.full((1, self._num_states), np.nan, dtype=np.float16)
            sts["hs2"] = np.concatenate((sts["hs2"], gap), axis=0)
            sts["weight"] = np.append(sts["weight"], np.float16(np.nan))
            sts["fobj"] = np.append(sts["fobj"], np.float32(np.nan))
            sts["grad"] = np.append(sts["grad"], np.float32(np.nan))


class SpLHSObjectiveBase(ABC):
    """
    Common state of the state-vector objectives (objective_base.py:630-833): target and
    V^H target on the GPU, theta cache, surrogate weight, service object.
    ``user_parameters`` keys used: num_qubits, max_flips, state_prep_func, enable_optim_stats,
    verbose, maxiter, and optionally ``device`` (CUDA ordinal, default 0).
    """

    def __init__(self, user_parameters: dict, circuit: ParametricCircuit, use_mps: bool = False, verbose: bool = False):
        assert isinstance(user_parameters, dict) and is_parametric_circuit(circuit)
        self._params = user_parameters
        self._circuit = circuit
        self._use_mps = bool(use_mps)
        self._verbose = bool(verbose)
        self._target = None
        self._target_seed = None  # set_target_random: the target is regenerated after a structure change
        self._blocks_seen = circuit.blocks
        self._last_thetas = np.empty(0)
        num_qubits = user_parameters["num_qubits"]
        assert num_qubits == circuit.num_qubits
        max_flips = user_parameters["max_flips"]
        self._device = int(user_parameters.get("device", 0))
        self._ws: Optional[SvWorkspace] = None
        if not use_mps:
            self._state_handler = make_state_handler(
                num_qubits, max_flips, user_parameters.get("state_prep_func", None)
            )
            self._num_states = self._state_handler.num_states
            self._dense = isinstance(self._state_handler, DenseStateHandler)
            self._ws = SvWorkspace(circuit, num_slots=5 if self._dense else 4, device=self._device)
            self._structure = circuit_signature(circuit)
            self._init_common()
        else:
            self._num_states = num_qubits + 1  # finalised by the MPS subclass
            self._structure = circuit_signature(circuit)

    def _init_common(self):
        self._service = SpService(self._params, self._circuit, self._num_states, verbose=self._verbose)
        self._hs2 = np.zeros(self._num_states)
        self._fobj = 1.0
        self._weight = 1.0

    # -- GPU residency ----------------------------------------------------------------------
    def _structure_changed(self) -> bool:
        """True once after ``insert_unit_blocks`` / ``update_structure`` changed the circuit."""
        blocks = self._circuit.blocks
        if blocks is self._blocks_seen:  # both code bases REPLACE the array on a structure change
            return False
        self._blocks_seen = blocks
        sig = circuit_signature(self._circuit)
        if sig == self._structure:
            return False
        self._structure = sig
        self._last_thetas = np.empty(0)  # cached objective belongs to the old structure
        self._early_thetas = None
        return True

    def _structure_stale(self) -> bool:
        """Cheap test (object identity of ``blocks``) whether a structure change may have happened."""
        return self._circuit.blocks is not self._blocks_seen

    def _refresh_workspace(self):
        """Re-creates the GPU workspace if the circuit structure changed (insert_unit_blocks)."""
        if self._structure_changed():
            self._ws.close()
            self._ws = SvWorkspace(self._circuit, num_slots=5 if self._dense else 4, device=self._device)
            if self._target_seed is not None:
                self._ws.fill_random(SLOT_TARGET, self._target_seed)
            elif self._target is not None:
                self._ws.upload(SLOT_TARGET, self._target)

    def _hs_products(self, thetas: np.ndarray) -> np.ndarray:
        """z0 = V^H target (kept in SLOT_VH_TARGET); returns <state_i|z0> for all states."""
        if self._target is None:
            raise RuntimeError("set_target() must be called before objective()")
        self._refresh_workspace()
        ws = self._ws
        if not self._dense:
            return ws.objective(thetas, SLOT_TARGET, SLOT_VH_TARGET, self._state_handler.state_indices)[0]
        ws.apply(thetas, SLOT_TARGET, SLOT_VH_TARGET, dagger=True)
        hs = np.zeros(self._num_states, dtype=np.complex128)
        for i in range(self._num_states):
            ws.upload(SLOT_STATE, self._state_handler.init_state(i))
            hs[i] = ws.vdot(SLOT_STATE, SLOT_VH_TARGET)[0]
        return hs

    def _raw_gradient(self, thetas: np.ndarray, state_no: int) -> np.ndarray:
        """Complex gradient of <V state|target> given the cached V^H target."""
        ws = self._ws
        if not self._dense:
            idx = int(self._state_handler.state_indices[state_no])
            return ws.grad(thetas, x_basis=idx, z0=SLOT_VH_TARGET, w=SLOT_W, z=SLOT_Z)[0]
        ws.upload(SLOT_STATE, self._state_handler.init_state(state_no))
        return ws.grad(thetas, x_slot=SLOT_STATE, z0=SLOT_VH_TARGET, w=SLOT_W, z=SLOT_Z)[0]

    # -- theta cache (objective_base.py:705-734) ------------------------------------------------
    def _store_latest_thetas(self, thetas: np.ndarray):
        self._last_thetas = np.array(thetas, dtype=np.float64, copy=True)

    def _calc_objective_before_gradient(self, thetas: np.ndarray):
        if not self._use_mps:
            self._refresh_workspace()  # a structure change invalidates the cached objective
        last = self._last_thetas
        if last.size == thetas.size:
            if np.array_equal(thetas, last):  # the optimiser's fun(x), jac(x) pair: nothing to do
                return
            tol = _SQRT_EPS
            if np.allclose(thetas, last, atol=tol, rtol=tol):
                return
        self.objective(thetas)

    @abstractmethod
    def objective(self, thetas: np.ndarray) -> float:
        raise NotImplementedError()

    @abstractmethod
    def gradient(self, thetas: np.ndarray) -> np.ndarray:
        raise NotImplementedError()

    def set_status_trackers(self, timeout=None, stopper=None):
        self._service.set_status_trackers(timeout, stopper)

    @property
    def num_thetas(self) -> int:
        return self._circuit.num_thetas

    @property
    def num_states(self) -> int:
        return self._num_states

    @property
    def target(self):
        return self._target

    def set_target(self, target) -> None:
        """Target state: complex128 vector of 2^n entries; uploaded to HBM once."""
        assert not self._use_mps
        assert chk.complex_1d(target, target.size == self._circuit.dimension)
        self._target = target
        self._target_seed = None
        self._refresh_workspace()
        self._ws.upload(SLOT_TARGET, target)
        self._last_thetas = np.empty(0)  # cached V^H target is stale
        self._early_thetas = None  # an early gradient sweep (if any) belongs to the old target

    def set_target_random(self, seed: int) -> None:
        """Synthetic target generated on the device (distribution of utils.rand_state)."""
        self._target = "device-random"
        self._target_seed = int(seed)
        self._refresh_workspace()
        self._ws.fill_random(SLOT_TARGET, seed)
        self._last_thetas = np.empty(0)
        self._early_thetas = None

    @property
    def statistics(self) -> dict:
        return self._service.statistics

    def on_epoch_end(self):
        self._service.on_epoch_end()

    @property
    def workspace(self) -> Optional[SvWorkspace]:
        """The GPU workspace (bench/profiling introspection)."""
        return self._ws


# ------------------------------------------------------------------------------------------------
# State handlers under the names of the reference (objective_base.py:42-429), for code that
# constructs them directly.  The objective classes above use make_state_handler().
# ------------------------------------------------------------------------------------------------
def _check_coefs(coefs: np.ndarray, count: int) -> None:
    assert chk.complex_or_float_1d(coefs, coefs.size == count)
    assert abs(np.linalg.norm(coefs) - 1) < np.sqrt(np.finfo(np.float64).eps)


class ThinStateHandler(BasisStateHandler):
    """
    |0> and the flip states X_i|0>, X_i X_j|0>, ... kept as basis indices and materialised on demand
    (objective_base.py:42-256), including the sparse linear combinations of the reference.
    """

    def __init__(self, num_qubits: int, max_flips: int, verbose: bool = False):
        super().__init__(num_qubits, max_flips, 0, verbose=verbose)
        self._state = np.zeros(2**num_qubits, dtype=np.complex128)

    def init_state(self, state_no: int) -> np.ndarray:
        """The internal array re-initialised to the requested state (:95-112)."""
        assert chk.is_int(state_no, 0 <= state_no < self.num_states)
        self._state.fill(0)
        self._state[self._state_idx[state_no]] = 1
        return self._state

    def init_composite_state_no_zero(self, coefs: np.ndarray) -> np.ndarray:
        """sum_i coefs_i |flip state i>, |0> excluded (:119-140)."""
        _check_coefs(coefs, self.num_states - 1)
        self._state.fill(0)
        self._state[self._state_idx[1:]] = coefs
        return self._state

    def init_composite_state(self, coefs: np.ndarray) -> np.ndarray:
        """sum_i coefs_i |state i> over all states (:142-162)."""
        _check_coefs(coefs, self.num_states)
        self._state.fill(0)
        self._state[self._state_idx] = coefs
        return self._state

    def state_dot_vector(self, state_no: int, vec: np.ndarray) -> np.complex128:
        assert chk.is_int(state_no, 0 <= state_no < self.num_states)
        assert chk.complex_1d(vec, vec.size == 2**self._num_qubits)
        return vec[self._state_idx[state_no]]

    def composite_state_dot_vector_no_zero(self, coefs: np.ndarray, vec: np.ndarray) -> np.complex128:
        """<composite state|vec> over the few non-zero entries (:175-195)."""
        _check_coefs(coefs, self.num_states - 1)
        assert chk.complex_1d(vec, vec.size == 2**self._num_qubits)
        return np.complex128(np.vdot(coefs, vec[self._state_idx[1:]]))

    def composite_state_dot_vector(self, coefs: np.ndarray, vec: np.ndarray) -> np.complex128:
        """(:197-216)"""
        _check_coefs(coefs, self.num_states)
        assert chk.complex_1d(vec, vec.size == 2**self._num_qubits)
        return np.complex128(np.vdot(coefs, vec[self._state_idx]))


def _prepared_handler(num_qubits: int, max_flips: int, state_prep_func: Optional[Callable]):
    assert chk.is_int(num_qubits, num_qubits >= 2)
    assert chk.is_int(max_flips, 0 <= max_flips <= num_qubits)
    assert state_prep_func is None or callable(state_prep_func)
    if max_flips > 1:
        raise ValueError("expects 'max_flips <= 1' to save memory")
    return make_state_handler(num_qubits, max_flips, state_prep_func)


class GenericStateHandler(DenseStateHandler):
    """
    Explicit dense states S|0>, S X_i|0> (objective_base.py:258-342).  ``state_prep_func(num_qubits)``
    returns what make_state_handler() accepts: a basis index, the X-gate positions, an (n+1, 2^n)
    array, or a Qiskit circuit when Qiskit is installed.
    """

    def __init__(self, num_qubits: int, max_flips: int, state_prep_func: Optional[Callable] = None,
                 verbose: bool = False):
        inner = _prepared_handler(num_qubits, max_flips, state_prep_func)
        states = np.array([inner.init_state(i) for i in range(inner.num_states)], dtype=np.complex128)
        super().__init__(states)

    def state_dot_vector(self, state_no: int, vec: np.ndarray) -> np.complex128:
        assert chk.is_int(state_no, 0 <= state_no < self.num_states)
        return super().state_dot_vector(state_no, vec)

    def init_composite_state_no_zero(self, _: np.ndarray) -> np.ndarray:
        raise NotImplementedError("composite states exist for ThinStateHandler only, as in the reference")

    init_composite_state = init_composite_state_no_zero

    def composite_state_dot_vector_no_zero(self, _: np.ndarray, __: np.ndarray) -> np.complex128:
        raise NotImplementedError("composite states exist for ThinStateHandler only, as in the reference")

    composite_state_dot_vector = composite_state_dot_vector_no_zero


class MpsStateHandler:
    """
    The same n+1 states in MPS format (objective_base.py:345-429).  Preparations that are products of
    X gates give bond-dimension-1 states, written down directly; ``state_dot_vector`` is one
    ``mps_dot`` on the GPU.
    """

    def __init__(self, num_qubits: int, max_flips: int, state_prep_func: Optional[Callable] = None,
                 verbose: bool = False):
        inner = _prepared_handler(num_qubits, max_flips, state_prep_func)
        if not isinstance(inner, BasisStateHandler):
            raise ValueError("MpsStateHandler supports basis-state (X-type) preparations only")
        self._num_qubits = num_qubits
        self._state_idx = inner.state_indices
        one, zero = np.ones((1, 1), dtype=np.complex128), np.zeros((1, 1), dtype=np.complex128)
        self._states = []
        for index in self._state_idx:
            gam = [((zero, one) if (int(index) >> q) & 1 else (one, zero)) for q in range(num_qubits)]
            lam = [np.ones(1, dtype=np.float64) for _ in range(num_qubits - 1)]
            self._states.append((gam, lam))

    @property
    def num_states(self) -> int:
        return len(self._states)

    @property
    def state_indices(self) -> np.ndarray:
        return self._state_idx

    def init_state(self, state_no: int):
        assert chk.is_int(state_no, 0 <= state_no < self.num_states)
        return self._states[state_no]

    @property
    def state0(self):
        return self._states[0]

    def state_dot_vector(self, state_no: int, vec) -> np.complex128:
        from ..mps_operations import mps_dot  # pylint: disable=import-outside-toplevel

        assert chk.is_int(state_no, 0 <= state_no < self.num_states)
        return np.complex128(mps_dot(self._states[state_no], vec))

    def init_composite_state_no_zero(self, _: np.ndarray) -> np.ndarray:
        raise NotImplementedError("composite states exist for ThinStateHandler only, as in the reference")

    init_composite_state = init_composite_state_no_zero

    def composite_state_dot_vector_no_zero(self, _: np.ndarray, __: np.ndarray) -> np.complex128:
        raise NotImplementedError("composite states exist for ThinStateHandler only, as in the reference")

    composite_state_dot_vector = composite_state_dot_vector_no_zero
