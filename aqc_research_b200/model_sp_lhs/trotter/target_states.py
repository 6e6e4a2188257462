"""
Computation, caching and loading of the target states of the time-evolution driver.
Reference: aqc_research/model_sp_lhs/trotter/target_states.py (TargetMpsState :44-132,
generate_all_mps_targets :135-231, get_target_mps_states :234-277, TargetClassicState :285-371,
generate_classic_target :374-455, get_target_classic_states :458-512, get_target_states :520-545).

What changes: the Trotter circuits are never built as Qiskit objects.  A ``TrotterAnsatz`` with the
angles of ``init_ansatz_to_trotter`` IS the Trotter circuit (test_trotter_initial_point.py:90-96 of
the reference), so an incremental evolution step is one ``v_mul_mps`` / ``v_mul_vec`` on the GPU.
Initial states are given as X-gate positions (``trotter.neel_init_state``), not QuantumCircuits.

On-disk format: the reference pickles lists of ``TargetMpsState`` / ``TargetClassicState`` objects
under the module path ``aqc_research.model_sp_lhs.trotter.target_states``.  ``load_targets`` maps
that path onto the classes below, ``save_targets`` writes it, so cached target files of either
code base open in the other (same attribute names, plain NumPy payloads).
"""

import io
import os
import pickle
import sys
import types
from typing import Any, List, Optional, Union
import numpy as np
from ... import checking as chk
from ...circuit_structures import make_trotter_like_circuit
from ...parametric_circuit import TrotterAnsatz
from . import trotter as trotop

REFERENCE_MODULE = "aqc_research.model_sp_lhs.trotter.target_states"


def precise_multiplier() -> int:
    """Ground-truth Trotter circuits use this many times more steps (target_states.py:30-36)."""
    return int(10)


def _trotter_ansatz(num_qubits: int, evol_time: float, num_steps: int, delta: float, second_order: bool):
    circ = TrotterAnsatz(num_qubits, make_trotter_like_circuit(num_qubits, int(num_steps)), bool(second_order))
    thetas = trotop.init_ansatz_to_trotter(
        circ, np.zeros(circ.num_thetas), evol_time=float(evol_time), delta=float(delta)
    )
    return circ, thetas


def _ini_positions(opts: Any, num_qubits: int):
    return opts.ini_state_func[0](num_qubits)


class _TargetBase:
    _FIELDS = ()

    @classmethod
    def _match(cls, dat, num_qubits, steps, time, index, opts) -> bool:
        return (
            isinstance(dat, cls)
            and all(hasattr(dat, f) for f in cls._FIELDS)
            and dat.num_qubits == num_qubits
            and dat.num_trot_steps == steps
            and dat.precise_multiplier == precise_multiplier()
            and bool(np.isclose(dat.delta / opts.delta, 1))
            and chk.is_float(dat.evol_time, bool(np.isclose(dat.evol_time / time, 1)))
            and chk.is_int(dat.my_id, dat.my_id == index)
            and isinstance(dat.second_order, bool)
        )


class TargetMpsState(_TargetBase):
    """Target state |t1> in MPS format and related data (target_states.py:44-132)."""

    _FIELDS = ("num_qubits", "num_trot_steps", "precise_multiplier", "trunc_thr", "delta", "evol_time",
               "my_id", "t1_gt", "t1", "second_order")

    def __init__(self, *, opts: Any, num_qubits: int, num_trot_steps: int, evol_time: float, my_id: int,
                 t1_gt, t1, second_order: bool):
        from ...mps_operations import check_mps  # pylint: disable=import-outside-toplevel

        assert chk.is_int(num_qubits, num_qubits >= 2)
        assert chk.is_int(num_trot_steps, num_trot_steps in opts.trotter_steps)
        assert chk.is_float(evol_time, evol_time in opts.evol_times)
        assert chk.is_int(my_id, my_id >= 0)
        assert check_mps(t1_gt) and check_mps(t1)
        assert isinstance(second_order, bool)
        self.num_qubits = int(num_qubits)
        self.num_trot_steps = int(num_trot_steps)
        self.precise_multiplier = precise_multiplier()
        self.trunc_thr = float(opts.trunc_thr_target)
        self.delta = float(opts.delta)
        self.evol_time = float(evol_time)
        self.my_id = int(my_id)
        self.t1_gt = t1_gt
        self.t1 = t1
        self.second_order = second_order

    @staticmethod
    def check_cached_data(opts: Any, num_qubits: int, data: List[Any]) -> bool:
        from ...mps_operations import check_mps  # pylint: disable=import-outside-toplevel

        assert chk.is_int(num_qubits, num_qubits >= 2) and chk.is_list(data)
        for i in range(min(len(data), len(opts.evol_times), len(opts.trotter_steps))):
            dat = data[i]
            if not (
                TargetMpsState._match(dat, num_qubits, opts.trotter_steps[i], opts.evol_times[i], i, opts)
                and bool(np.isclose(dat.trunc_thr / opts.trunc_thr_target, 1))
                and check_mps(dat.t1_gt)
                and check_mps(dat.t1)
            ):
                return False
        return True


class TargetClassicState(_TargetBase):
    """Target state |t1> as a dense vector and related data (target_states.py:285-371)."""

    _FIELDS = ("num_qubits", "num_trot_steps", "precise_multiplier", "delta", "evol_time", "my_id",
               "t1_gt", "t1", "second_order")

    def __init__(self, *, opts: Any, num_qubits: int, num_trot_steps: int, evol_time: float, my_id: int,
                 t1_gt: np.ndarray, t1: np.ndarray, second_order: bool):
        assert chk.is_int(num_qubits, num_qubits >= 2)
        assert chk.is_int(num_trot_steps, num_trot_steps in opts.trotter_steps)
        assert chk.is_float(evol_time, evol_time in opts.evol_times)
        assert chk.is_int(my_id, my_id >= 0)
        assert isinstance(t1_gt, np.ndarray) and isinstance(t1, np.ndarray)
        assert isinstance(second_order, bool)
        self.num_qubits = int(num_qubits)
        self.num_trot_steps = int(num_trot_steps)
        self.precise_multiplier = precise_multiplier()
        self.delta = float(opts.delta)
        self.evol_time = float(evol_time)
        self.my_id = int(my_id)
        self.t1_gt = t1_gt
        self.t1 = t1
        self.second_order = second_order

    @staticmethod
    def check_cached_data(opts: Any, num_qubits: int, data: List[Any]) -> bool:
        assert chk.is_int(num_qubits, num_qubits >= 2) and chk.is_list(data)
        for i in range(min(len(data), len(opts.evol_times), len(opts.trotter_steps))):
            dat = data[i]
            if not (
                TargetClassicState._match(dat, num_qubits, opts.trotter_steps[i], opts.evol_times[i], i, opts)
                and isinstance(dat.t1_gt, np.ndarray)
                and isinstance(dat.t1, np.ndarray)
            ):
                return False
        return True


# ---------------------------------------------------------------------------------------------
# pickle interoperability with the reference's result folders
# ---------------------------------------------------------------------------------------------
class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == REFERENCE_MODULE and name in ("TargetMpsState", "TargetClassicState"):
            return globals()[name]
        return super().find_class(module, name)


def load_targets(path: str) -> list:
    """Loads a target file written by this package OR by the reference."""
    with open(path, "rb") as fld:
        return _RefUnpickler(fld).load()


class _RefPickler(pickle.Pickler):
    """Writes TargetMpsState / TargetClassicState under the reference's module path."""

    def reducer_override(self, obj):
        if isinstance(obj, (TargetMpsState, TargetClassicState)):
            return _rebuild, (type(obj).__name__, dict(obj.__dict__))
        return NotImplemented


def _rebuild(name: str, state: dict):
    obj = object.__new__(globals()[name])
    obj.__dict__.update(state)
    return obj


def save_targets(data: list, path: str, reference_compatible: bool = True) -> None:
    """
    Pickles a list of targets.  With ``reference_compatible`` the classes are recorded under
    ``aqc_research.model_sp_lhs.trotter.target_states`` exactly as ``pickle.dump`` in the reference
    does (target_states.py:274-275, 509-510), so the reference loads the file with a plain
    ``pickle.load``; an alias module is registered for the duration of the dump when the reference
    package is not importable here.
    """
    if not reference_compatible:
        with open(path, "wb") as fld:
            _RefPickler(fld).dump(data)
        return
    created = []
    names = REFERENCE_MODULE.split(".")
    try:
        for i in range(1, len(names) + 1):
            mod = ".".join(names[:i])
            if mod not in sys.modules:
                sys.modules[mod] = types.ModuleType(mod)
                created.append(mod)
        alias = sys.modules[REFERENCE_MODULE]
        saved = {}
        for cls in (TargetMpsState, TargetClassicState):
            saved[cls] = (cls.__module__, getattr(alias, cls.__name__, None))
            setattr(alias, cls.__name__, cls)
            cls.__module__ = REFERENCE_MODULE
        buf = io.BytesIO()
        pickle.dump(data, buf)
        with open(path, "wb") as fld:
            fld.write(buf.getvalue())
    finally:
        for cls, (mod, prev) in saved.items():
            cls.__module__ = mod
            if prev is not None:
                setattr(sys.modules[REFERENCE_MODULE], cls.__name__, prev)
        for mod in created:
            sys.modules.pop(mod, None)


# ---------------------------------------------------------------------------------------------
# MPS targets
# ---------------------------------------------------------------------------------------------
def product_mps(num_qubits: int, x_positions) -> tuple:
    """Bond-dimension-1 MPS of the basis state with X gates at ``x_positions``."""
    bits = set(int(q) for q in x_positions)
    gam = []
    for q in range(num_qubits):
        one = np.ones((1, 1), dtype=np.complex128)
        zero = np.zeros((1, 1), dtype=np.complex128)
        gam.append((zero, one) if q in bits else (one, zero))
    lam = [np.ones(1) for _ in range(num_qubits - 1)]
    return gam, lam


def trotter_mul_mps(mps, *, num_qubits: int, evol_time: float, num_steps: int, delta: float,
                    second_order: bool, trunc_thr: float):
    """Applies a Trotter circuit (``num_steps`` steps over ``evol_time``) to an MPS on the GPU."""
    from ...mps_operations import v_mul_mps  # pylint: disable=import-outside-toplevel

    circ, thetas = _trotter_ansatz(num_qubits, evol_time, num_steps, delta, second_order)
    return v_mul_mps(circ, thetas, mps, trunc_thr=trunc_thr)


def generate_all_mps_targets(*, opts: Any, num_qubits: int, second_order: bool) -> List[TargetMpsState]:
    """
    All targets in MPS format by precise and normal Trotterisation, re-using the MPS of the previous
    horizon for the next one (target_states.py:135-231).
    """
    trotter_steps = np.asarray(opts.trotter_steps)
    evol_times = np.asarray(opts.evol_times)
    assert chk.is_int(num_qubits, num_qubits >= 2)
    assert evol_times.size == trotter_steps.size
    assert isinstance(second_order, bool)
    assert np.unique(np.diff(trotter_steps)).size <= 1, "expects uniform stepping"
    assert np.allclose(np.diff(evol_times), evol_times[0]), "expects equal intervals"
    thr = opts.trunc_thr_target
    t1_gt = product_mps(num_qubits, _ini_positions(opts, num_qubits))
    t1 = product_mps(num_qubits, _ini_positions(opts, num_qubits))
    interval, nsteps = evol_times[0], trotter_steps[0]
    targets = []
    for i in range(max(evol_times.size, trotter_steps.size)):
        if i > 0:
            interval = evol_times[i] - evol_times[i - 1]
            nsteps = trotter_steps[i] - trotter_steps[i - 1]
        common = dict(num_qubits=num_qubits, evol_time=float(interval), delta=opts.delta,
                      second_order=second_order, trunc_thr=thr)
        t1_gt = trotter_mul_mps(t1_gt, num_steps=int(nsteps) * precise_multiplier(), **common)
        t1 = trotter_mul_mps(t1, num_steps=int(nsteps), **common)
        targets.append(
            TargetMpsState(opts=opts, num_qubits=num_qubits, num_trot_steps=trotter_steps[i],
                           evol_time=evol_times[i], my_id=i, t1_gt=t1_gt, t1=t1, second_order=second_order)
        )
    return targets


def _cached_or_compute(filename, input_file, check, compute):
    if not bool(isinstance(input_file, str) and os.path.isfile(input_file)):
        input_file = filename
    if os.path.isfile(input_file):
        data = load_targets(input_file)
        if check(data):
            return data
    data = compute()
    assert check(data)
    os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
    save_targets(data, filename)
    return data


def get_target_mps_states(opts: Any, num_qubits: int, second_order: bool,
                          input_file: Optional[str] = None) -> List[TargetMpsState]:
    """Loads precomputed MPS targets or computes and stores them (target_states.py:234-277)."""
    filename = os.path.join(opts.result_dir, f"target_mps_states_n{num_qubits}.pkl")
    return _cached_or_compute(
        filename, input_file,
        lambda d: TargetMpsState.check_cached_data(opts, num_qubits, d),
        lambda: generate_all_mps_targets(opts=opts, num_qubits=num_qubits, second_order=second_order),
    )


# ---------------------------------------------------------------------------------------------
# classic targets
# ---------------------------------------------------------------------------------------------
def generate_classic_target(*, opts: Any, num_qubits: int, num_trot_steps: int, evol_time: float,
                            my_id: int, second_order: bool) -> TargetClassicState:
    """Accurate and normal Trotter states as dense vectors (target_states.py:374-455)."""
    assert chk.is_int(num_qubits, num_qubits >= 2)
    assert chk.is_int(num_trot_steps, num_trot_steps >= 1)
    assert chk.is_float(evol_time, evol_time > 0)
    assert chk.is_int(my_id, my_id >= 0) and isinstance(second_order, bool)
    ini = _ini_positions(opts, num_qubits)
    common = dict(evol_time=float(evol_time), delta=opts.delta, second_order=second_order, ini_state=ini)
    t1_gt = trotop.trotter_state(num_qubits, num_steps=int(num_trot_steps) * precise_multiplier(), **common)
    t1 = trotop.trotter_state(num_qubits, num_steps=int(num_trot_steps), **common)
    return TargetClassicState(opts=opts, num_qubits=num_qubits, num_trot_steps=num_trot_steps,
                              evol_time=evol_time, my_id=my_id, t1_gt=t1_gt, t1=t1, second_order=second_order)


def get_target_classic_states(opts: Any, num_qubits: int, second_order: bool,
                              input_file: Optional[str] = None) -> List[TargetClassicState]:
    """Loads precomputed dense targets or computes and stores them (target_states.py:458-512)."""
    filename = os.path.join(opts.result_dir, f"target_classic_states_n{num_qubits}.pkl")

    def compute():
        return [
            generate_classic_target(opts=opts, num_qubits=num_qubits, num_trot_steps=nts, evol_time=etm,
                                    my_id=i, second_order=second_order)
            for i, (nts, etm) in enumerate(zip(opts.trotter_steps, opts.evol_times))
        ]

    return _cached_or_compute(
        filename, input_file, lambda d: TargetClassicState.check_cached_data(opts, num_qubits, d), compute
    )


def get_target_states(opts: Any) -> Union[List[TargetClassicState], List[TargetMpsState]]:
    """Loads or recomputes the list of target states (target_states.py:520-545)."""
    fn = get_target_mps_states if opts.use_mps else get_target_classic_states
    return fn(opts=opts, num_qubits=opts.num_qubits, second_order=opts.second_order_trotter,
              input_file=opts.targets_file)
