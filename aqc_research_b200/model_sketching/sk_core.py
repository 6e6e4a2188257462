"""
Unitary-AQC objective ``f = 1 - Re Tr(<V Q | U Q>) / m`` and its gradient on the GPU.
Reference: aqc_research/model_sketching/sk_core.py (SketchingVectorsBase :34-91,
SketchingObjectiveEx :94-297, FullRangeSketchingVectors :300-326).

The (2^n, m) matrices live in an ``SvWorkspace`` with ``log2_cols = log2 m`` (gates act on the
row-index bits).  With ``FullRangeSketchingVectors`` nothing but thetas crosses PCIe per
evaluation: Y = U stays resident, X = I is regenerated on the device.
``BatchedSketchingObjective`` evaluates many independent starts (multistart, one slice of
starts per GPU) in one launch sequence.
"""

from abc import ABC, abstractmethod
from time import perf_counter
from typing import Tuple
import numpy as np
from .. import checking as chk
from ..engine import SvWorkspace
from ..parametric_circuit import ParametricCircuit, is_parametric_circuit

_SLOT_Y, _SLOT_Z, _SLOT_X, _SLOT_TMP = 0, 1, 2, 3


class SketchingVectorsBase(ABC):
    """Generator of sketching matrices X (2^n x m) and Y = U X (sk_core.py:34-91)."""

    def __init__(self, num_skvecs: int, target_mat: np.ndarray):
        assert chk.is_int(num_skvecs) and chk.complex_2d_square(target_mat)
        num_skvecs = min(max(num_skvecs, 1), target_mat.shape[0])
        if num_skvecs & (num_skvecs - 1):
            raise ValueError("'num_skvecs' must be a power of 2 number")
        self._num_skvecs = int(num_skvecs)
        self._target_mat = target_mat

    @property
    def num_skvecs(self) -> int:
        return self._num_skvecs

    @property
    def target_matrix(self) -> np.ndarray:
        return self._target_mat

    #: True if generate() always returns X = I and Y = U (lets the objective keep them on the GPU)
    is_full_range = False
    #: True if generate_device() is implemented (X and Y are produced in GPU slots, no host round trip)
    on_device = False

    def generate_device(self, ws: SvWorkspace, x: int, y: int, tmp: int, circ=None, thetas=None) -> None:
        """Writes X into slot ``x`` and Y = U X into slot ``y`` of ``ws`` (``tmp``: scratch slot)."""
        raise NotImplementedError("this generator has no device path")

    def _own_workspace(self) -> SvWorkspace:
        """Workspace for stand-alone generate() calls (the objective class passes its own)."""
        ws = getattr(self, "_ws_own", None)
        if ws is None:
            n = _log2(self._target_mat.shape[0])
            dummy = ParametricCircuit(n, "cx", np.array([[0], [1]], dtype=np.int64))
            ws = SvWorkspace(dummy, num_slots=4, log2_cols=_log2(self._num_skvecs), as_generic=True)
            ws.set_dense_target(self._target_mat)
            self._ws_own = ws
        return ws

    def _download_xy(self, ws: SvWorkspace, x: int, y: int):
        shape = (self._target_mat.shape[0], self._num_skvecs)
        return ws.download(x).reshape(shape), ws.download(y).reshape(shape)

    @abstractmethod
    def generate(self, circ=None, thetas=None) -> Tuple[np.ndarray, np.ndarray]:
        raise NotImplementedError("abstract method")


class FullRangeSketchingVectors(SketchingVectorsBase):
    """X = I, Y = U: the full AQC objective (sk_core.py:300-326)."""

    is_full_range = True

    def __init__(self, target_mat: np.ndarray):
        super().__init__(target_mat.shape[0], target_mat)

    def generate(self, circ=None, thetas=None) -> Tuple[np.ndarray, np.ndarray]:
        dim = self._target_mat.shape[0]
        return np.eye(dim, dtype=np.complex128), np.array(self._target_mat, dtype=np.complex128)


class RandomSketchingVectors(SketchingVectorsBase):
    """
    New random sketching vectors upon every request (sk_core.py:329-359):
    ``X = qr(rand + 1j rand)``, ``Y = U X``.  The random matrix comes from the global NumPy RNG in
    the reference's call order; QR and the GEMM run on the GPU.  X equals the reference's up to a
    unitary m x m factor (column space identical), which the objective and gradient do not see.
    """

    on_device = True

    def __init__(self, num_skvecs: int, target_mat: np.ndarray):
        super().__init__(num_skvecs, target_mat)
        assert target_mat.shape[0] % self.num_skvecs == 0

    def generate_device(self, ws, x, y, tmp, circ=None, thetas=None) -> None:
        dim, m = self._target_mat.shape[0], self.num_skvecs
        ws.upload(x, np.random.rand(dim, m) + 1j * np.random.rand(dim, m))
        ws.orthonormalize(x, tmp)
        ws.target_matmul(x, y)

    def generate(self, circ=None, thetas=None):
        ws = self._own_workspace()
        self.generate_device(ws, 0, 1, 2)
        return self._download_xy(ws, 0, 1)


class AlternatingSketchingVectors(SketchingVectorsBase):
    """A random subset of the target's columns per request, cycling through a permutation (sk_core.py:362-407)."""

    on_device = True

    def __init__(self, num_skvecs: int, target_mat: np.ndarray):
        super().__init__(num_skvecs, target_mat)
        dim = target_mat.shape[0]
        assert dim % self.num_skvecs == 0
        self._offset = 0
        self._indices = np.random.permutation(dim)

    def generate_device(self, ws, x, y, tmp, circ=None, thetas=None) -> None:
        dim = self._target_mat.shape[0]
        if self._offset >= dim:
            self._offset = 0
            self._indices = np.random.permutation(dim)
        idx = self._indices[self._offset : self._offset + self.num_skvecs]
        ws.gather_target_columns(idx, x, y)
        self._offset += self.num_skvecs

    def generate(self, circ=None, thetas=None):
        ws = self._own_workspace()
        self.generate_device(ws, 0, 1, 2)
        return self._download_xy(ws, 0, 1)


class EigenSketchingVectors(SketchingVectorsBase):
    """
    Sketching vectors spanning the dominant range of ``V^H - U^H`` (randomised range finder,
    sk_core.py:410-462): ``X = qr((V^H - U^H) Omega)``, ``Omega`` complex normal, ``Y = U X``.
    """

    on_device = True

    def generate_device(self, ws, x, y, tmp, circ=None, thetas=None) -> None:
        assert is_parametric_circuit(circ)
        assert chk.float_1d(thetas, thetas.size == circ.num_thetas)
        assert circ.dimension == self._target_mat.shape[0]
        dim, m = circ.dimension, self.num_skvecs
        omega = 1j * np.random.randn(dim, m)  # same draw order as the reference (:435-438)
        omega += np.random.randn(dim, m)
        ws.upload(x, omega)
        ws.target_matmul(x, tmp, conj_transpose=True)  # U^H Omega
        ws.apply(thetas, x, x, dagger=True)  # V^H Omega
        ws.sub(x, tmp)
        ws.orthonormalize(x, tmp)
        ws.target_matmul(x, y)

    def generate(self, circ=None, thetas=None):
        ws = getattr(self, "_ws_own", None)
        if ws is None or ws.circuit.signature() != _signature(circ):
            ws = SvWorkspace(circ, num_slots=4, log2_cols=_log2(self.num_skvecs), as_generic=True)
            ws.set_dense_target(self._target_mat)
            self._ws_own = ws
        self.generate_device(ws, 0, 1, 2, circ, thetas)
        return self._download_xy(ws, 0, 1)


def skvecs_generator(skvecs_type: str, num_skvecs: int, target_mat: np.ndarray) -> SketchingVectorsBase:
    """Factory of the reference (sk_core.py:465-497): one of 'full', 'rand', 'alt', 'eigen'."""
    assert isinstance(skvecs_type, str)
    if skvecs_type == "full" or num_skvecs == target_mat.shape[0]:
        return FullRangeSketchingVectors(target_mat)
    if skvecs_type == "rand":
        return RandomSketchingVectors(num_skvecs, target_mat)
    if skvecs_type == "alt":
        return AlternatingSketchingVectors(num_skvecs, target_mat)
    if skvecs_type == "eigen":
        return EigenSketchingVectors(num_skvecs, target_mat)
    raise ValueError(
        f"unknown type of sketching vectors generator, expects one of: "
        f"['full', 'rand', 'alt', 'eigen'], got {skvecs_type}"
    )


def _signature(circ):
    from ..engine import CircuitHandle

    return CircuitHandle(circ, as_generic=True).signature()


def _log2(m: int) -> int:
    k = int(m).bit_length() - 1
    assert 1 << k == m
    return k


class SketchingObjectiveEx:
    """Drop-in for the reference class of the same name (sk_core.py:94-297)."""

    def __init__(
        self,
        circ: ParametricCircuit,
        skvecs: SketchingVectorsBase,
        *,
        enable_stats: bool = False,
        grad_scaler=None,
        stop_timeout=None,
        stop_stagnant=None,
        stop_small_fobj=None,
        logger=None,
        device: int = 0,
    ):
        assert is_parametric_circuit(circ) and isinstance(skvecs, SketchingVectorsBase)
        self._circ = circ
        self._skvecs = skvecs
        self._target = skvecs.target_matrix
        self._enable_stats = bool(enable_stats)
        self._grad_scaler = grad_scaler
        self._stop_timeout = stop_timeout
        self._stop_stagnant = stop_stagnant
        self._stop_small_fobj = stop_small_fobj
        self._logger = logger
        self._ws = SvWorkspace(
            circ, num_slots=4 if skvecs.on_device else 3, device=device,
            log2_cols=_log2(skvecs.num_skvecs), as_generic=True
        )
        if skvecs.is_full_range:
            self._ws.upload(_SLOT_Y, np.ascontiguousarray(self._target, dtype=np.complex128))
        elif skvecs.on_device:
            self._ws.set_dense_target(self._target)
        self._fobj_best = float(np.inf)
        self._thetas_best = np.zeros(circ.num_thetas)
        self._nit = 0
        self._fobj_profile = []
        self._fobj_latest = float(1e30)
        self._grad_latest = np.empty(0)
        self._thetas_latest = np.empty(0)
        self._elapsed_time = perf_counter()
        self._period = int(round(10 + 60.0 / (1 + 2.0 ** (6 - circ.num_qubits))))

    def objective_and_gradient(self, thetas: np.ndarray) -> Tuple[float, np.ndarray]:
        now = perf_counter()
        if self._elapsed_time + self._period < now:
            print(".", end="", flush=True)
            self._elapsed_time = now
        ws, m = self._ws, self._skvecs.num_skvecs
        if self._skvecs.is_full_range:
            ws.set_identity(_SLOT_X)
        elif self._skvecs.on_device:
            self._skvecs.generate_device(ws, _SLOT_X, _SLOT_Y, _SLOT_TMP, self._circ, thetas)
        else:
            x, y = self._skvecs.generate(self._circ, thetas)
            ws.upload(_SLOT_X, x)
            ws.upload(_SLOT_Y, y)
        ws.apply(thetas, _SLOT_Y, _SLOT_Z, dagger=True)  # vh_y = V^H y
        fobj = float(1 - np.real(ws.vdot(_SLOT_X, _SLOT_Z)[0]) / m)
        grad = ws.grad(thetas, x_slot=_SLOT_X, z0=_SLOT_Z, w=_SLOT_X, z=_SLOT_Z)[0]
        grad = np.ascontiguousarray(-np.real(grad) / m)
        if self._grad_scaler:
            grad *= self._grad_scaler.estimate(fobj)
        if fobj < self._fobj_best:
            self._fobj_best = fobj
            np.copyto(self._thetas_best, thetas)
        self._nit += 1
        if self._enable_stats:
            self._fobj_profile.append(fobj)
        if self._logger is not None:
            print(f"\riter: {self._nit:4d}, fobj: {fobj:0.4f}, |grad|: {np.linalg.norm(grad):0.5f}")
        if self._stop_timeout:
            self._stop_timeout.check()
        if self._stop_stagnant:
            self._stop_stagnant.check(fobj=fobj, iter_no=self._nit)
        if self._stop_small_fobj:
            self._stop_small_fobj.check(fobj=fobj)
        return fobj, grad

    def objective(self, thetas: np.ndarray) -> float:
        self._thetas_latest = np.array(thetas, dtype=np.float64, copy=True)
        self._fobj_latest, self._grad_latest = self.objective_and_gradient(thetas)
        return self._fobj_latest

    def gradient(self, thetas: np.ndarray) -> np.ndarray:
        tol = float(10.0 * np.finfo(np.float64).eps)
        last = self._thetas_latest
        if last.size != thetas.size or not np.allclose(thetas, last, atol=tol, rtol=tol):
            self.objective(thetas)
        return self._grad_latest

    @property
    def statistics(self) -> dict:
        return {"convergence_profile": np.asarray(self._fobj_profile, dtype=np.float32), "nit": self._nit}

    @property
    def num_iterations(self) -> int:
        return int(self._nit)

    @property
    def optim_results(self) -> dict:
        return {
            "cost": float(self._fobj_best),
            "num_fun_ev": int(self._nit),
            "num_grad_ev": int(self._nit),
            "num_iters": int(self._nit),
            "thetas": self._thetas_best,
            "entangler": self._circ.entangler,
            "blocks": self._circ.blocks.copy(),
        }

    def set_status_trackers(self, timeout, stopper):
        """Compatibility with AqcOptimizer (sk_core.py:295-297)."""

    @property
    def workspace(self) -> SvWorkspace:
        return self._ws


class BatchedSketchingObjective:
    """
    Full-range objective and gradient for ``batch`` independent angle sets at once -- the
    multistart fan-out of aqc_sketching.py:266-272 mapped onto one GPU (one slice of starts per
    GPU, no communication).  ``evaluate(thetas[batch, T])`` -> (f[batch], grad[batch, T]).
    """

    def __init__(self, circ: ParametricCircuit, target_mat: np.ndarray, batch: int, device: int = 0):
        assert chk.complex_2d_square(target_mat, target_mat.shape[0] == circ.dimension)
        self._circ = circ
        self._dim = circ.dimension
        self._batch = int(batch)
        self._ws = SvWorkspace(
            circ, num_slots=3, device=device, log2_cols=circ.num_qubits, batch=batch, as_generic=True
        )
        self._ws.upload(_SLOT_Y, np.ascontiguousarray(target_mat, dtype=np.complex128))

    def evaluate(self, thetas: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        thetas = np.ascontiguousarray(thetas, dtype=np.float64).reshape(self._batch, self._circ.num_thetas)
        ws = self._ws
        ws.set_identity(_SLOT_X)
        ws.apply(thetas, _SLOT_Y, _SLOT_Z, dagger=True)
        fobj = 1 - np.real(ws.vdot(_SLOT_X, _SLOT_Z)) / self._dim
        grad = ws.grad(thetas, x_slot=_SLOT_X, z0=_SLOT_Z, w=_SLOT_X, z=_SLOT_Z)
        return fobj, -np.real(grad) / self._dim

    @property
    def workspace(self) -> SvWorkspace:
        return self._ws
