"""
Surrogate state-preparation objective with a max-projection term, GPU edition.
Reference: aqc_research/model_sp_lhs/objective_lhs_sur_max.py:42-196.

    hs_i   = <state_i | V^H | target>,   i = 0..num_states-1
    f      = 1 - (1 - w) |hs_0|^2 - w |hs_max|^2            (w = weight, max = leading flip state)
    grad f = Re(-2 (1-w) conj(hs_0) d<V s_0|t>) + Re(-2 w conj(hs_max) d<V s_max|t>)

The O(num_thetas) scalar logic (hysteresis on ``max_no``, weight smoothing, amplifier, stats,
stoppers) runs on the host exactly as in the reference; the vector work runs on the GPU.
"""

from typing import Optional, Tuple
import numpy as np
from .. import checking as chk
from ..core_operations import mask_gradient
from ..parametric_circuit import ParametricCircuit
from .objective_base import SLOT_STATE, SLOT_TARGET, SLOT_VH_TARGET, SLOT_W, SLOT_Z, SpLHSObjectiveBase


class SpSurrogateObjectiveMax(SpLHSObjectiveBase):
    """Drop-in for the reference class of the same name."""

    _gamma = 0.1  # exponential smoothing rate of the weight (:40)

    def __init__(
        self,
        *,
        user_parameters: dict,
        circ: ParametricCircuit,
        block_range: Optional[Tuple[int, int]] = None,
        front_layer: bool = False,
        verbose: bool = False,
        grad_scaler=None,
    ):
        super().__init__(user_parameters, circ, verbose=verbose)
        block_range = (0, circ.num_blocks) if block_range is None else block_range
        assert chk.is_tuple(block_range, len(block_range) == 2)
        assert 0 <= block_range[0] < block_range[1] <= circ.num_blocks
        assert chk.is_bool(front_layer)
        assert grad_scaler is None or hasattr(grad_scaler, "estimate")
        self._block_range = block_range
        self._front_layer = front_layer
        self._fidelity = float(-1)
        self._grad_scaler = grad_scaler
        self._hs = np.zeros(self._num_states, dtype=np.complex128)
        self._max_no = 0
        # Early start of the gradient sweep: scipy's L-BFGS-B asks for jac(theta) right after
        # fun(theta) (optimizer.py:585-590), so once that pattern has been seen twice objective()
        # enqueues the first gradient term (state 0, always needed, :150-165) before returning and
        # gradient() only collects it; an objective() call that finds an uncollected sweep (an
        # optimiser that does not ask for gradients) switches the early start off again.
        self._early_hits = 0
        self._early_on = False
        self._early_thetas = None
        self._sweep_key = None

    def objective(self, thetas: np.ndarray) -> float:
        self._store_latest_thetas(thetas)
        if self._early_thetas is not None:  # the previous early sweep was never collected
            self._early_thetas, self._early_on, self._early_hits = None, False, 0
        early = self._early_on and not self._dense
        # Leading state |s_0> (the rule near convergence): the WHOLE evaluation -- V^H sweep, hs gather and
        # the gradient sweep from |s_0> -- is enqueued at once; hs comes back as soon as the gather is
        # done while the gradient sweep keeps the GPU busy (aqc_sv_eval_begin).  If the hysteresis below
        # then picks another leader, that sweep is simply not collected.
        speculative = early and self._max_no == 0 and getattr(self._ws, "can_eval", True)
        if speculative:
            if self._target is None:
                raise RuntimeError("set_target() must be called before objective()")
            self._refresh_workspace()
            idx = self._state_handler.state_indices
            self._hs[:] = self._ws.eval_begin(thetas, SLOT_TARGET, SLOT_VH_TARGET, idx, x_basis=int(idx[0]),
                                              w=SLOT_W, z=SLOT_Z)[0]
            self._sweep_key = (0, self._weight)
        else:
            self._hs[:] = self._hs_products(thetas)
        np.copyto(self._hs2, np.abs(self._hs) ** 2)
        # hysteresis: the leader changes only if a state is better by 10% (:110-117); the scan runs only
        # when some state qualifies at all (it is sequential: the bar rises with every change)
        best = self._hs2[self._max_no]
        if (self._hs2 > 1.1 * best).any():
            for i in range(self._num_states):
                if 1.1 * best < self._hs2[i]:
                    best = self._hs2[i]
                    self._max_no = i
        w = self._weight
        self._fobj = float(1.0 - (1.0 - w) * self._hs2[0] - w * self._hs2[self._max_no])
        self._fidelity = float(self._hs2[0])
        if early:
            if not speculative or self._max_no != 0:
                self._begin_sweep(thetas)  # (drops a speculative sweep that assumed the wrong leader)
            self._early_thetas = self._last_thetas
        self._service.on_end_objective()
        return self._fobj

    def _begin_sweep(self, thetas: np.ndarray):
        """
        Enqueues the one gradient sweep an evaluation needs (basis flip states).  <V x|t> is
        antilinear in x, so the two weighted terms of :150-186,
            Re(-2(1-w) conj(hs_0) d<V s_0|t>) + Re(-2w conj(hs_max) d<V s_max|t>),
        are the real part of d<V x|t> for x = -2(1-w) hs_0 |s_0> - 2w hs_max |s_max>: one sweep of
        (w, z) instead of the reference's two.  With max_no == 0 the sweep starts from |s_0> and
        the factor -2 conj(hs_0) is applied afterwards, as in the reference.
        """
        idx = self._state_handler.state_indices
        if self._max_no == 0:
            self._ws.grad_begin(thetas, x_basis=int(idx[0]), z0=SLOT_VH_TARGET, w=SLOT_W, z=SLOT_Z)
        else:
            w, m = self._weight, self._max_no
            self._ws.set_sparse(SLOT_W, [int(idx[0]), int(idx[m])],
                                [-2.0 * (1.0 - w) * self._hs[0], -2.0 * w * self._hs[m]])
            self._ws.grad_begin(thetas, x_slot=SLOT_W, z0=SLOT_VH_TARGET, w=SLOT_W, z=SLOT_Z)
        self._sweep_key = (self._max_no, self._weight)

    def _sweep(self, thetas: np.ndarray, at_last_thetas: bool) -> np.ndarray:
        """
        Raw result of ``_begin_sweep``: collected from the early sweep if one is in flight.  An early
        sweep was started at ``_last_thetas``; ``at_last_thetas`` says that those are exactly ``thetas``.
        """
        early, self._early_thetas = self._early_thetas, None
        if (early is None or early is not self._last_thetas or not at_last_thetas
                or self._sweep_key != (self._max_no, self._weight)):
            # (a stale early sweep is dropped by the next workspace call)
            self._begin_sweep(thetas)
        return self._ws.grad_end()[0]

    def _dense_terms(self, thetas: np.ndarray) -> np.ndarray:
        """Generic (dense) flip states: the same single sweep, started from the uploaded combination."""
        hnd, w, m = self._state_handler, self._weight, self._max_no
        if m == 0:
            return -2.0 * np.conj(self._hs[0]) * self._raw_gradient(thetas, 0)
        x = (-2.0 * (1.0 - w) * self._hs[0]) * hnd.init_state(0) + (-2.0 * w * self._hs[m]) * hnd.init_state(m)
        self._ws.upload(SLOT_STATE, x)
        return self._ws.grad(thetas, x_slot=SLOT_STATE, z0=SLOT_VH_TARGET, w=SLOT_W, z=SLOT_Z)[0]

    def gradient(self, thetas: np.ndarray) -> np.ndarray:
        self._service.on_begin_gradient(self._fobj, thetas, self._fidelity)  # may raise: early stop
        last = self._last_thetas
        same = last.size == np.size(thetas) and np.array_equal(last, thetas)
        self._early_hits = self._early_hits + 1 if same else 0
        self._early_on = self._early_hits >= 2
        if not same or self._structure_stale():
            self._calc_objective_before_gradient(thetas)  # may recompute the objective (and restart the sweep)
            last = self._last_thetas
            same = last.size == np.size(thetas) and np.array_equal(last, thetas)  # False: within sqrt(eps) only
        circ = self._circuit
        front = bool(self._front_layer or self._block_range == (0, circ.num_blocks))

        if self._dense:
            raw = self._dense_terms(thetas)
        else:
            raw = self._sweep(thetas, same)
            if self._max_no == 0:
                raw = -2.0 * np.conj(self._hs[0]) * raw
        full = np.ascontiguousarray(np.real(mask_gradient(circ, raw, self._block_range, front)), dtype=np.float64)
        if self._grad_scaler:
            full *= self._grad_scaler.estimate(self._fobj)
        self._weight += self._gamma * (float(np.sqrt(abs(self._fobj))) - self._weight)
        self._service.on_end_gradient(self._fobj, self._fidelity, full, self._hs2, self._weight)
        return full

    @property
    def fidelity(self) -> float:
        return self._fidelity

    @property
    def max_no(self) -> int:
        return self._max_no

    @property
    def weight(self) -> float:
        return self._weight
