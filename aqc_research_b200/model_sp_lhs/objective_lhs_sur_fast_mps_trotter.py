"""
Surrogate state-preparation objective on MPS states, GPU edition.
Reference: aqc_research/model_sp_lhs/objective_lhs_sur_fast_mps_trotter.py:57-232 (the default
``UserOptions.objective``, user_options.py:91).  Same surrogate as ``SpSurrogateObjectiveMax``;
the target, V^H target and the sweep states are device-resident MPS in an ``MpsWorkspace``.

``state_prep_func`` must describe an X-type preparation (basis state): None, an int basis index
or a list of flipped qubits (the reference's default is the Neel state, trotter.py:389-398).
"""

from typing import Optional, Tuple
import numpy as np
from .. import checking as chk
from ..core_operations import mask_gradient
from ..mps_engine import MpsWorkspace
from ..mps_operations import check_mps
from ..parametric_circuit import TrotterAnsatz, first_layer_included, is_trotter_ansatz, layer_to_block_range
from .objective_base import BasisStateHandler, SpLHSObjectiveBase, make_state_handler

_SLOT_TARGET, _SLOT_VH, _SLOT_W, _SLOT_Z = 0, 1, 2, 3


class SpSurrogateObjectiveFastMpsTrotter(SpLHSObjectiveBase):
    """Drop-in for the reference class of the same name."""

    _gamma = 0.1

    def __init__(
        self,
        *,
        user_parameters: dict,
        circ: TrotterAnsatz,
        layer_range: Optional[Tuple[int, int]] = None,
        alt_layers: bool = False,
        verbose: bool = False,
        grad_scaler=None,
    ):
        if not is_trotter_ansatz(circ):
            raise ValueError("expects Trotterized ansatz")
        super().__init__(user_parameters, circ, use_mps=True, verbose=verbose)
        if user_parameters["max_flips"] != 1:
            raise ValueError("expects max_flips=1 in case of using MPS")
        handler = make_state_handler(circ.num_qubits, 1, user_parameters.get("state_prep_func", None))
        if not isinstance(handler, BasisStateHandler):
            raise ValueError("the MPS objective supports basis-state (X-type) preparations only")
        self._state_handler = handler
        self._num_states = handler.num_states
        self._init_common()
        assert layer_range is None or chk.is_tuple(layer_range, len(layer_range) == 2)
        self._layer_range = layer_range
        self._alt_layers = bool(alt_layers)
        self._trunc_thr = float(user_parameters["trunc_thr"])
        self._chi_max = int(user_parameters.get("chi_max", 64))
        self._fidelity = float(-1)
        self._grad_scaler = grad_scaler
        self._hs = np.zeros(self._num_states, dtype=np.complex128)
        self._max_no = 0
        self._mps = MpsWorkspace(circ, num_slots=4, chi_max=self._chi_max, trunc_thr=self._trunc_thr,
                                 device=self._device)

    def _refresh_mps_workspace(self):
        """Re-creates the MPS workspace after a structure change (insert_unit_blocks / update_structure)."""
        if self._structure_changed():
            self._mps.close()
            self._mps = MpsWorkspace(self._circuit, num_slots=4, chi_max=self._chi_max, trunc_thr=self._trunc_thr,
                                     device=self._device)
            if self._target is not None:
                self._mps.upload(_SLOT_TARGET, self._target)

    def set_target(self, target) -> None:
        assert check_mps(target) and len(target[0]) == self._circuit.num_qubits
        self._target = target
        self._refresh_mps_workspace()
        self._mps.upload(_SLOT_TARGET, target)
        self._last_thetas = np.empty(0)

    def objective(self, thetas: np.ndarray) -> float:
        if self._target is None:
            raise RuntimeError("set_target() must be called before objective()")
        self._refresh_mps_workspace()
        self._store_latest_thetas(thetas)
        self._hs[:] = self._mps.objective(thetas, _SLOT_TARGET, _SLOT_VH, self._state_handler.state_indices)
        np.copyto(self._hs2, np.abs(self._hs) ** 2)
        best = self._hs2[self._max_no]
        for i in range(self._num_states):
            if 1.1 * best < self._hs2[i]:
                best = self._hs2[i]
                self._max_no = i
        w = self._weight
        self._fobj = float(1.0 - (1.0 - w) * self._hs2[0] - w * self._hs2[self._max_no])
        self._fidelity = float(self._hs2[0])
        self._service.on_end_objective()
        return self._fobj

    def _raw(self, thetas, state_no, block_range, front):
        idx = int(self._state_handler.state_indices[state_no])
        g = self._mps.grad(thetas, x_basis=idx, z0=_SLOT_VH, w=_SLOT_W, z=_SLOT_Z)
        return mask_gradient(self._circuit, g, block_range, front)

    def _raw_pair(self, thetas, block_range, front):
        """
        Both weighted terms of the reference (:190-227) from ONE sweep: <V x|t> is antilinear in x
        and  -2(1-w) hs_0 |s_0> - 2w hs_max |s_max>  is a product state again (the two basis states
        differ in one qubit), so the sweep starts from that bond-1 MPS, normalised, and the norm
        is multiplied back.
        """
        idx, w, m = self._state_handler.state_indices, self._weight, self._max_no
        i0, flip = int(idx[0]), int(idx[0]) ^ int(idx[m])
        if flip == 0 or flip & (flip - 1):  # not a single flip: two sweeps, as the reference
            g0 = self._raw(thetas, 0, block_range, front)
            gm = self._raw(thetas, m, block_range, front)
            return -2.0 * (1.0 - w) * np.conj(self._hs[0]) * g0 - 2.0 * w * np.conj(self._hs[m]) * gm
        site = flip.bit_length() - 1
        a0, am = -2.0 * (1.0 - w) * self._hs[0], -2.0 * w * self._hs[m]
        scale = float(np.hypot(abs(a0), abs(am)))
        if scale == 0.0:
            return np.zeros(self._circuit.num_thetas, dtype=np.complex128)
        amps = (a0 / scale, am / scale) if (i0 >> site) & 1 == 0 else (am / scale, a0 / scale)
        self._mps.set_product_site(_SLOT_W, i0, site, amps[0], amps[1])
        g = self._mps.grad(thetas, x_slot=_SLOT_W, z0=_SLOT_VH, w=_SLOT_W, z=_SLOT_Z)
        return scale * mask_gradient(self._circuit, g, block_range, front)

    def gradient(self, thetas: np.ndarray) -> np.ndarray:
        self._service.on_begin_gradient(self._fobj, thetas, self._fidelity)
        self._refresh_mps_workspace()
        self._calc_objective_before_gradient(thetas)
        circ = self._circuit
        block_range = layer_to_block_range(circ, self._layer_range)
        front = first_layer_included(circ, self._layer_range)
        if self._max_no == 0:
            full = np.real(-2.0 * np.conj(self._hs[0]) * self._raw(thetas, 0, block_range, front))
        else:
            full = np.real(self._raw_pair(thetas, block_range, front))
        full = np.ascontiguousarray(full, dtype=np.float64)
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

    @property
    def mps_workspace(self) -> MpsWorkspace:
        return self._mps
