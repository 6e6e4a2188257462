"""
Coordinate-descent optimisation of the full AQC objective ``fobj = 1 - |<V(thetas), U>|^2 / dim^2``
(reference: aqc_research/model_sketching/aqc_coord_descent.py:32-124 ``_single_simulation`` around
core_op_matrix.py:765-917 ``coord_descent_single_sweep``).

The reference runs one start per *process* (job_executor.py:141); here the starts of a multistart
run are the batch dimension of one GPU workspace: every sweep is one V^H-apply launch sequence plus
ONE kernel (one CTA per start) that updates all angles of all starts on the device.
"""

from typing import Optional
import numpy as np
from ..engine import SvWorkspace
from ..parametric_circuit import ParametricCircuit, is_parametric_circuit

SLOT_TARGET, SLOT_W, SLOT_Z = 0, 1, 2


class BatchedCoordinateDescent:
    """``batch`` independent coordinate-descent runs against one target unitary."""

    def __init__(self, circ: ParametricCircuit, target: np.ndarray, batch: int = 1, device: int = 0):
        assert is_parametric_circuit(circ)
        if circ.entangler == "cp":
            raise NotImplementedError("CPhase entangler is not supported yet")
        target = np.ascontiguousarray(target, dtype=np.complex128)
        assert target.shape == (circ.dimension, circ.dimension)
        self.circ = circ
        self.batch = int(batch)
        self.workspace = SvWorkspace(circ, 3, device=device, log2_cols=circ.num_qubits, batch=batch)
        self.workspace.upload(SLOT_TARGET, target)  # broadcast to every batch element

    @property
    def num_thetas(self) -> int:
        return self.circ.num_thetas

    def sweep(self, thetas: np.ndarray, num_sweeps: int = 1):
        """Returns (fobj[num_sweeps, batch], thetas[batch, T]) after ``num_sweeps`` sweeps."""
        return self.workspace.coord_descent(
            thetas, target=SLOT_TARGET, w=SLOT_W, z=SLOT_Z, num_sweeps=num_sweeps
        )

    def run(
        self,
        thetas_0: np.ndarray,
        maxiter: int,
        *,
        thetas_change_threshold: float = 1e-8,
        fobj_thr: Optional[float] = 1e-2,
    ) -> dict:
        """
        The loop of the reference's ``_single_simulation`` (aqc_coord_descent.py:66-103) for every
        start: sweep until the largest angle change drops below ``thetas_change_threshold``, the
        objective drops below ``fobj_thr`` (SmallObjectiveStopper) or ``maxiter`` sweeps are done;
        the best objective and its angles are kept per start.
        """
        th = np.array(thetas_0, dtype=np.float64).reshape(self.batch, self.num_thetas).copy()
        best_f = np.full(self.batch, np.inf)
        best_th = th.copy()
        active = np.ones(self.batch, dtype=bool)
        nit = np.zeros(self.batch, dtype=np.int64)
        profile = []
        for _ in range(int(maxiter)):
            fobj, new_th = self.sweep(th)
            f = fobj[0]
            change = np.max(np.abs(new_th - th), axis=1)
            upd = active & (f < best_f)
            best_f[upd] = f[upd]
            best_th[upd] = new_th[upd]
            nit[active] += 1
            profile.append(np.where(active, f, np.nan))
            th[active] = new_th[active]  # finished starts stay frozen
            done = change < thetas_change_threshold
            if fobj_thr is not None:
                done |= f < fobj_thr
            active &= ~done
            if not active.any():
                break
        return {"cost": best_f, "thetas": best_th, "nit": nit, "convergence_profile": np.array(profile)}

    def close(self):
        self.workspace.close()
