"""
Drop-in of ``fast_dot_gradient`` (aqc_research/mps_dot_objective.py:41-242) on the GPU.
"""

from typing import Optional, Tuple
import numpy as np
from . import checking as chk
from .core_operations import mask_gradient
from .mps_engine import MpsWorkspace, QiskitMPS
from .mps_operations import no_truncation_threshold
from .parametric_circuit import ParametricCircuit


def fast_dot_gradient(
    circ: ParametricCircuit,
    thetas: np.ndarray,
    lvec: QiskitMPS,
    vh_phi: QiskitMPS,
    *,
    trunc_thr: Optional[float] = no_truncation_threshold(),
    block_range: Optional[Tuple[int, int]] = None,
    front_layer: Optional[bool] = True,
) -> np.ndarray:
    """
    Complex gradient of ``<lvec|V^H|phi>`` given ``vh_phi = V^H|phi>``; entries outside
    ``block_range`` / of a disabled front layer are zero, as in the reference.
    """
    assert isinstance(circ, ParametricCircuit)
    assert chk.float_1d(thetas, thetas.size == circ.num_thetas)
    assert isinstance(lvec, tuple) and isinstance(vh_phi, tuple)
    block_range = (0, circ.num_blocks) if block_range is None else block_range
    assert chk.is_tuple(block_range, len(block_range) == 2)
    assert 0 <= block_range[0] < block_range[1] <= circ.num_blocks
    ws = MpsWorkspace(circ, num_slots=4, chi_max=64, trunc_thr=float(trunc_thr))
    ws.upload(0, lvec)
    ws.upload(1, vh_phi)
    grad = ws.grad(thetas, x_slot=0, z0=1, w=2, z=3)
    ws.close()
    return mask_gradient(circ, grad, block_range, bool(front_layer))
