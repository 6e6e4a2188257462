"""
Function-level drop-ins for the reference's state-vector numeric core
(reference: aqc_research/core_operations.py:606-1019).  Same names, argument meaning and
error behaviour; host arrays in/out; the arithmetic runs on the GPU through
``libaqc_b200.so``.  There is no CPU fallback.

For repeated evaluations use the objective classes (``model_sp_lhs``) which keep the target and
work vectors resident in HBM; these shims copy their vector arguments over PCIe on each call.
"""

from collections import OrderedDict
from typing import Optional, Tuple
import numpy as np
from . import checking as chk
from .engine import SvWorkspace
from .parametric_circuit import ParametricCircuit

_CACHE_SIZE = 4
_workspaces: "OrderedDict[tuple, SvWorkspace]" = OrderedDict()


def _workspace(circ: ParametricCircuit, log2_cols: int = 0, device: int = 0) -> SvWorkspace:
    """Small LRU of GPU workspaces keyed by the circuit structure."""
    key = (
        circ.num_qubits,
        circ.entangler,
        type(circ).__name__,
        bool(getattr(circ, "is_second_order", False)),
        np.ascontiguousarray(circ.blocks, dtype=np.int32).tobytes(),
        log2_cols,
        device,
    )
    ws = _workspaces.get(key)
    if ws is None:
        ws = SvWorkspace(circ, num_slots=3, device=device, log2_cols=log2_cols)
        _workspaces[key] = ws
        while len(_workspaces) > _CACHE_SIZE:
            _, old = _workspaces.popitem(last=False)
            old.close()
    else:
        _workspaces.move_to_end(key)
    return ws


def clear_workspace_cache():
    """Frees the cached GPU workspaces."""
    while _workspaces:
        _, ws = _workspaces.popitem()
        ws.close()


def _check_vectors(circ, thetas, *vecs):
    assert isinstance(circ, ParametricCircuit)
    assert chk.float_1d(thetas, thetas.size == circ.num_thetas)
    for v in vecs:
        assert chk.complex_1d(v, v.size == circ.dimension) and v.flags.c_contiguous


def v_mul_vec(
    circ: ParametricCircuit,
    thetas: np.ndarray,
    vec: np.ndarray,
    out: np.ndarray,
    workspace: Optional[np.ndarray] = None,
) -> np.ndarray:
    """``out = V(thetas) @ vec`` (core_operations.py:606-710). ``out`` may alias ``vec``."""
    _check_vectors(circ, thetas, vec, out)
    ws = _workspace(circ)
    ws.upload(0, vec)
    ws.apply(thetas, 0, 0, dagger=False)
    return ws.download(0, 0, out)


def v_dagger_mul_vec(
    circ: ParametricCircuit,
    thetas: np.ndarray,
    vec: np.ndarray,
    out: np.ndarray,
    workspace: Optional[np.ndarray] = None,
) -> np.ndarray:
    """``out = V(thetas)^H @ vec`` (core_operations.py:713-820)."""
    _check_vectors(circ, thetas, vec, out)
    ws = _workspace(circ)
    ws.upload(0, vec)
    ws.apply(thetas, 0, 0, dagger=True)
    return ws.download(0, 0, out)


def grad_of_dot_product(
    circ: ParametricCircuit,
    thetas: np.ndarray,
    x_vec: np.ndarray,
    vh_y_vec: np.ndarray,
    workspace: Optional[np.ndarray] = None,
    block_range: Optional[Tuple[int, int]] = None,
    front_layer: bool = True,
) -> np.ndarray:
    """
    Complex gradient of ``<V x, y>`` given ``vh_y_vec = V^H y`` (core_operations.py:823-1019).
    Entries of blocks outside ``block_range`` (and of the front layer if disabled) are zero.
    The inputs are not modified.
    """
    _check_vectors(circ, thetas, x_vec, vh_y_vec)
    assert chk.is_bool(front_layer)
    block_range = (0, circ.num_blocks) if block_range is None else block_range
    assert chk.is_tuple(block_range, len(block_range) == 2)
    assert 0 <= block_range[0] < block_range[1] <= circ.num_blocks
    ws = _workspace(circ)
    ws.upload(0, x_vec)
    ws.upload(1, vh_y_vec)
    grad = ws.grad(thetas, x_slot=0, z0=1, w=0, z=1)[0]
    return mask_gradient(circ, grad, block_range, front_layer)


def mask_gradient(circ, grad, block_range, front_layer) -> np.ndarray:
    """Zeroes the entries the reference does not record (core_operations.py:936-949,996-1013)."""
    g2 = circ.subset2q(grad)
    g2[: block_range[0]] = 0
    g2[block_range[1] :] = 0
    if not front_layer:
        circ.subset1q(grad)[:] = 0
    return grad
