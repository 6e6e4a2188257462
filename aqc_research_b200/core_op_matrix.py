"""
Function-level drop-ins for the reference's matrix numeric core
(reference: aqc_research/core_op_matrix.py:480-762): the same sweeps applied to the columns of
a row-major ``(2^n, m)`` matrix.  On the GPU the matrix is a state of ``n + ceil(log2 m)`` index
bits whose low bits are the column index (columns are zero-padded to a power of two; zero
columns contribute nothing to the Frobenius inner products).
"""

from typing import Optional
import numpy as np
from . import checking as chk
from .core_operations import _workspace
from .parametric_circuit import ParametricCircuit


def _pad_cols(mat: np.ndarray, log2_cols: int) -> np.ndarray:
    cols = 1 << log2_cols
    if mat.shape[1] == cols:
        return np.ascontiguousarray(mat)
    out = np.zeros((mat.shape[0], cols), dtype=np.complex128)
    out[:, : mat.shape[1]] = mat
    return out


def _log2_cols(m: int) -> int:
    return max(0, int(m - 1).bit_length())


def _check(circ, thetas, *mats):
    assert isinstance(circ, ParametricCircuit)
    assert chk.float_1d(thetas, thetas.size == circ.num_thetas)
    for m in mats:
        assert chk.complex_2d(m, m.shape[0] == circ.dimension and 1 <= m.shape[1] <= m.shape[0])
        assert m.flags.c_contiguous
        assert m.shape == mats[0].shape


def _apply(circ, thetas, mat, dagger):
    k = _log2_cols(mat.shape[1])
    ws = _workspace(circ, log2_cols=k)
    ws.upload(0, _pad_cols(mat, k))
    ws.apply(thetas, 0, 0, dagger=dagger)
    res = ws.download(0, 0).reshape(mat.shape[0], 1 << k)
    mat[:, :] = res[:, : mat.shape[1]]
    return mat


def v_mul_mat(circ, thetas: np.ndarray, mat: np.ndarray, workspace: Optional[np.ndarray] = None):
    """``mat <- V @ mat`` in place (core_op_matrix.py:480-559)."""
    _check(circ, thetas, mat)
    return _apply(circ, thetas, mat, False)


def v_dagger_mul_mat(circ, thetas: np.ndarray, mat: np.ndarray, workspace: Optional[np.ndarray] = None):
    """``mat <- V^H @ mat`` in place (core_op_matrix.py:562-642)."""
    _check(circ, thetas, mat)
    return _apply(circ, thetas, mat, True)


def grad_of_matrix_dot_product(
    circ,
    thetas: np.ndarray,
    x_mat: np.ndarray,
    vh_y_mat: np.ndarray,
    workspace: Optional[np.ndarray] = None,
) -> np.ndarray:
    """
    Complex gradient of ``<V X, Y>_F`` given ``vh_y_mat = V^H Y`` (core_op_matrix.py:645-762).
    The reference overwrites ``x_mat`` and ``vh_y_mat`` with its swept work matrices; callers
    cannot rely on their contents afterwards.  Here the sweep runs on device copies (which hold
    rescaled work states) and the host arrays are left untouched.
    """
    _check(circ, thetas, x_mat, vh_y_mat)
    k = _log2_cols(x_mat.shape[1])
    ws = _workspace(circ, log2_cols=k)
    ws.upload(0, _pad_cols(x_mat, k))
    ws.upload(1, _pad_cols(vh_y_mat, k))
    return ws.grad(thetas, x_slot=0, z0=1, w=0, z=1)[0]


def coord_descent_single_sweep(
    circ, thetas: np.ndarray, target: np.ndarray, workspace: Optional[np.ndarray] = None
) -> float:
    """
    One coordinate-descent sweep over all angles for ``fobj = 1 - |<V, U>|^2 / dim^2``
    (core_op_matrix.py:765-917): ``thetas`` is modified IN PLACE, the objective at the end of the
    sweep is returned.  cx / cz entanglers only, as in the reference (:818-819).
    """
    assert isinstance(circ, ParametricCircuit)
    assert chk.float_1d(thetas, thetas.size == circ.num_thetas)
    assert chk.complex_2d(target, target.shape[0] == target.shape[1] == circ.dimension)
    if circ.entangler == "cp":
        raise NotImplementedError("CPhase entangler is not supported yet")
    ws = _workspace(circ, log2_cols=circ.num_qubits)
    ws.upload(0, np.ascontiguousarray(target))
    fobj, new_thetas = ws.coord_descent(thetas, target=0, w=1, z=2, num_sweeps=1)
    thetas[:] = new_thetas[0]
    return float(fobj[0, 0])
