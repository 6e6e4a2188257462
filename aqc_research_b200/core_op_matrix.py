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
from .parametric_circuit import ParametricCircuit, is_parametric_circuit


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
    assert is_parametric_circuit(circ)
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
    assert is_parametric_circuit(circ)
    assert chk.float_1d(thetas, thetas.size == circ.num_thetas)
    assert chk.complex_2d(target, target.shape[0] == target.shape[1] == circ.dimension)
    if circ.entangler == "cp":
        raise NotImplementedError("CPhase entangler is not supported yet")
    ws = _workspace(circ, log2_cols=circ.num_qubits)
    ws.upload(0, np.ascontiguousarray(target))
    fobj, new_thetas = ws.coord_descent(thetas, target=0, w=1, z=2, num_sweeps=1)
    thetas[:] = new_thetas[0]
    return float(fobj[0, 0])


# ------------------------------------------------------------------------------------------------
# Gate-by-gate primitives on matrices (core_op_matrix.py:32-477 of the reference).  A gate acts on
# bit ``qubit_no`` of the ROW index of the row-major (2^n, m) matrix, i.e. on the flat-index stride
# m * 2^qubit_no, for any column count m <= 2^n.  In place on the host array, computed on the GPU
# (csrc/aqc_prim.cu); the workspace arguments are checked as in the reference but not needed.
# ------------------------------------------------------------------------------------------------
from . import _prim  # noqa: E402
from . import elementary_operations as _eo  # noqa: E402

_P11 = np.array([[0, 0], [0, 1]], dtype=np.complex128)


def _check_mat(qubits, workspace, *mats) -> int:
    m0 = mats[0]
    for m in mats:
        assert chk.complex_2d(m, m.shape[1] <= m.shape[0]) and m.flags.c_contiguous and m.shape == m0.shape
    n = int(round(np.log2(m0.shape[0])))
    assert 2**n == m0.shape[0]
    for q in qubits:
        assert chk.is_int(q, 0 <= q < n)
    if workspace is not None:
        assert chk.complex_array(workspace, workspace.size >= m0.size)
        assert not np.may_share_memory(m0, workspace)
    return n


def _mstride(mat: np.ndarray, qubit_no: int) -> int:
    return mat.shape[1] << qubit_no


def gate2x2_mul_mat(qubit_no: int, gate2x2: np.ndarray, mat: np.ndarray, workspace: np.ndarray) -> np.ndarray:
    """``mat <- gate @ mat`` with the 2x2 gate on row-index bit ``qubit_no`` (core_op_matrix.py:392-427)."""
    assert chk.complex_or_float_2d(gate2x2, gate2x2.shape == (2, 2))
    _check_mat([qubit_no], workspace, mat)
    return _prim.apply_gates(mat, [(_mstride(mat, qubit_no), 0, 0, gate2x2)])


def _rot_mul_mat(make, angle, qubit_no, mat, workspace):
    assert chk.is_float(angle)
    _check_mat([qubit_no], workspace, mat)
    return _prim.apply_gates(mat, [(_mstride(mat, qubit_no), 0, 0, make(float(angle)))])


def rx_mul_mat(angle: float, qubit_no: int, mat: np.ndarray, workspace: np.ndarray) -> np.ndarray:
    """Rx(angle) on ``qubit_no``, in place (:32-63)."""
    return _rot_mul_mat(_eo.np_rx, angle, qubit_no, mat, workspace)


def ry_mul_mat(angle: float, qubit_no: int, mat: np.ndarray, workspace: np.ndarray) -> np.ndarray:
    """Ry(angle) on ``qubit_no``, in place (:66-97)."""
    return _rot_mul_mat(_eo.np_ry, angle, qubit_no, mat, workspace)


def rz_mul_mat(angle: float, qubit_no: int, mat: np.ndarray, ___: Optional[np.ndarray] = None) -> np.ndarray:
    """Rz(angle) on ``qubit_no``, in place (:100-127)."""
    return _rot_mul_mat(_eo.np_rz, angle, qubit_no, mat, None)


def _ctrl_mul_mat(gate, ctrl, targ, mat, workspace):
    assert ctrl != targ
    _check_mat([ctrl, targ], workspace, mat)
    return _prim.apply_gates(mat, [(_mstride(mat, targ), _mstride(mat, ctrl), 1, gate)])


def cx_mul_mat(ctrl: int, targ: int, ___: float, mat: np.ndarray, workspace: np.ndarray) -> np.ndarray:
    """CX, in place (:130-178)."""
    return _ctrl_mul_mat(_eo.np_x(), ctrl, targ, mat, workspace)


def cz_mul_mat(ctrl: int, targ: int, ___: float, mat: np.ndarray, workspace: np.ndarray) -> np.ndarray:
    """CZ, in place (:181-229)."""
    return _ctrl_mul_mat(_eo.np_z(), ctrl, targ, mat, workspace)


def cp_mul_mat(ctrl: int, targ: int, angle: float, mat: np.ndarray, workspace: np.ndarray) -> np.ndarray:
    """CPhase(angle), in place (:232-281)."""
    assert chk.is_float(angle)
    return _ctrl_mul_mat(_eo.np_phase(float(angle)), ctrl, targ, mat, workspace)


def _pauli_dot_mat(pauli, qubit_no, w_mat, z_mat, workspace) -> np.complex128:
    _check_mat([qubit_no], workspace, w_mat, z_mat)
    return np.complex128(0.5j * _prim.gate_vdot(w_mat, z_mat, (_mstride(w_mat, qubit_no), 0, 0, pauli)))


def x_dot_mat(qubit_no: int, w_mat: np.ndarray, z_mat: np.ndarray, workspace: np.ndarray) -> np.complex128:
    """``0.5j <X w|z>`` (Frobenius, :284-317)."""
    return _pauli_dot_mat(_eo.np_x(), qubit_no, w_mat, z_mat, workspace)


def y_dot_mat(qubit_no: int, w_mat: np.ndarray, z_mat: np.ndarray, workspace: np.ndarray) -> np.complex128:
    """``0.5j <Y w|z>`` (:320-353)."""
    return _pauli_dot_mat(_eo.np_y(), qubit_no, w_mat, z_mat, workspace)


def z_dot_mat(qubit_no: int, w_mat: np.ndarray, z_mat: np.ndarray, workspace: np.ndarray) -> np.complex128:
    """``0.5j <Z w|z>`` (:356-389)."""
    return _pauli_dot_mat(_eo.np_z(), qubit_no, w_mat, z_mat, workspace)


def derv_cphase(ctrl: int, targ: int, w_mat: np.ndarray, z_mat: np.ndarray, workspace: np.ndarray) -> np.complex128:
    """Derivative of ``<w|z>`` by the CPhase angle: ``-1j <(|1><1|_c (x) |1><1|_t) w|z>`` (:430-477)."""
    assert ctrl != targ
    _check_mat([ctrl, targ], workspace, w_mat, z_mat)
    op = (_mstride(w_mat, targ), _mstride(w_mat, ctrl), 2, _P11)
    return np.complex128(-1j * _prim.gate_vdot(w_mat, z_mat, op))
