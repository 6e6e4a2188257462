"""
2x2 gate matrices and small dense unit-block matrices (host side, NumPy).
Reference: aqc_research/elementary_operations.py:39-291 -- same names and conventions:

    Rx(phi) = [[c, -i s], [-i s, c]],  Ry(phi) = [[c, -s], [s, c]],  Rz(phi) = diag(e^{-i phi/2}, e^{i phi/2}),
    P(phi)  = diag(1, e^{i phi}),      c = cos(phi/2), s = sin(phi/2)

Dense n-qubit matrices put qubit 0 on the LEFT of the Kronecker product (most significant index bit),
which is why the vector kernels act on bit ``n - 1 - pos`` (core_operations.bit2bit_transform).
These are the oracles the reference's tests compare the fast kernels with; nothing here is on the
hot path (the device computes cos/sin itself, csrc/aqc_sv.cu trig_kernel).
"""

from typing import Dict
import numpy as np

_C128 = np.complex128


def _fill(out: np.ndarray, a, b, c, d) -> np.ndarray:
    assert isinstance(out, np.ndarray) and out.shape == (2, 2) and out.dtype == _C128
    out[0, 0], out[0, 1], out[1, 0], out[1, 1] = a, b, c, d
    return out


def make_rx(phi: float, out: np.ndarray) -> np.ndarray:
    """Rx gate written into ``out`` (elementary_operations.py:143-165)."""
    c, s = np.cos(0.5 * phi), np.sin(0.5 * phi)
    return _fill(out, c, -1j * s, -1j * s, c)


def make_ry(phi: float, out: np.ndarray) -> np.ndarray:
    """Ry gate written into ``out`` (:188-210)."""
    c, s = np.cos(0.5 * phi), np.sin(0.5 * phi)
    return _fill(out, c, -s, s, c)


def make_rz(phi: float, out: np.ndarray) -> np.ndarray:
    """Rz gate written into ``out`` (:230-251)."""
    return _fill(out, np.exp(-0.5j * phi), 0.0, 0.0, np.exp(0.5j * phi))


def np_rx(phi: float) -> np.ndarray:
    return make_rx(phi, np.empty((2, 2), dtype=_C128))


def np_ry(phi: float) -> np.ndarray:
    return make_ry(phi, np.empty((2, 2), dtype=_C128))


def np_rz(phi: float) -> np.ndarray:
    return make_rz(phi, np.empty((2, 2), dtype=_C128))


def np_phase(phi: float) -> np.ndarray:
    """Phase gate diag(1, e^{i phi}) (:254-266)."""
    return _fill(np.empty((2, 2), dtype=_C128), 1.0, 0.0, 0.0, np.exp(1j * phi))


def np_x() -> np.ndarray:
    return np.array([[0, 1], [1, 0]], dtype=_C128)


def np_y() -> np.ndarray:
    return np.array([[0, -1j], [1j, 0]], dtype=_C128)


def np_z() -> np.ndarray:
    return np.array([[1, 0], [0, -1]], dtype=_C128)


def _embed(n: int, ops: Dict[int, np.ndarray]) -> np.ndarray:
    """Kronecker product over qubits 0..n-1 (qubit 0 leftmost) with identities where ``ops`` is silent."""
    mat = np.ones((1, 1), dtype=_C128)
    for q in range(n):
        mat = np.kron(mat, ops.get(q, np.eye(2, dtype=_C128)))
    return mat


def np_block_matrix(n: int, c: int, t: int, c_mat: np.ndarray, t_mat: np.ndarray, gate_mat: np.ndarray) -> np.ndarray:
    """
    Dense matrix of a unit block, (c_mat (x) t_mat) . controlled-G, on qubits ``c`` (control) and
    ``t`` (target) of ``n`` (:39-81):  |0><0|_c c-part (x) t_mat  +  |1><1|_c c-part (x) t_mat G.
    """
    assert 0 <= c < n and 0 <= t < n and c != t
    p0 = np.array([[1, 0], [0, 0]], dtype=_C128)
    p1 = np.array([[0, 0], [0, 1]], dtype=_C128)
    return _embed(n, {c: c_mat @ p0, t: np.asarray(t_mat, dtype=_C128)}) + _embed(n, {c: c_mat @ p1, t: t_mat @ gate_mat})


def np_cx_matrix(n: int, c: int, t: int) -> np.ndarray:
    """Dense CX with control ``c`` and target ``t`` (:84-120)."""
    eye = np.eye(2, dtype=_C128)
    return np_block_matrix(n, c, t, eye, eye, np_x())
