"""
Dense matrix of an ansatz (reference: aqc_research/circuit_transform.py:249-390), computed on the GPU.

``ansatz_to_numpy_fast``    -- V applied to the identity by the matrix path (generic ansatz; the
                               reference does exactly this with its NumPy ``v_mul_mat``, :273-287);
``ansatz_to_numpy_trotter`` -- column k is ``v_mul_vec(e_k)``, which also covers ``TrotterAnsatz``
                               (the reference builds the same matrix from Kronecker products and its
                               tests assert the equality, test_core_operations.py:283-321).
Both are O(4^n) testing aids.  The Qiskit converters of the reference (``ansatz_to_qcircuit``,
``qcircuit_to_matrix`` ...) need Qiskit, which is not part of this package.
"""

import numpy as np
from . import checking as chk
from .core_op_matrix import v_mul_mat
from .core_operations import v_mul_vec
from .parametric_circuit import ParametricCircuit, TrotterAnsatz, is_parametric_circuit, is_trotter_ansatz


def ansatz_to_numpy_fast(circ: ParametricCircuit, thetas: np.ndarray) -> np.ndarray:
    """Circuit matrix of a generic (non-Trotter) ansatz (:273-287)."""
    assert is_parametric_circuit(circ) and chk.float_1d(thetas)
    if is_trotter_ansatz(circ):
        raise ValueError("ansatz_to_numpy_fast does not support TrotterAnsatz; use ansatz_to_numpy_trotter")
    mat = np.eye(circ.dimension, dtype=np.complex128)
    return v_mul_mat(circ, thetas, mat, workspace=None)


def ansatz_to_numpy_trotter(circ: ParametricCircuit, thetas: np.ndarray) -> np.ndarray:
    """Circuit matrix of any ansatz, Trotterized ones included (:290-390)."""
    assert is_parametric_circuit(circ) and chk.float_1d(thetas)
    dim = circ.dimension
    mat = np.empty((dim, dim), dtype=np.complex128)
    col, out = np.zeros(dim, dtype=np.complex128), np.empty(dim, dtype=np.complex128)
    for k in range(dim):
        col[:] = 0
        col[k] = 1
        mat[:, k] = v_mul_vec(circ, thetas, col, out)
    return mat


def _needs_qiskit(name: str):
    def _fn(*_, **__):
        raise NotImplementedError(f"{name} converts to / from Qiskit objects; Qiskit is not part of this package")

    _fn.__name__ = name
    _fn.__doc__ = "Qiskit converter of the reference (circuit_transform.py); unavailable without Qiskit."
    return _fn


qcircuit_to_state = _needs_qiskit("qcircuit_to_state")
qcircuit_to_matrix = _needs_qiskit("qcircuit_to_matrix")
state_preparation_qcircuit = _needs_qiskit("state_preparation_qcircuit")
ansatz_to_qcircuit = _needs_qiskit("ansatz_to_qcircuit")
ansatz_to_numpy_by_qiskit = _needs_qiskit("ansatz_to_numpy_by_qiskit")
