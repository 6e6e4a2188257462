"""
Drop-in of ``fast_dot_gradient`` (aqc_research/mps_dot_objective.py:41-242) on the GPU.
"""

from typing import Optional, Tuple
import numpy as np
from . import checking as chk
from .core_operations import mask_gradient
from .mps_engine import MpsWorkspace, QiskitMPS
from .mps_operations import no_truncation_threshold
from .parametric_circuit import ParametricCircuit, is_parametric_circuit


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
    assert is_parametric_circuit(circ)
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


# ------------------------------------------------------------------------------------------------
# Gate-by-gate helpers of the reference (mps_dot_objective.py:245-516): one qiskit-aer run per gate
# there.  Here every one of them is a tiny ansatz pushed through the GPU MPS engine:
#   * a 1-qubit gate is the front layer Rz(t0) Ry(t1) Rz(t2) of an ansatz whose two unit blocks
#     (same pair, zero angles) cancel: CX CX = 1.  Rx(a) = Rz(-pi/2) Ry(a) Rz(pi/2); the Pauli gates
#     are i R(pi), the factor i goes into the first site tensor;
#   * CX / CZ / CP(angle) is a single unit block with zero rotation angles, on any (ctrl, targ) pair
#     (non-adjacent pairs run through the engine's swap network, csrc/aqc_mps.cu build_mps_program).
# The hot path (fast_dot_gradient above) does not use them: it takes all derivatives of a pair-run
# from one reduced overlap matrix.
# ------------------------------------------------------------------------------------------------
from .mps_operations import mps_dot, v_mul_mps  # noqa: E402


def _one_qubit_gate(angles, qubit: int, mps_vec: QiskitMPS, phase: complex = 1.0) -> QiskitMPS:
    num_qubits = len(mps_vec[0])
    assert chk.is_int(qubit, 0 <= qubit < num_qubits)
    circ = ParametricCircuit(num_qubits, "cx", np.array([[0, 0], [1, 1]]))
    thetas = np.zeros(circ.num_thetas)
    thetas[3 * qubit : 3 * qubit + 3] = angles
    gam, lam = v_mul_mps(circ, thetas, mps_vec, trunc_thr=no_truncation_threshold())
    if phase != 1.0:
        gam = [tuple(phase * np.asarray(g) for g in gam[0])] + list(gam[1:])
    return (list(gam), list(lam))


def x_mul_mps(qubit: int, mps_vec: QiskitMPS) -> QiskitMPS:
    """``X_qubit @ mps_vec`` (:245-264)."""
    return _one_qubit_gate((-0.5 * np.pi, np.pi, 0.5 * np.pi), qubit, mps_vec, 1j)


def y_mul_mps(qubit: int, mps_vec: QiskitMPS) -> QiskitMPS:
    """``Y_qubit @ mps_vec`` (:267-286)."""
    return _one_qubit_gate((0.0, np.pi, 0.0), qubit, mps_vec, 1j)


def z_mul_mps(qubit: int, mps_vec: QiskitMPS) -> QiskitMPS:
    """``Z_qubit @ mps_vec`` (:289-308)."""
    return _one_qubit_gate((np.pi, 0.0, 0.0), qubit, mps_vec, 1j)


def rx_mul_mps(angle: float, qubit: int, mps_vec: QiskitMPS) -> QiskitMPS:
    """``Rx(angle)_qubit @ mps_vec`` (:311-331)."""
    assert chk.is_float(angle)
    return _one_qubit_gate((-0.5 * np.pi, float(angle), 0.5 * np.pi), qubit, mps_vec)


def ry_mul_mps(angle: float, qubit: int, mps_vec: QiskitMPS) -> QiskitMPS:
    """``Ry(angle)_qubit @ mps_vec`` (:334-354)."""
    assert chk.is_float(angle)
    return _one_qubit_gate((0.0, float(angle), 0.0), qubit, mps_vec)


def rz_mul_mps(angle: float, qubit: int, mps_vec: QiskitMPS) -> QiskitMPS:
    """``Rz(angle)_qubit @ mps_vec`` (:357-377)."""
    assert chk.is_float(angle)
    return _one_qubit_gate((float(angle), 0.0, 0.0), qubit, mps_vec)


def _entangler(kind: str, angle: float, ctrl: int, targ: int, mps_vec: QiskitMPS, trunc_thr: float) -> QiskitMPS:
    num_qubits = len(mps_vec[0])
    assert 0 <= ctrl < num_qubits and 0 <= targ < num_qubits and ctrl != targ
    circ = ParametricCircuit(num_qubits, kind, np.array([[ctrl], [targ]]))
    thetas = np.zeros(circ.num_thetas)
    if kind == "cp":
        thetas[3 * num_qubits + 4] = angle
    return v_mul_mps(circ, thetas, mps_vec, trunc_thr=trunc_thr)


def cx_mul_mps(_: float, ctrl: int, targ: int, mps_vec: QiskitMPS, *,
               trunc_thr: float = no_truncation_threshold()) -> QiskitMPS:
    """``CX(ctrl, targ) @ mps_vec`` (:380-407)."""
    return _entangler("cx", 0.0, ctrl, targ, mps_vec, trunc_thr)


def cp_mul_mps(angle: float, ctrl: int, targ: int, mps_vec: QiskitMPS, *,
               trunc_thr: float = no_truncation_threshold()) -> QiskitMPS:
    """``CPhase(angle)(ctrl, targ) @ mps_vec`` (:410-438)."""
    assert chk.is_float(angle)
    return _entangler("cp", float(angle), ctrl, targ, mps_vec, trunc_thr)


def cz_mul_mps(_: float, ctrl: int, targ: int, mps_vec: QiskitMPS, *,
               trunc_thr: float = no_truncation_threshold()) -> QiskitMPS:
    """``CZ(ctrl, targ) @ mps_vec`` (:441-468)."""
    return _entangler("cz", 0.0, ctrl, targ, mps_vec, trunc_thr)


def dot_x(qubit: int, w_vec: QiskitMPS, z_vec: QiskitMPS) -> np.complex128:
    """``0.5j <X w|z>`` (:471-484)."""
    return np.complex128(0.5j * mps_dot(x_mul_mps(qubit, w_vec), z_vec))


def dot_y(qubit: int, w_vec: QiskitMPS, z_vec: QiskitMPS) -> np.complex128:
    """``0.5j <Y w|z>`` (:487-500)."""
    return np.complex128(0.5j * mps_dot(y_mul_mps(qubit, w_vec), z_vec))


def dot_z(qubit: int, w_vec: QiskitMPS, z_vec: QiskitMPS) -> np.complex128:
    """``0.5j <Z w|z>`` (:503-516)."""
    return np.complex128(0.5j * mps_dot(z_mul_mps(qubit, w_vec), z_vec))
