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
from .parametric_circuit import ParametricCircuit, is_parametric_circuit

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
    assert is_parametric_circuit(circ)
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


# ------------------------------------------------------------------------------------------------
# Gate-by-gate primitives (core_operations.py:34-603 of the reference): what its unit tests and tools
# call directly.  Same names, argument order and in-place behaviour; the gate runs on the GPU
# (csrc/aqc_prim.cu) on a copy of the host vector.  The workspace arguments are checked like the
# reference checks them but are not needed.  Qubit ``pos`` is index bit ``n - 1 - pos``.
# ------------------------------------------------------------------------------------------------
from . import _prim  # noqa: E402
from . import elementary_operations as _eo  # noqa: E402

_P00 = np.array([[1, 0], [0, 0]], dtype=np.complex128)
_P11 = np.array([[0, 0], [0, 1]], dtype=np.complex128)


def bit2bit_transform(n: int, i: int) -> int:
    """Qubit position -> bit of the flat index (Qiskit bit order, core_operations.py:34-43)."""
    return n - 1 - i


def _stride(n: int, pos: int) -> int:
    return 1 << (n - 1 - pos)


def _check_vec(n: int, *vecs) -> None:
    for v in vecs:
        assert isinstance(v, np.ndarray) and v.shape == (2**n,) and v.dtype == np.complex128
        assert v.flags.c_contiguous


def gate2x2_mul_vec(num_qubits: int, pos: int, gate2x2: np.ndarray, vec: np.ndarray, out: np.ndarray,
                    inplace: bool) -> np.ndarray:
    """``(I (x) G (x) I) @ vec`` into ``vec`` (inplace) or into ``out`` (core_operations.py:46-119)."""
    assert 0 <= pos < num_qubits and gate2x2.shape == (2, 2)
    _check_vec(num_qubits, vec, out)
    assert not np.may_share_memory(vec, out)
    dst = vec
    if not inplace:
        np.copyto(out, vec)
        dst = out
    return _prim.apply_gates(dst, [(_stride(num_qubits, pos), 0, 0, gate2x2)])


def proj00_mul_vec(num_qubits: int, pos: int, vec: np.ndarray) -> np.ndarray:
    """|0><0| on qubit ``pos``, in place (:122-140)."""
    assert 0 <= pos < num_qubits
    _check_vec(num_qubits, vec)
    return _prim.apply_gates(vec, [(_stride(num_qubits, pos), 0, 0, _P00)])


def proj11_mul_vec(num_qubits: int, pos: int, vec: np.ndarray) -> np.ndarray:
    """|1><1| on qubit ``pos``, in place (:143-161)."""
    assert 0 <= pos < num_qubits
    _check_vec(num_qubits, vec)
    return _prim.apply_gates(vec, [(_stride(num_qubits, pos), 0, 0, _P11)])


def _rot_mul_vec(make, n, pos, angle, vec, temp):
    assert 0 <= pos < n and chk.is_float(angle)
    _check_vec(n, vec)
    assert temp is None or (isinstance(temp, np.ndarray) and temp.shape == vec.shape)
    return _prim.apply_gates(vec, [(_stride(n, pos), 0, 0, make(float(angle)))])


def rx_mul_vec(n: int, pos: int, angle: float, vec: np.ndarray, temp: np.ndarray) -> np.ndarray:
    """Rx(angle) on qubit ``pos``, in place (:164-197)."""
    return _rot_mul_vec(_eo.np_rx, n, pos, angle, vec, temp)


def ry_mul_vec(n: int, pos: int, angle: float, vec: np.ndarray, temp: np.ndarray) -> np.ndarray:
    """Ry(angle) on qubit ``pos``, in place (:200-233)."""
    return _rot_mul_vec(_eo.np_ry, n, pos, angle, vec, temp)


def rz_mul_vec(n: int, pos: int, angle: float, vec: np.ndarray, _: Optional[np.ndarray] = None) -> np.ndarray:
    """Rz(angle) on qubit ``pos``, in place (:236-264)."""
    return _rot_mul_vec(_eo.np_rz, n, pos, angle, vec, None)


def _pauli_dot(pauli, n, pos, w_vec, z_vec, temp) -> np.complex128:
    assert 0 <= pos < n
    _check_vec(n, w_vec, z_vec, temp)
    return np.complex128(0.5j * _prim.gate_vdot(w_vec, z_vec, (_stride(n, pos), 0, 0, pauli)))


def dot_x(n: int, pos: int, w_vec: np.ndarray, z_vec: np.ndarray, temp: np.ndarray) -> np.complex128:
    """``0.5j <X w|z>`` (:267-293)."""
    return _pauli_dot(_eo.np_x(), n, pos, w_vec, z_vec, temp)


def dot_y(n: int, pos: int, w_vec: np.ndarray, z_vec: np.ndarray, temp: np.ndarray) -> np.complex128:
    """``0.5j <Y w|z>`` (:296-322)."""
    return _pauli_dot(_eo.np_y(), n, pos, w_vec, z_vec, temp)


def dot_z(n: int, pos: int, w_vec: np.ndarray, z_vec: np.ndarray, temp: np.ndarray) -> np.complex128:
    """``0.5j <Z w|z>`` (:325-351)."""
    return _pauli_dot(_eo.np_z(), n, pos, w_vec, z_vec, temp)


def block_mul_vec(n: int, c: int, t: int, c_mat: np.ndarray, t_mat: np.ndarray, g_mat: np.ndarray,
                  vec: np.ndarray, workspace: np.ndarray, dagger: bool) -> np.ndarray:
    """
    ``vec <- (c_mat (x) t_mat) . controlled-g_mat . vec`` in place; ``dagger=True`` only flips the
    order of the two factors, the matrices are taken as given (:354-419).
    """
    assert 0 <= c < n and 0 <= t < n and c != t
    assert c_mat.shape == t_mat.shape == g_mat.shape == (2, 2)
    _check_vec(n, vec)
    assert workspace.shape == (2, vec.size) and not np.may_share_memory(vec, workspace)
    ent = (_stride(n, t), _stride(n, c), 1, g_mat)
    local = [(_stride(n, c), 0, 0, c_mat), (_stride(n, t), 0, 0, t_mat)]
    return _prim.apply_gates(vec, local + [ent] if dagger else [ent] + local)


def _ctrl_mul_vec(gate, n, c, t, vec, temp):
    assert 0 <= c < n and 0 <= t < n and c != t
    _check_vec(n, vec)
    assert temp is None or (isinstance(temp, np.ndarray) and temp.shape == vec.shape)
    return _prim.apply_gates(vec, [(_stride(n, t), _stride(n, c), 1, gate)])


def cx_mul_vec(n: int, c: int, t: int, _: float, vec: np.ndarray, temp: np.ndarray) -> np.ndarray:
    """CX with control ``c`` and target ``t``, in place (:422-465)."""
    return _ctrl_mul_vec(_eo.np_x(), n, c, t, vec, temp)


def cz_mul_vec(n: int, c: int, t: int, _: float, vec: np.ndarray, temp: np.ndarray) -> np.ndarray:
    """CZ, in place (:468-511)."""
    return _ctrl_mul_vec(_eo.np_z(), n, c, t, vec, temp)


def cp_mul_vec(n: int, c: int, t: int, angle: float, vec: np.ndarray, temp: np.ndarray) -> np.ndarray:
    """CPhase(angle), in place (:514-558)."""
    assert chk.is_float(angle)
    return _ctrl_mul_vec(_eo.np_phase(float(angle)), n, c, t, vec, temp)


def derv_cphase_mul_vec(n: int, c: int, t: int, angle: float, vec: np.ndarray, out: np.ndarray) -> np.ndarray:
    """``out = (|1><1|_c (x) dP(angle)/d angle _t) vec`` (:561-603); ``vec`` is not modified."""
    assert 0 <= c < n and 0 <= t < n and c != t and chk.is_float(angle)
    _check_vec(n, vec, out)
    assert not np.may_share_memory(vec, out)
    np.copyto(out, vec)
    dgate = np.array([[0, 0], [0, 1j * np.exp(1j * float(angle))]], dtype=np.complex128)
    return _prim.apply_gates(out, [(_stride(n, t), _stride(n, c), 2, dgate)])
