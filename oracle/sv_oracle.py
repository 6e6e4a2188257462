"""
CPU ORACLE (test infrastructure only) -- NumPy restatement of the reference's state-vector and
matrix objective/gradient algorithm.  It is NOT part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import it.  The product path (aqc_research_b200) never does.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function below against
outputs of the unmodified reference (generated in the build container by
``tests/golden/make_golden.py`` from /root/reference, committed as ``tests/golden/*.npz``) and,
when /root/reference is present, against the reference imported live.

Reference (qiskit-community/aqc-research) locations restated here, paths relative to its root:
  gate conventions            aqc_research/elementary_operations.py:143-291
  bit order (qubit q = bit q) aqc_research/core_operations.py:34-43,77
  V @ vec, V^H @ vec          aqc_research/core_operations.py:606-710, 713-820
  gradient sweep              aqc_research/core_operations.py:823-1019
  matrix variants             aqc_research/core_op_matrix.py:480-762
  surrogate objective         aqc_research/model_sp_lhs/objective_lhs_sur_max.py:82-191
  sketching objective         aqc_research/model_sketching/sk_core.py:167-222
  coordinate descent          aqc_research/core_op_matrix.py:765-917   (coord_descent_sweep; pinned by
                              tests/golden/cd_cases.npz: 8 reference runs of 4 consecutive sweeps)
  sketching generators        aqc_research/model_sketching/sk_core.py:329-464   (SketchOracle; pinned by
                              tests/golden/sketch_cases.npz: 7 generator + objective sequences)

The restatement is deliberately written differently from the reference (one generic
``(-1, 2, stride)`` reshape per gate instead of half-slice arithmetic) so that agreement is
evidence about the algorithm rather than about shared code.
"""

from typing import Optional, Tuple
import numpy as np

C128 = np.complex128


# ------------------------------------------------------------------------------------------------
# gates (elementary_operations.py:143-291)
# ------------------------------------------------------------------------------------------------
def rx(phi: float) -> np.ndarray:
    c, s = np.cos(0.5 * phi), np.sin(0.5 * phi)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=C128)


def ry(phi: float) -> np.ndarray:
    c, s = np.cos(0.5 * phi), np.sin(0.5 * phi)
    return np.array([[c, -s], [s, c]], dtype=C128)


def rz(phi: float) -> np.ndarray:
    return np.array([[np.exp(-0.5j * phi), 0], [0, np.exp(0.5j * phi)]], dtype=C128)


def phase(phi: float) -> np.ndarray:
    return np.array([[1, 0], [0, np.exp(1j * phi)]], dtype=C128)


PAULI_X = np.array([[0, 1], [1, 0]], dtype=C128)
PAULI_Y = np.array([[0, -1j], [1j, 0]], dtype=C128)
PAULI_Z = np.array([[1, 0], [0, -1]], dtype=C128)


# ------------------------------------------------------------------------------------------------
# circuit description helpers (duck-typed: works with our classes and with the reference's)
# ------------------------------------------------------------------------------------------------
def _is_trotter(circ) -> bool:
    return hasattr(circ, "is_second_order")


def _num_extra_blocks(circ) -> int:
    """Blocks of the implied trailing half-layer (core_operations.py:638-641)."""
    if _is_trotter(circ) and circ.is_second_order and circ.num_blocks > 0:
        return 3 * (circ.num_qubits // 2)
    return 0


def _layout(circ, ncols: int):
    """
    (trotterized, total number of blocks).  The matrix routines of the reference
    (core_op_matrix.py:480-762) have no Trotter special-casing: with ncols > 1 every circuit is
    treated as a generic ParametricCircuit, exactly as the reference does.
    """
    if ncols == 1 and _is_trotter(circ):
        return True, circ.num_blocks + _num_extra_blocks(circ)
    return False, circ.num_blocks


def _swappable(circ):
    """Rs gate / Pauli that commutes with the entangler on the target (core_operations.py:652-668)."""
    return (rx, PAULI_X) if circ.entangler == "cx" else (rz, PAULI_Z)


# ------------------------------------------------------------------------------------------------
# primitive operations on a flat array of 2^n * ncols numbers; qubit q <-> bit q of the ROW index
# ------------------------------------------------------------------------------------------------
def op1(state: np.ndarray, q: int, g: np.ndarray, ncols: int = 1) -> np.ndarray:
    """state <- (I x g_q x I) state, returns a new array."""
    v = state.reshape(-1, 2, (1 << q) * ncols)
    return np.einsum("ab,xbz->xaz", g, v).reshape(state.shape)


def ctrl_op(state: np.ndarray, c: int, t: int, g: np.ndarray, ncols: int = 1) -> np.ndarray:
    """state <- (|0><0|_c x I + |1><1|_c x g_t) state (cx/cz/cp_mul_vec, core_operations.py:422-558)."""
    out = state.copy()
    n_idx = out.size // ncols
    rows = np.arange(n_idx)
    sel = (rows >> c) & 1 == 1
    sub = out.reshape(n_idx, ncols)
    lo = rows[sel & ((rows >> t) & 1 == 0)]
    hi = lo | (1 << t)
    a, b = sub[lo].copy(), sub[hi].copy()
    sub[lo] = g[0, 0] * a + g[0, 1] * b
    sub[hi] = g[1, 0] * a + g[1, 1] * b
    return out


def pauli_dot(w: np.ndarray, z: np.ndarray, q: int, pauli: np.ndarray, ncols: int = 1) -> complex:
    """0.5j * <P_q w | z>   (dot_x / dot_y / dot_z, core_operations.py:267-351)."""
    return 0.5j * np.vdot(op1(w, q, pauli, ncols), z)


def _entangler_gate(circ, tht, dagger: bool) -> np.ndarray:
    if circ.entangler == "cx":
        return PAULI_X
    if circ.entangler == "cz":
        return PAULI_Z
    return phase(-tht[4] if dagger else tht[4])


# ------------------------------------------------------------------------------------------------
# V @ state and V^H @ state
# ------------------------------------------------------------------------------------------------
def apply_v(circ, thetas: np.ndarray, state: np.ndarray, dagger: bool = False, ncols: int = 1):
    """
    Returns V(thetas) @ state (dagger=False; v_mul_vec core_operations.py:606-710, v_mul_mat
    core_op_matrix.py:480-559) or V^H @ state (dagger=True; core_operations.py:713-820,
    core_op_matrix.py:562-642).  ``state`` is a flat complex128 array of 2^n * ncols entries
    (row-major (2^n, ncols) matrix when ncols > 1).
    """
    n, nb = circ.num_qubits, circ.num_blocks
    tpb = 5 if circ.entangler == "cp" else 4
    th1 = thetas[: 3 * n].reshape(n, 3)
    th2 = thetas[3 * n :].reshape(nb, tpb)
    make_rs, _ = _swappable(circ)
    trot, total = _layout(circ, ncols)
    s = np.array(state, dtype=C128).ravel().copy()

    def front(s):
        for q in range(n):
            a0, a1, a2 = th1[q]
            g = rz(a0) @ ry(a1) @ rz(a2)  # core_operations.py:671-677
            s = op1(s, q, g.conj().T if dagger else g, ncols)
        return s

    def block(s, i):
        k = i % nb
        c, t = int(circ.blocks[0, k]), int(circ.blocks[1, k])
        tht = th2[k]
        cm = rz(tht[1]) @ ry(tht[0])
        tm = make_rs(tht[3]) @ ry(tht[2])
        e = _entangler_gate(circ, tht, dagger)
        if not dagger:
            if trot and i % 3 == 0:
                s = op1(s, c, rz(-np.pi / 2), ncols)
            s = ctrl_op(s, c, t, e, ncols)
            s = op1(s, c, cm, ncols)
            s = op1(s, t, tm, ncols)
            if trot and i % 3 == 2:
                s = op1(s, t, rz(np.pi / 2), ncols)
        else:
            if trot and i % 3 == 2:
                s = op1(s, t, rz(-np.pi / 2), ncols)
            s = op1(s, t, tm.conj().T, ncols)
            s = op1(s, c, cm.conj().T, ncols)
            s = ctrl_op(s, c, t, e, ncols)
            if trot and i % 3 == 0:
                s = op1(s, c, rz(np.pi / 2), ncols)
        return s

    if not dagger:
        s = front(s)
        for i in range(total):
            s = block(s, i)
    else:
        for i in range(total - 1, -1, -1):
            s = block(s, i)
        s = front(s)
    return s.reshape(np.shape(state))


# ------------------------------------------------------------------------------------------------
# gradient sweep
# ------------------------------------------------------------------------------------------------
def grad_sweep(
    circ,
    thetas: np.ndarray,
    x: np.ndarray,
    z0: np.ndarray,
    block_range: Optional[Tuple[int, int]] = None,
    front_layer: bool = True,
    ncols: int = 1,
) -> np.ndarray:
    """
    Complex gradient of <V x | y> given z0 = V^H y (grad_of_dot_product
    core_operations.py:823-1019; grad_of_matrix_dot_product core_op_matrix.py:645-762).
    Both w (from x) and z (from z0) are pushed through the circuit; after each rotation R_P the
    entry 0.5j <P w|z> is recorded.  Derivatives of the implied trailing half-layer are ADDED to
    those of the leading half-layer (:966-994).
    """
    n, nb = circ.num_qubits, circ.num_blocks
    tpb = 5 if circ.entangler == "cp" else 4
    th1 = thetas[: 3 * n].reshape(n, 3)
    th2 = thetas[3 * n :].reshape(nb, tpb)
    make_rs, pauli_s = _swappable(circ)
    trot, total = _layout(circ, ncols)
    lo_b, hi_b = (0, nb) if block_range is None else block_range
    w = np.array(x, dtype=C128).ravel().copy()
    z = np.array(z0, dtype=C128).ravel().copy()
    grad = np.zeros(thetas.size, dtype=C128)
    g1 = grad[: 3 * n].reshape(n, 3)
    g2 = grad[3 * n :].reshape(nb, tpb)

    def rot(q, gate, pauli):
        nonlocal w, z
        w, z = op1(w, q, gate, ncols), op1(z, q, gate, ncols)
        return pauli_dot(w, z, q, pauli, ncols)

    for q in range(n):  # :919-949
        d2 = rot(q, rz(th1[q, 2]), PAULI_Z)
        d1 = rot(q, ry(th1[q, 1]), PAULI_Y)
        d0 = rot(q, rz(th1[q, 0]), PAULI_Z)
        if front_layer:
            g1[q] = d0, d1, d2

    for i in range(total):  # :956-1017
        k = i % nb
        c, t = int(circ.blocks[0, k]), int(circ.blocks[1, k])
        tht = th2[k]
        rec = lo_b <= k < hi_b
        if trot and i % 3 == 0:
            w, z = op1(w, c, rz(-np.pi / 2), ncols), op1(z, c, rz(-np.pi / 2), ncols)
        e = _entangler_gate(circ, tht, False)
        z = ctrl_op(z, c, t, e, ncols)
        if circ.entangler == "cp":
            # derivative of the phase gate: |1><1|_c x i e^{i phi}|1><1|_t  (:561-603, 972-975)
            rows = np.arange(w.size // ncols)
            both = ((rows >> c) & 1 == 1) & ((rows >> t) & 1 == 1)
            dw = np.zeros_like(w).reshape(-1, ncols)
            dw[both] = 1j * np.exp(1j * tht[4]) * w.reshape(-1, ncols)[both]
            if rec:
                g2[k, 4] += np.vdot(dw.ravel(), z)
        w = ctrl_op(w, c, t, e, ncols)
        d = [
            rot(c, ry(tht[0]), PAULI_Y),
            rot(c, rz(tht[1]), PAULI_Z),
            rot(t, ry(tht[2]), PAULI_Y),
            rot(t, make_rs(tht[3]), pauli_s),
        ]
        if rec:
            g2[k, 0:4] += d
        if trot and i % 3 == 2:
            w, z = op1(w, t, rz(np.pi / 2), ncols), op1(z, t, rz(np.pi / 2), ncols)
    return grad


# ------------------------------------------------------------------------------------------------
# dense matrix of the circuit by Kronecker products (independent small-n check)
# ------------------------------------------------------------------------------------------------
def dense_unitary(circ, thetas: np.ndarray) -> np.ndarray:
    """2^n x 2^n matrix of V(thetas): columns V e_k (test_core_operations.py:283-321 identity)."""
    dim = 1 << circ.num_qubits
    eye = np.eye(dim, dtype=C128)
    return np.stack([apply_v(circ, thetas, eye[:, k]) for k in range(dim)], axis=1)


def kron_1q(n: int, q: int, g: np.ndarray) -> np.ndarray:
    """Full matrix of a 1-qubit gate on qubit q (bit q, little-endian => leftmost factor is qubit n-1)."""
    m = np.eye(1, dtype=C128)
    for k in range(n - 1, -1, -1):
        m = np.kron(m, g if k == q else np.eye(2, dtype=C128))
    return m


def kron_ctrl(n: int, c: int, t: int, g: np.ndarray) -> np.ndarray:
    """Full matrix of |0><0|_c x I + |1><1|_c x g_t."""
    p0 = np.array([[1, 0], [0, 0]], dtype=C128)
    p1 = np.array([[0, 0], [0, 1]], dtype=C128)
    return kron_1q(n, c, p0) + kron_1q(n, c, p1) @ kron_1q(n, t, g)


def dense_unitary_kron(circ, thetas: np.ndarray) -> np.ndarray:
    """Same matrix as ``dense_unitary`` but assembled from explicit Kronecker products."""
    n, nb = circ.num_qubits, circ.num_blocks
    tpb = 5 if circ.entangler == "cp" else 4
    th1 = thetas[: 3 * n].reshape(n, 3)
    th2 = thetas[3 * n :].reshape(nb, tpb)
    make_rs, _ = _swappable(circ)
    trot = _is_trotter(circ)
    m = np.eye(1 << n, dtype=C128)
    for q in range(n):
        m = kron_1q(n, q, rz(th1[q, 0]) @ ry(th1[q, 1]) @ rz(th1[q, 2])) @ m
    for i in range(nb + _num_extra_blocks(circ)):
        k = i % nb
        c, t = int(circ.blocks[0, k]), int(circ.blocks[1, k])
        tht = th2[k]
        if trot and i % 3 == 0:
            m = kron_1q(n, c, rz(-np.pi / 2)) @ m
        m = kron_ctrl(n, c, t, _entangler_gate(circ, tht, False)) @ m
        m = kron_1q(n, c, rz(tht[1]) @ ry(tht[0])) @ m
        m = kron_1q(n, t, make_rs(tht[3]) @ ry(tht[2])) @ m
        if trot and i % 3 == 2:
            m = kron_1q(n, t, rz(np.pi / 2)) @ m
    return m


# ------------------------------------------------------------------------------------------------
# objectives (stateless pieces; the stateful shells live in the product and in the reference)
# ------------------------------------------------------------------------------------------------
def basis_state_indices(num_qubits: int, init_index: int = 0) -> np.ndarray:
    """
    Indices of |s>, X_0|s>, ..., X_{n-1}|s> for the basis state s = init_index
    (ThinStateHandler with max_flips = 1, objective_base.py:42-173; little-endian bits).
    """
    return np.array([init_index] + [init_index ^ (1 << q) for q in range(num_qubits)], dtype=np.int64)


def sur_max_value_and_grad(circ, thetas, target, weight: float, max_no: int, init_index: int = 0):
    """
    One stateless evaluation of the surrogate of objective_lhs_sur_max.py:82-191 for FIXED
    ``weight`` and ``max_no`` (the reference evolves them across calls; parity tests replay the
    same sequence on both sides).  Returns (f, hs, real gradient, raw complex grads).
    """
    idx = basis_state_indices(circ.num_qubits, init_index)
    z0 = apply_v(circ, thetas, target, dagger=True)
    hs = z0[idx]
    hs2 = np.abs(hs) ** 2
    f = 1.0 - (1.0 - weight) * hs2[0] - weight * hs2[max_no]
    e0 = np.zeros(z0.size, dtype=C128)
    e0[idx[0]] = 1
    g0 = grad_sweep(circ, thetas, e0, z0)
    if max_no == 0:
        grad = np.real(-2.0 * np.conj(hs[0]) * g0)
        gm = None
    else:
        em = np.zeros(z0.size, dtype=C128)
        em[idx[max_no]] = 1
        gm = grad_sweep(circ, thetas, em, z0)
        grad = np.real(-2.0 * (1.0 - weight) * np.conj(hs[0]) * g0) + np.real(
            -2.0 * weight * np.conj(hs[max_no]) * gm
        )
    return f, hs, grad, (g0, gm)


def sketch_full_value_and_grad(circ, thetas, target_matrix):
    """
    f = 1 - Re Tr(V^H U)/d, grad = -Re(grad <V I|U>)/d  (sk_core.py:167-222 with
    FullRangeSketchingVectors :300-326).
    """
    d = target_matrix.shape[0]
    vh_y = apply_v(circ, thetas, target_matrix.ravel(), dagger=True, ncols=d).reshape(d, d)
    f = 1.0 - np.real(np.trace(vh_y)) / d
    g = grad_sweep(circ, thetas, np.eye(d, dtype=C128).ravel(), vh_y.ravel(), ncols=d)
    return f, -np.real(g) / d


# ------------------------------------------------------------------------------------------------
# coordinate descent for unitary AQC
# ------------------------------------------------------------------------------------------------
CD_LEARN_RATE = np.pi / 16
CD_MAX_DELTA = np.pi / 4


def cd_delta_theta(prod: complex, grad: complex, dim: int) -> float:
    """Angle increment of one coordinate (``_delta_theta``, core_op_matrix.py:833-850)."""
    tol = float(np.sqrt(np.finfo(np.float64).eps))
    derv1 = (-2.0 * np.real(np.conj(prod) * grad)) / (dim**2)
    derv2 = (-2.0 * abs(grad) ** 2 + 0.5 * abs(prod) ** 2) / (dim**2)
    if derv2 < tol:
        derv1 /= max(abs(derv1), 1.0)
        dt = -CD_LEARN_RATE * derv1
    else:
        dt = -derv1 / derv2
    a = abs(dt / CD_MAX_DELTA)
    return float(dt if a <= 1 else dt / a)


def coord_descent_sweep(circ, thetas: np.ndarray, target: np.ndarray):
    """
    One coordinate-descent sweep over all angles (coord_descent_single_sweep,
    core_op_matrix.py:765-917): w = I, z = V^H U; before every rotation R_P(theta_k) the product
    <w|z>_F and 0.5j<P w|z>_F give a Newton / clipped gradient step for theta_k; z is rotated
    with the OLD angle, w with the NEW one.  Returns (fobj, new_thetas); cx and cz only.
    """
    if circ.entangler == "cp":
        raise NotImplementedError("CPhase entangler is not supported yet")
    n, nb = circ.num_qubits, circ.num_blocks
    dim = 1 << n
    th = np.array(thetas, dtype=np.float64).copy()
    th1 = th[: 3 * n].reshape(n, 3)
    th2 = th[3 * n :].reshape(nb, 4)
    make_rs, pauli_s = _swappable(circ)
    w = np.eye(dim, dtype=C128).ravel()
    z = apply_v(circ, th, np.array(target, dtype=C128).ravel(), dagger=True, ncols=dim)

    def step(q, make_gate, pauli, tht, k):
        nonlocal w, z
        grad = pauli_dot(w, z, q, pauli, dim)
        prod = np.vdot(w, z)
        z = op1(z, q, make_gate(tht[k]), dim)
        tht[k] += cd_delta_theta(prod, grad, dim)
        w = op1(w, q, make_gate(tht[k]), dim)

    for q in range(n):
        step(q, rz, PAULI_Z, th1[q], 2)
        step(q, ry, PAULI_Y, th1[q], 1)
        step(q, rz, PAULI_Z, th1[q], 0)
    e = PAULI_X if circ.entangler == "cx" else PAULI_Z
    for i in range(nb):
        c, t = int(circ.blocks[0, i]), int(circ.blocks[1, i])
        z = ctrl_op(z, c, t, e, dim)
        w = ctrl_op(w, c, t, e, dim)
        step(c, ry, PAULI_Y, th2[i], 0)
        step(c, rz, PAULI_Z, th2[i], 1)
        step(t, ry, PAULI_Y, th2[i], 2)
        step(t, make_rs, pauli_s, th2[i], 3)
    return float(1 - np.abs(np.vdot(w, z) / dim) ** 2), th


# ------------------------------------------------------------------------------------------------
# sketching-vector generators: sk_core.py:329-464, same global-RNG call order
# ------------------------------------------------------------------------------------------------
class SketchOracle:
    """
    NumPy restatement of Random / Alternating / Eigen SketchingVectors.generate followed by
    SketchingObjectiveEx.objective_and_gradient (sk_core.py:167-222).  Draws from the GLOBAL NumPy
    RNG in the reference's order, so after ``np.random.seed(s)`` it sees the same random numbers.
    """

    def __init__(self, kind: str, num_skvecs: int, target: np.ndarray):
        self.kind, self.m, self.target = kind, int(num_skvecs), np.asarray(target, dtype=C128)
        self.dim = self.target.shape[0]
        if kind == "alt":  # sk_core.py:375-379
            self.offset = 0
            self.indices = np.random.permutation(self.dim)

    def generate(self, circ=None, thetas=None):
        d, m, u = self.dim, self.m, self.target
        if self.kind == "rand":  # :347-359
            x, _ = np.linalg.qr(np.random.rand(d, m) + 1j * np.random.rand(d, m))
        elif self.kind == "alt":  # :381-407
            if self.offset >= d:
                self.offset = 0
                self.indices = np.random.permutation(d)
            idx = self.indices[self.offset : self.offset + m]
            x = np.zeros((d, m), dtype=C128)
            x[idx, np.arange(idx.size)] = 1
            self.offset += m
        elif self.kind == "eigen":  # :424-462
            omega = 1j * np.random.randn(d, m)
            omega = omega + np.random.randn(d, m)
            vh_om = apply_v(circ, thetas, omega.ravel(), dagger=True, ncols=m).reshape(d, m)
            x, _ = np.linalg.qr(vh_om - u.conj().T @ omega)
        else:
            raise ValueError(self.kind)
        return x, u @ x

    def value_and_grad(self, circ, thetas):
        x, y = self.generate(circ, thetas)
        m = self.m
        vh_y = apply_v(circ, thetas, y.ravel(), dagger=True, ncols=m)
        f = 1.0 - np.real(np.vdot(x.ravel(), vh_y)) / m
        g = grad_sweep(circ, thetas, x.ravel(), vh_y, ncols=m)
        return float(f), -np.real(g) / m
