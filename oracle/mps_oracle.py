"""
CPU ORACLE (test infrastructure only) for the MPS path.  NOT part of the product.

What is restated
  * the MPS data format and its interpretation: ``QiskitMPS = (gammas, lambdas)``, Vidal form,
    site k <-> qubit k <-> bit k of the flat index; amplitude
    <b|psi> = G_0[b_0] diag(l_0) G_1[b_1] ... G_{n-1}[b_{n-1}]
    (reference: aqc_research/mps_operations.py:33, check_mps :87-123, _preprocess_mps :126-156,
    mps_to_vector :159-189);
  * ``mps_dot`` (mps_operations.py:192-213): left-to-right transfer contraction;
  * the gate-by-gate gradient sweep ``fast_dot_gradient`` (mps_dot_objective.py:41-242): every
    gate applied to w and z separately, every derivative 0.5j <P w|z> by a full-chain mps_dot;
    CPhase derivative by the parameter-shift difference of :186-196;
  * ``v_mul_mps`` / ``v_dagger_mul_mps`` (mps_operations.py:326-371) with the gate order of
    ansatz_to_qcircuit (circuit_transform.py:200-243).

Parity status
  * format, mps_to_vector, mps_dot: PINNED against the reference's own pure-NumPy functions
    (golden vectors tests/golden/mps_cases.npz + live import when /root/reference exists);
  * gate application: the reference delegates it to qiskit-aer (third-party C++, not in
    /root/reference, no pinned version: requirements.txt:1-8 lists neither qiskit-aer nor a
    version).  Its published algorithm is restated here: contract the two sites with the
    neighbouring lambdas, apply the 4x4 gate, SVD, drop singular values <= 1e-16, drop the smallest
    remaining Schmidt values while the sum of their squares stays below ``trunc_thr``, renormalise
    if anything was dropped, divide the outer lambdas back out.  The reference runs the simulator
    WITHOUT a bond cap; the GPU engine's capacity ``chi_max`` (<= 64) is applied AFTER that rule, so it
    binds only where the rule alone would keep more (the engine counts and reports those cases).
    UNTRUNCATED results (trunc_thr = 1e-16) are pinned: they must equal the state-vector path
    (what test_mps.py / test_mps_fast_dot_gradient.py of the reference assert).  TRUNCATED
    results are "parity unpinned" with respect to qiskit-aer (no bit-for-bit pin is possible without
    it); the GPU tests bound them against the exact state-vector oracle through the discarded weight.
"""

from typing import List, Optional, Tuple
import numpy as np
from . import sv_oracle as O

C128 = np.complex128
CHOP = 1e-16  # singular values at or below this are treated as exact zeros

QiskitMPS = Tuple[List[Tuple[np.ndarray, np.ndarray]], List[np.ndarray]]


# ------------------------------------------------------------------------------------------------
# format
# ------------------------------------------------------------------------------------------------
def site_matrices(mps: QiskitMPS) -> List[np.ndarray]:
    """A_k[b] = G_k[b] diag(l_k) (no lambda after the last site); shape (2, chi_k, chi_{k+1})."""
    gam, lam = mps
    out = []
    for k, (g0, g1) in enumerate(gam):
        a = np.stack((np.asarray(g0, dtype=C128), np.asarray(g1, dtype=C128)))
        if k < len(gam) - 1:
            a = a * np.asarray(lam[k], dtype=np.float64).ravel()[None, None, :]
        out.append(a)
    return out


def mps_to_vector(mps: QiskitMPS) -> np.ndarray:
    """State vector of 2^n amplitudes; bit k of the index = physical index of site k."""
    mats = site_matrices(mps)
    n = len(mats)
    # psi[b_0 ... b_{k}, bond]: grow one site at a time; new site index is the SLOWEST so far
    psi = mats[0].reshape(2, -1)  # (b0, chi_1)   (chi_0 = 1)
    for k in range(1, n):
        psi = np.einsum("xa,bac->bxc", psi, mats[k]).reshape(-1, mats[k].shape[2])
    return psi.reshape(-1)


def mps_dot(m1: QiskitMPS, m2: QiskitMPS) -> complex:
    """<m1|m2> by the transfer-matrix recursion E <- sum_b A1[b]^H E A2[b]."""
    a, b = site_matrices(m1), site_matrices(m2)
    env = np.ones((1, 1), dtype=C128)
    for x, y in zip(a, b):
        env = sum(x[s].conj().T @ env @ y[s] for s in range(2))
    return complex(env.item())


def vector_to_mps(vec: np.ndarray, chop: float = CHOP) -> QiskitMPS:
    """Exact Vidal decomposition of a state vector (numerical-rank bonds)."""
    n = int(round(np.log2(vec.size)))
    gam, lam = [], []
    rest = np.asarray(vec, dtype=C128).reshape(1, -1)  # (chi_left, remaining amplitudes)
    prev_lam = np.ones(1)
    for k in range(n - 1):
        chi = rest.shape[0]
        # split off bit k (fastest varying of the remaining index)
        m = rest.reshape(chi, -1, 2).transpose(0, 2, 1).reshape(chi * 2, -1)  # rows (alpha, b)
        u, s, vh = np.linalg.svd(m, full_matrices=False)
        keep = max(1, int(np.sum(s > chop)))
        u, s, vh = u[:, :keep], s[:keep], vh[:keep]
        u = u.reshape(chi, 2, keep) / prev_lam[:, None, None]
        gam.append((u[:, 0, :].copy(), u[:, 1, :].copy()))
        lam.append(s.copy())
        rest = s[:, None] * vh
        prev_lam = s
    last = rest.reshape(rest.shape[0], 2) / prev_lam[:, None]
    gam.append((last[:, 0:1].copy(), last[:, 1:2].copy()))
    return gam, lam


def product_state(num_qubits: int, index: int) -> QiskitMPS:
    """Basis state |index> as a bond-dimension-1 MPS."""
    gam = []
    for q in range(num_qubits):
        bit = (index >> q) & 1
        g0 = np.array([[1.0 - bit]], dtype=C128)
        g1 = np.array([[float(bit)]], dtype=C128)
        gam.append((g0, g1))
    return gam, [np.ones(1) for _ in range(num_qubits - 1)]


def random_mps(num_qubits: int, chi: int, rng: np.random.RandomState) -> QiskitMPS:
    """
    Random normalised MPS in Vidal form with bond dimensions min(chi, 2^k, 2^(n-k)): random site
    tensors, then a right-to-left QR sweep and a left-to-right SVD sweep (canonicalisation).
    """
    n = num_qubits
    dims = [min(chi, 2 ** min(k, n - k)) for k in range(n + 1)]
    a = [rng.randn(2, dims[k], dims[k + 1]) + 1j * rng.randn(2, dims[k], dims[k + 1]) for k in range(n)]
    # right-canonicalise
    for k in range(n - 1, 0, -1):
        m = a[k].transpose(1, 0, 2).reshape(dims[k], -1)  # (left, b*right)
        q, r = np.linalg.qr(m.conj().T)  # m^H = q r  ->  m = r^H q^H
        a[k] = q.conj().T.reshape(dims[k], 2, dims[k + 1]).transpose(1, 0, 2)
        a[k - 1] = np.einsum("bxy,yz->bxz", a[k - 1], r.conj().T)
    a[0] /= np.linalg.norm(a[0])
    # left-to-right SVD sweep -> Vidal form
    gam, lam = [], []
    prev = np.ones(1)
    carry = np.eye(1, dtype=C128)
    for k in range(n - 1):
        t = np.einsum("xy,byz->bxz", carry, a[k])  # (b, left, right)
        m = t.transpose(1, 0, 2).reshape(t.shape[1] * 2, -1)
        u, s, vh = np.linalg.svd(m, full_matrices=False)
        u = u.reshape(t.shape[1], 2, -1) / prev[:, None, None]
        gam.append((u[:, 0, :].copy(), u[:, 1, :].copy()))
        lam.append(s.copy())
        carry = s[:, None] * vh
        prev = s
    t = np.einsum("xy,byz->bxz", carry, a[n - 1]) / prev[None, :, None]
    gam.append((t[0].copy(), t[1].copy()))
    return gam, lam


def copy_mps(mps: QiskitMPS) -> QiskitMPS:
    return [(g0.copy(), g1.copy()) for g0, g1 in mps[0]], [np.array(l, dtype=np.float64).copy() for l in mps[1]]


def bond_dims(mps: QiskitMPS) -> List[int]:
    return [int(np.asarray(l).size) for l in mps[1]]


# ------------------------------------------------------------------------------------------------
# gate application (restated qiskit-aer MPS algorithm, see header)
# ------------------------------------------------------------------------------------------------
def apply_1q(mps: QiskitMPS, q: int, g: np.ndarray) -> QiskitMPS:
    gam, lam = mps
    g0, g1 = gam[q]
    gam[q] = (g[0, 0] * g0 + g[0, 1] * g1, g[1, 0] * g0 + g[1, 1] * g1)
    return mps


def truncate_rule(s: np.ndarray, trunc_thr: float, chi_max: Optional[int]) -> Tuple[int, np.ndarray]:
    """Number of kept singular values and the (possibly renormalised) kept values."""
    total = int(np.sum(s > CHOP))
    keep = max(1, total)
    acc = 0.0
    while keep > 1 and acc + s[keep - 1] ** 2 < trunc_thr:  # the reference's rule (no bond cap there)
        acc += s[keep - 1] ** 2
        keep -= 1
    if chi_max is not None:  # the GPU engine's bond capacity binds only if the rule alone keeps more
        keep = min(keep, chi_max)
    kept = s[:keep].copy()
    if keep < total:
        kept /= np.linalg.norm(kept)
    return keep, kept


def apply_2q(mps: QiskitMPS, k: int, gate4: np.ndarray, trunc_thr: float, chi_max: Optional[int]) -> QiskitMPS:
    """
    4x4 gate on sites (k, k+1); gate index = b_k + 2 b_{k+1} (little-endian like the flat index).
    """
    gam, lam = mps
    n = len(gam)
    l_left = np.asarray(lam[k - 1]).ravel() if k > 0 else np.ones(1)
    l_mid = np.asarray(lam[k]).ravel()
    l_right = np.asarray(lam[k + 1]).ravel() if k + 2 < n else np.ones(1)
    a = np.stack(gam[k])  # (b1, L, M)
    b = np.stack(gam[k + 1])  # (b2, M, R)
    theta = np.einsum("l,xlm,m,ymr,r->xylr", l_left, a, l_mid, b, l_right)  # (b1, b2, L, R)
    g = gate4.reshape(2, 2, 2, 2)  # [b2', b1', b2, b1]
    theta = np.einsum("vuyx,xylr->uvlr", g, theta)  # (b1', b2', L, R)
    cl, cr = theta.shape[2], theta.shape[3]
    m = theta.transpose(0, 2, 1, 3).reshape(2 * cl, 2 * cr)  # rows (b1, L), cols (b2, R)
    u, s, vh = np.linalg.svd(m, full_matrices=False)
    keep, kept = truncate_rule(s, trunc_thr, chi_max)
    u = u[:, :keep].reshape(2, cl, keep) / l_left[None, :, None]
    vh = vh[:keep].reshape(keep, 2, cr) / l_right[None, None, :]
    gam[k] = (u[0].copy(), u[1].copy())
    gam[k + 1] = (vh[:, 0, :].copy(), vh[:, 1, :].copy())
    lam[k] = kept
    return mps


def _embed(g_lo: np.ndarray, g_hi: np.ndarray) -> np.ndarray:
    """4x4 matrix of g_hi (site k+1) x g_lo (site k) with index b_k + 2 b_{k+1}."""
    return np.kron(g_hi, g_lo)


def ctrl_gate4(ctrl_is_hi: bool, g: np.ndarray) -> np.ndarray:
    """|0><0|_c x I + |1><1|_c x g_t on a pair of adjacent sites."""
    p0, p1, eye = np.diag([1.0, 0.0]).astype(C128), np.diag([0.0, 1.0]).astype(C128), np.eye(2, dtype=C128)
    if ctrl_is_hi:
        return _embed(eye, p0) + _embed(g, p1)
    return _embed(p0, eye) + _embed(p1, g)


def _entangler(circ, c: int, t: int, angle: float) -> Tuple[int, np.ndarray]:
    if abs(c - t) != 1:
        raise NotImplementedError("MPS path supports unit-blocks on adjacent qubits only")
    g = O.PAULI_X if circ.entangler == "cx" else (O.PAULI_Z if circ.entangler == "cz" else O.phase(angle))
    return min(c, t), ctrl_gate4(c > t, g)


def _total_blocks(circ) -> int:
    return circ.num_blocks + O._num_extra_blocks(circ)


def apply_v(circ, thetas, mps: QiskitMPS, dagger: bool = False, trunc_thr: float = CHOP,
            chi_max: Optional[int] = None) -> QiskitMPS:
    """V @ mps or V^H @ mps, gate by gate (v_mul_mps / v_dagger_mul_mps, mps_operations.py:326-371)."""
    s = copy_mps(mps)
    n, nb = circ.num_qubits, circ.num_blocks
    tpb = 5 if circ.entangler == "cp" else 4
    th1, th2 = thetas[: 3 * n].reshape(n, 3), thetas[3 * n :].reshape(nb, tpb)
    make_rs, _ = O._swappable(circ)
    trot = O._is_trotter(circ)
    total = _total_blocks(circ)
    if not dagger:
        for q in range(n):
            for g in (O.rz(th1[q, 2]), O.ry(th1[q, 1]), O.rz(th1[q, 0])):
                apply_1q(s, q, g)
        for i in range(total):
            k = i % nb
            c, t = int(circ.blocks[0, k]), int(circ.blocks[1, k])
            tht = th2[k]
            if trot and i % 3 == 0:
                apply_1q(s, c, O.rz(-np.pi / 2))
            site, e4 = _entangler(circ, c, t, tht[4] if tpb == 5 else 0.0)
            apply_2q(s, site, e4, trunc_thr, chi_max)
            for q, g in ((c, O.ry(tht[0])), (c, O.rz(tht[1])), (t, O.ry(tht[2])), (t, make_rs(tht[3]))):
                apply_1q(s, q, g)
            if trot and i % 3 == 2:
                apply_1q(s, t, O.rz(np.pi / 2))
    else:
        for i in range(total - 1, -1, -1):
            k = i % nb
            c, t = int(circ.blocks[0, k]), int(circ.blocks[1, k])
            tht = th2[k]
            if trot and i % 3 == 2:
                apply_1q(s, t, O.rz(-np.pi / 2))
            for q, g in ((t, make_rs(-tht[3])), (t, O.ry(-tht[2])), (c, O.rz(-tht[1])), (c, O.ry(-tht[0]))):
                apply_1q(s, q, g)
            site, e4 = _entangler(circ, c, t, -tht[4] if tpb == 5 else 0.0)
            apply_2q(s, site, e4, trunc_thr, chi_max)
            if trot and i % 3 == 0:
                apply_1q(s, c, O.rz(np.pi / 2))
        for q in range(n):
            for g in (O.rz(-th1[q, 0]), O.ry(-th1[q, 1]), O.rz(-th1[q, 2])):
                apply_1q(s, q, g)
    return s


def _pauli_dot(w: QiskitMPS, z: QiskitMPS, q: int, pauli: np.ndarray) -> complex:
    """0.5j <P_q w|z> by a full-chain contraction (dot_x/y/z, mps_dot_objective.py:471-516)."""
    pw = apply_1q(copy_mps(w), q, pauli)
    return 0.5j * mps_dot(pw, z)


def grad_sweep(circ, thetas, lvec: QiskitMPS, vh_phi: QiskitMPS, trunc_thr: float = CHOP,
               chi_max: Optional[int] = None, block_range=None, front_layer: bool = True) -> np.ndarray:
    """Gate-by-gate MPS gradient sweep (fast_dot_gradient, mps_dot_objective.py:41-242)."""
    n, nb = circ.num_qubits, circ.num_blocks
    tpb = 5 if circ.entangler == "cp" else 4
    th1, th2 = thetas[: 3 * n].reshape(n, 3), thetas[3 * n :].reshape(nb, tpb)
    make_rs, pauli_s = O._swappable(circ)
    trot = O._is_trotter(circ)
    lo_b, hi_b = (0, nb) if block_range is None else block_range
    w, z = copy_mps(lvec), copy_mps(vh_phi)
    grad = np.zeros(thetas.size, dtype=C128)
    g1, g2 = grad[: 3 * n].reshape(n, 3), grad[3 * n :].reshape(nb, tpb)

    def rot(q, gate, pauli):
        apply_1q(w, q, gate)
        apply_1q(z, q, gate)
        return _pauli_dot(w, z, q, pauli)

    for q in range(n):
        d2 = rot(q, O.rz(th1[q, 2]), O.PAULI_Z)
        d1 = rot(q, O.ry(th1[q, 1]), O.PAULI_Y)
        d0 = rot(q, O.rz(th1[q, 0]), O.PAULI_Z)
        if front_layer:
            g1[q] = d0, d1, d2
    for i in range(_total_blocks(circ)):
        k = i % nb
        c, t = int(circ.blocks[0, k]), int(circ.blocks[1, k])
        tht = th2[k]
        rec = lo_b <= k < hi_b
        if trot and i % 3 == 0:
            apply_1q(w, c, O.rz(-np.pi / 2))
            apply_1q(z, c, O.rz(-np.pi / 2))
        angle = tht[4] if tpb == 5 else 0.0
        site, e4 = _entangler(circ, c, t, angle)
        apply_2q(z, site, e4, trunc_thr, chi_max)
        if tpb == 5 and rec:
            _, e4s = _entangler(circ, c, t, angle + np.pi)
            w2 = apply_2q(copy_mps(w), site, e4s, trunc_thr, chi_max)
            apply_2q(w, site, e4, trunc_thr, chi_max)
            g2[k, 4] += -0.5j * (mps_dot(w, z) - mps_dot(w2, z))  # :186-196
        else:
            apply_2q(w, site, e4, trunc_thr, chi_max)
        d = [rot(c, O.ry(tht[0]), O.PAULI_Y), rot(c, O.rz(tht[1]), O.PAULI_Z),
             rot(t, O.ry(tht[2]), O.PAULI_Y), rot(t, make_rs(tht[3]), pauli_s)]
        if rec:
            g2[k, 0:4] += d
        if trot and i % 3 == 2:
            apply_1q(w, t, O.rz(np.pi / 2))
            apply_1q(z, t, O.rz(np.pi / 2))
    return grad
