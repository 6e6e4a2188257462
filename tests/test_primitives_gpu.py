"""
Gate-by-gate primitives through the C-ABI against golden outputs of
the UNMODIFIED reference (tests/golden/primitive_cases.npz, made by make_golden.py primitive_cases):
every function of core_operations.py:34-603 and core_op_matrix.py:32-477 at every qubit position,
vectors and matrices (square and ragged, m = 5 columns).  Tolerance 1e-12 relative (element-wise
complex arithmetic; only FMA contraction differs).
"""

import numpy as np
import pytest

from golden_util import load, rel
from aqc_research_b200 import core_op_matrix as cpm
from aqc_research_b200 import core_operations as cop

pytestmark = pytest.mark.gpu
TOL = 1e-12


def test_vector_primitives_match_the_reference():
    g = load("primitive_cases.npz")
    n, gate, ang = int(g["n"]), g["gate"], float(g["angle"])
    v0, z0 = g["vec"], g["zvec"]
    tmp = np.zeros_like(v0)
    assert cop.bit2bit_transform(n, 1) == n - 2
    for pos in range(n):
        v = v0.copy()
        assert cop.gate2x2_mul_vec(n, pos, gate, v, tmp.copy(), True) is v
        assert rel(v, g[f"gate2x2_{pos}"]) < TOL
        v, o = v0.copy(), np.zeros_like(v0)
        assert cop.gate2x2_mul_vec(n, pos, gate, v, o, False) is o
        assert rel(o, g[f"gate2x2_out_{pos}"]) < TOL and np.array_equal(v, v0)
        assert rel(cop.proj00_mul_vec(n, pos, v0.copy()), g[f"proj00_{pos}"]) < TOL
        assert rel(cop.proj11_mul_vec(n, pos, v0.copy()), g[f"proj11_{pos}"]) < TOL
        for nm in ("rx", "ry", "rz"):
            v = v0.copy()
            assert getattr(cop, nm + "_mul_vec")(n, pos, ang, v, tmp.copy()) is v
            assert rel(v, g[f"{nm}_{pos}"]) < TOL, (nm, pos)
        for nm in ("dot_x", "dot_y", "dot_z"):
            w, z = v0.copy(), z0.copy()
            val = getattr(cop, nm)(n, pos, w, z, tmp.copy())
            assert abs(val - complex(g[f"{nm}_{pos}"])) < TOL * np.linalg.norm(v0) * np.linalg.norm(z0)
            assert np.array_equal(w, v0) and np.array_equal(z, z0)
    for c in range(n):
        for t in range(n):
            if c == t:
                continue
            for nm in ("cx", "cz", "cp"):
                v = v0.copy()
                assert getattr(cop, nm + "_mul_vec")(n, c, t, ang, v, tmp.copy()) is v
                assert rel(v, g[f"{nm}_{c}{t}"]) < TOL, (nm, c, t)
            v, o = v0.copy(), np.zeros_like(v0)
            assert cop.derv_cphase_mul_vec(n, c, t, ang, v, o) is o
            assert rel(o, g[f"dcp_{c}{t}"]) < TOL and np.array_equal(v, v0)
            for dag in (False, True):
                v = v0.copy()
                ws = np.zeros((2, v.size), dtype=np.complex128)
                cop.block_mul_vec(n, c, t, g["c_mat"], g["t_mat"], g["g_mat"], v, ws, dag)
                assert rel(v, g[f"block_{c}{t}_{int(dag)}"]) < TOL, (c, t, dag)


def test_projector_does_not_spread_non_finite_values():
    """Zero gate entries are skipped (core_operations.py:76-119): inf in the discarded half stays out."""
    n = 3
    v = np.ones(2**n, dtype=np.complex128)
    v[1 << (n - 1)] = np.inf  # qubit 0 = 1 half
    out = cop.proj00_mul_vec(n, 0, v)
    assert np.all(np.isfinite(out)) and np.all(out[: 2 ** (n - 1)] == 1) and np.all(out[2 ** (n - 1):] == 0)


def test_degenerate_gates_and_blocks_vs_kronecker():
    """
    gate2x2_mul_vec with every zero pattern of the 2x2 gate (the reference special-cases them,
    core_operations.py:76-119; its test_core_operations.py:124-196) and block_mul_vec for all
    (ctrl, targ) against the dense Kronecker matrices of elementary_operations (qubit 0 leftmost).
    """
    from aqc_research_b200 import elementary_operations as eo

    rng = np.random.RandomState(5)
    n = 3
    a, b, c, d = (rng.randn(4) + 1j * rng.randn(4))
    patterns = [[[a, 0], [0, d]], [[0, b], [c, 0]], [[a, b], [0, 0]], [[0, 0], [c, d]],
                [[a, 0], [c, 0]], [[0, b], [0, d]], [[a, 0], [0, 0]], [[0, 0], [0, d]], [[0, 0], [0, 0]]]
    v0 = rng.randn(2**n) + 1j * rng.randn(2**n)
    for pat in patterns:
        g = np.array(pat, dtype=np.complex128)
        for pos in range(n):
            want = eo._embed(n, {pos: g}) @ v0
            got = cop.gate2x2_mul_vec(n, pos, g, v0.copy(), np.zeros_like(v0), True)
            assert np.linalg.norm(got - want) <= TOL * np.linalg.norm(v0), (pat, pos)
    cm, tm, gm = [rng.randn(2, 2) + 1j * rng.randn(2, 2) for _ in range(3)]
    for ctrl in range(n):
        for targ in range(n):
            if ctrl == targ:
                continue
            ws = np.zeros((2, v0.size), dtype=np.complex128)
            want = eo.np_block_matrix(n, ctrl, targ, cm, tm, gm) @ v0
            assert rel(cop.block_mul_vec(n, ctrl, targ, cm, tm, gm, v0.copy(), ws, False), want) < TOL
            assert rel(cop.cx_mul_vec(n, ctrl, targ, 0.0, v0.copy(), ws[0]), eo.np_cx_matrix(n, ctrl, targ) @ v0) < TOL


@pytest.mark.parametrize("m", [16, 5])
def test_matrix_primitives_match_the_reference(m):
    g = load("primitive_cases.npz")
    n, gate, ang = int(g["n"]), g["gate"], float(g["angle"])
    m0, zm = g[f"mat_{m}"], g[f"zmat_{m}"]
    ws = np.zeros(m0.size, dtype=np.complex128)
    scale = np.linalg.norm(m0) * np.linalg.norm(zm)
    for q in range(n):
        a = m0.copy()
        assert cpm.gate2x2_mul_mat(q, gate, a, ws) is a
        assert rel(a, g[f"m{m}_gate2x2_{q}"]) < TOL
        for nm in ("rx", "ry", "rz"):
            a = m0.copy()
            getattr(cpm, nm + "_mul_mat")(ang, q, a, ws)
            assert rel(a, g[f"m{m}_{nm}_{q}"]) < TOL, (nm, q)
        for nm in ("x", "y", "z"):
            val = getattr(cpm, nm + "_dot_mat")(q, m0.copy(), zm.copy(), ws)
            assert abs(val - complex(g[f"m{m}_{nm}dot_{q}"])) < TOL * scale, (nm, q)
    for c in range(n):
        for t in range(n):
            if c == t:
                continue
            for nm in ("cx", "cz", "cp"):
                a = m0.copy()
                getattr(cpm, nm + "_mul_mat")(c, t, ang, a, ws)
                assert rel(a, g[f"m{m}_{nm}_{c}{t}"]) < TOL, (nm, c, t)
            val = cpm.derv_cphase(c, t, m0.copy(), zm.copy(), ws)
            assert abs(val - complex(g[f"m{m}_dcp_{c}{t}"])) < TOL * scale


def test_bad_arguments_raise():
    v = np.zeros(8, dtype=np.complex128)
    with pytest.raises(AssertionError):
        cop.rx_mul_vec(3, 3, 0.1, v, v.copy())
    with pytest.raises(AssertionError):
        cop.cx_mul_vec(3, 1, 1, 0.0, v, v.copy())
    with pytest.raises(AssertionError):
        cop.gate2x2_mul_vec(3, 0, np.eye(2, dtype=np.complex128), v, v, True)  # aliasing
