"""
Pins the CPU oracle (oracle/sv_oracle.py) to the reference: golden vectors produced by the
unmodified reference (tests/golden/make_golden.py), the SURVEY section 8(c) sanity values, the
identities the reference's own tests assert, and -- when /root/reference is present -- the
reference imported live.
"""

import numpy as np
import pytest

from golden_util import KINDS, circuit_from, load, rel
from oracle import sv_oracle as O
from aqc_research_b200 import circuit_structures as cs, utils
from aqc_research_b200.parametric_circuit import ParametricCircuit, TrotterAnsatz

TOL = 1e-12


def test_sv_golden():
    g = load("sv_cases.npz")
    for c in range(int(g["num_cases"])):
        p = f"c{c}_"
        n, kind, b0, b1 = [int(v) for v in g[p + "meta"]]
        circ = circuit_from(KINDS[kind], n, g[p + "blocks"])
        th, x, y = g[p + "thetas"], g[p + "x"], g[p + "y"]
        assert rel(O.apply_v(circ, th, y), g[p + "v_y"]) < TOL
        assert rel(O.apply_v(circ, th, y, dagger=True), g[p + "vh_y"]) < TOL
        assert rel(O.grad_sweep(circ, th, x, g[p + "vh_y"]), g[p + "grad"]) < TOL
        assert rel(O.grad_sweep(circ, th, x, g[p + "vh_y"], (b0, b1), False), g[p + "grad_part"]) < TOL


def test_matrix_golden():
    g = load("mat_cases.npz")
    for c in range(int(g["num_cases"])):
        p = f"c{c}_"
        n, kind, m = [int(v) for v in g[p + "meta"]]
        circ = circuit_from(KINDS[kind], n, g[p + "blocks"])
        th, X, Y = g[p + "thetas"], g[p + "x"], g[p + "y"]
        assert rel(O.apply_v(circ, th, Y.ravel(), False, m), g[p + "v_y"]) < TOL
        assert rel(O.apply_v(circ, th, Y.ravel(), True, m), g[p + "vh_y"]) < TOL
        assert rel(O.grad_sweep(circ, th, X.ravel(), g[p + "vh_y"].ravel(), ncols=m), g[p + "grad"]) < TOL


def test_objective_sequences_golden():
    g = load("objective_sequences.npz")
    for c in range(int(g["num_sp"])):
        p = f"sp{c}_"
        n, layers, steps = [int(v) for v in g[p + "meta"]]
        if n > 8:
            steps = 1  # keep the CPU suite fast
        circ = TrotterAnsatz(n, g[p + "blocks"], True)
        weight = 1.0
        for s in range(steps):
            th = g[p + "thetas"][s]
            f, hs, grad, _ = O.sur_max_value_and_grad(circ, th, g[p + "target"], weight, int(g[p + "max_no"][s]))
            assert rel(hs, g[p + "hs"][s]) < TOL
            assert abs(f - g[p + "f"][s]) < 1e-12
            assert rel(grad, g[p + "grad"][s]) < 1e-11
            weight = float(g[p + "weight"][s])
    for c in range(int(g["num_sk"])):
        p = f"sk{c}_"
        n, kind = [int(v) for v in g[p + "meta"]]
        circ = ParametricCircuit(n, KINDS[kind], g[p + "blocks"])
        for s in range(g[p + "thetas"].shape[0]):
            f, grad = O.sketch_full_value_and_grad(circ, g[p + "thetas"][s], g[p + "target"])
            assert abs(f - g[p + "f"][s]) < 1e-12
            assert rel(grad, g[p + "grad"][s]) < 1e-11


def test_survey_sanity_values():
    """probe values of the unmodified reference, seed 1234, n = 5."""
    np.random.seed(1234)
    n = 5
    target = utils.rand_state(n)
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    th = utils.rand_thetas(circ.num_thetas)
    f, hs, grad, (g0, _) = O.sur_max_value_and_grad(circ, th, target, 1.0, 0)
    assert abs(f - 0.88712450260286224) < 1e-13
    assert abs(hs[0] - (0.09920365985839553 - 0.32098930086194066j)) < 1e-13
    assert abs(np.linalg.norm(grad) - 0.34811735864678156) < 1e-12
    assert abs(g0[0] - (-0.12242549036285641 - 0.04724294087058167j)) < 1e-13


@pytest.mark.parametrize("n", [2, 3, 4])
def test_identities_of_reference_tests(n):
    """V V^H v = v and apply-by-columns == Kronecker matrix (test_core_operations.py:252-321)."""
    np.random.seed(n)
    circs = [
        TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True),
        TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 1), False),
    ] + [ParametricCircuit(n, e, utils.rand_circuit(n, 6)) for e in ("cx", "cz", "cp")]
    for circ in circs:
        th = utils.rand_thetas(circ.num_thetas)
        v = utils.rand_state(n)
        assert rel(O.apply_v(circ, th, O.apply_v(circ, th, v, dagger=True)), v) < TOL
        U = O.dense_unitary(circ, th)
        assert rel(U, O.dense_unitary_kron(circ, th)) < TOL
        assert rel(U.conj().T @ U, np.eye(2**n)) < TOL
        # gradient vs central finite differences of <V x|y> (utils_dot_gradient_test.py:166-238)
        x, y = utils.rand_state(n), utils.rand_state(n)
        g = O.grad_sweep(circ, th, x, O.apply_v(circ, th, y, dagger=True))
        k = np.random.randint(th.size)
        e = np.zeros_like(th)
        e[k] = 1e-5
        fd = (np.vdot(O.apply_v(circ, th + e, x), y) - np.vdot(O.apply_v(circ, th - e, x), y)) / 2e-5
        assert abs(fd - g[k]) < 1e-8


def test_live_reference_if_present():
    """Oracle vs the reference imported live (build container only)."""
    from ref_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("/root/reference not present")
    R = load_reference()
    np.random.seed(3)
    for n in (3, 5):
        blocks = R.cs.make_trotter_like_circuit(n, 2)
        rc = R.pc.TrotterAnsatz(n, blocks, True)
        th, x, y = R.utils.rand_thetas(rc.num_thetas), R.utils.rand_state(n), R.utils.rand_state(n)
        ws = np.zeros((3, 2**n), dtype=np.complex128)
        out = np.zeros(2**n, dtype=np.complex128)
        z0 = R.cop.v_dagger_mul_vec(rc, th, y, out, ws).copy()
        circ = TrotterAnsatz(n, blocks, True)
        assert rel(O.apply_v(circ, th, y, dagger=True), z0) < TOL
        assert rel(O.grad_sweep(circ, th, x, z0), R.cop.grad_of_dot_product(rc, th, x, z0, ws)) < TOL


def test_c_oracle_golden():
    """The C/OpenMP restatement against the reference's golden vectors and the NumPy oracle."""
    from oracle import c_oracle as C

    g = load("sv_cases.npz")
    for c in range(int(g["num_cases"])):
        p = f"c{c}_"
        n, kind, b0, b1 = [int(v) for v in g[p + "meta"]]
        circ = circuit_from(KINDS[kind], n, g[p + "blocks"])
        th, x, y = g[p + "thetas"], g[p + "x"], g[p + "y"]
        assert rel(C.apply_v(circ, th, y), g[p + "v_y"]) < TOL
        assert rel(C.apply_v(circ, th, y, dagger=True), g[p + "vh_y"]) < TOL
        grad, w, z = C.grad_sweep(circ, th, x, g[p + "vh_y"])
        assert rel(grad, g[p + "grad"]) < TOL
        assert rel(z, y) < 1e-11
    g = load("mat_cases.npz")
    for c in range(int(g["num_cases"])):
        p = f"c{c}_"
        n, kind, m = [int(v) for v in g[p + "meta"]]
        if m & (m - 1):
            continue  # the C oracle takes power-of-two column counts only
        circ = circuit_from(KINDS[kind], n, g[p + "blocks"])
        k = m.bit_length() - 1
        th, X, Y = g[p + "thetas"], g[p + "x"], g[p + "y"]
        assert rel(C.apply_v(circ, th, Y.ravel(), True, k), g[p + "vh_y"]) < TOL
        grad, _, _ = C.grad_sweep(circ, th, X.ravel(), g[p + "vh_y"].ravel(), k)
        assert rel(grad, g[p + "grad"]) < TOL
    # mid size: C oracle == NumPy oracle
    np.random.seed(14)
    n = 14
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 1), True)
    th, x, y = utils.rand_thetas(circ.num_thetas), utils.rand_state(n), utils.rand_state(n)
    z0 = C.apply_v(circ, th, y, dagger=True)
    assert rel(z0, O.apply_v(circ, th, y, dagger=True)) < TOL
    assert rel(C.grad_sweep(circ, th, x, z0)[0], O.grad_sweep(circ, th, x, z0)) < TOL


def test_coord_descent_golden():
    """oracle coord_descent_sweep == reference coord_descent_single_sweep over 4 consecutive sweeps."""
    g = load("cd_cases.npz")
    for c in range(int(g["num_cases"])):
        p = f"c{c}_"
        n, kind = [int(v) for v in g[p + "meta"]]
        circ = circuit_from(KINDS[kind], n, g[p + "blocks"])
        ths, fs = g[p + "thetas"], g[p + "fobj"]
        for s in range(len(fs)):
            f, th_new = O.coord_descent_sweep(circ, ths[s], g[p + "target"])
            assert abs(f - fs[s]) < 1e-11, (c, s, f, fs[s])
            assert rel(th_new, ths[s + 1]) < 1e-11, (c, s)


def test_sketching_generators_golden():
    """SketchOracle (same global-RNG call order) == reference generators + SketchingObjectiveEx."""
    g = load("sketch_cases.npz")
    for c in range(int(g["num_cases"])):
        p = f"c{c}_"
        n, m, kind, ent, seed = [int(v) for v in g[p + "meta"]]
        circ = ParametricCircuit(n, ["cx", "cz", "cp"][ent], g[p + "blocks"])
        np.random.seed(seed)
        orc = O.SketchOracle(["rand", "alt", "eigen"][kind], m, g[p + "target"])
        for s, th in enumerate(g[p + "thetas"]):
            f, grad = orc.value_and_grad(circ, th)
            assert abs(f - g[p + "f"][s]) < 1e-12, (c, s)
            assert rel(grad, g[p + "grad"][s]) < 1e-11, (c, s)
