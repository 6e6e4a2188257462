"""
GPU parity tests of the MPS path against the CPU oracles.
  * format / dot / amplitudes: vs oracle/mps_oracle.py (pinned to the reference's mps_dot,
    mps_to_vector);
  * gate application and gradient WITHOUT truncation (trunc_thr = 1e-16, as in the reference's
    test_mps.py / test_mps_fast_dot_gradient.py): must equal the state-vector oracle to 1e-10;
  * with truncation: the result must stay normalised, respect chi_max and approach the exact
    state as the threshold shrinks (parity with qiskit-aer is unpinned, see oracle header).
"""

import numpy as np
import pytest

from aqc_research_b200 import circuit_structures as cs
from aqc_research_b200.mps_engine import MpsWorkspace
from aqc_research_b200.parametric_circuit import ParametricCircuit, TrotterAnsatz
from oracle import mps_oracle as M
from oracle import sv_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300)


def _rand_vec(n, rng):
    v = rng.randn(2**n) + 1j * rng.randn(2**n)
    return v / np.linalg.norm(v)


def _circuits(n, rng):
    yield "t1", TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), False)
    yield "t2", TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    for ent in ("cx", "cz", "cp"):
        blocks = cs.create_ansatz_structure(n, "spin", "full", 2 * (n - 1) + 1)
        flip = rng.rand(blocks.shape[1]) < 0.5
        blocks[:, flip] = blocks[::-1, flip]
        yield ent, ParametricCircuit(n, ent, blocks)


@pytest.mark.parametrize("n", [2, 3, 5, 8])
def test_roundtrip_dot_amplitudes(n):
    rng = np.random.RandomState(n)
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 1), True)
    ws = MpsWorkspace(circ, num_slots=3)
    u, v = _rand_vec(n, rng), _rand_vec(n, rng)
    mu, mv = M.vector_to_mps(u), M.vector_to_mps(v)
    ws.upload(0, mu)
    ws.upload(1, mv)
    assert _rel(M.mps_to_vector(ws.download(0)), u) < TOL
    assert abs(ws.dot(0, 1) - np.vdot(u, v)) < TOL
    assert abs(ws.dot(1, 1) - 1) < TOL
    idx = rng.randint(0, 2**n, size=7)
    assert _rel(ws.amplitudes(1, idx), v[idx]) < TOL
    ws.set_product(2, 5 % 2**n)
    e = np.zeros(2**n, dtype=complex)
    e[5 % 2**n] = 1
    assert _rel(M.mps_to_vector(ws.download(2)), e) < TOL
    assert abs(ws.dot(2, 1) - v[5 % 2**n]) < TOL
    ws.close()


@pytest.mark.parametrize("n", [2, 3, 4, 6, 8, 12])
def test_untruncated_apply_and_gradient_vs_statevector(n):
    """n = 12: bonds up to 64, i.e. 128 x 128 SVDs shared by a cluster (all register layouts of the sweeps)."""
    rng = np.random.RandomState(100 + n)
    for name, circ in _circuits(n, rng):
        if n == 12 and name not in ("t2", "cx"):
            continue
        th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
        y = _rand_vec(n, rng)
        ws = MpsWorkspace(circ, num_slots=5, chi_max=64, trunc_thr=1e-16)
        ws.upload(0, M.vector_to_mps(y))
        ws.apply(th, 0, 1, dagger=False)
        assert _rel(M.mps_to_vector(ws.download(1)), O.apply_v(circ, th, y)) < TOL, (name, "V")
        idx = O.basis_state_indices(n, 1)
        hs = ws.objective(th, 0, 1, idx)
        z0 = O.apply_v(circ, th, y, dagger=True)
        assert _rel(M.mps_to_vector(ws.download(1)), z0) < TOL, (name, "VH")
        assert _rel(hs, z0[idx]) < TOL
        # gradient from a basis state and from a dense (entangled) state
        g = ws.grad(th, x_basis=int(idx[1]), z0=1, w=2, z=3)
        e = np.zeros(2**n, dtype=complex)
        e[idx[1]] = 1
        assert _rel(g, O.grad_sweep(circ, th, e, z0)) < TOL, (name, "grad basis")
        x = _rand_vec(n, rng)
        ws.upload(4, M.vector_to_mps(x))
        g = ws.grad(th, x_slot=4, z0=1, w=2, z=3)
        assert _rel(g, O.grad_sweep(circ, th, x, z0)) < TOL, (name, "grad dense")
        assert _rel(M.mps_to_vector(ws.download(3)), y) < 1e-9  # z ends as V V^H y
        assert _rel(M.mps_to_vector(ws.download(2)), O.apply_v(circ, th, x)) < 1e-9
        ws.close()


def test_truncation_behaviour():
    n = 10
    rng = np.random.RandomState(7)
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 3), True)
    th = 0.3 * (2 * rng.rand(circ.num_thetas) - 1)
    exact = O.apply_v(circ, th, np.eye(2**n, dtype=complex)[:, 0])
    errs = []
    for thr, chi in ((1e-16, 32), (1e-8, 32), (1e-4, 32), (1e-16, 4)):
        ws = MpsWorkspace(circ, num_slots=2, chi_max=chi, trunc_thr=thr)
        ws.set_product(0, 0)
        ws.apply(th, 0, 1)
        out = ws.download(1)
        assert max(M.bond_dims(out)) <= chi
        # simultaneous truncations leave the Vidal form only approximately canonical
        assert abs(ws.dot(1, 1) - 1) < (1e-10 if (thr < 1e-12 and chi >= 32) else 0.05)
        errs.append(_rel(M.mps_to_vector(out), exact))
        # same algorithmic rule on the CPU (gate by gate) stays close
        ref = M.apply_v(circ, th, M.product_state(n, 0), trunc_thr=thr, chi_max=chi)
        assert _rel(M.mps_to_vector(out), M.mps_to_vector(ref)) < max(50 * errs[-1], 1e-9)
        ws.close()
    assert errs[0] < TOL and errs[0] <= errs[1] * 1.0001 + 1e-12 and errs[1] <= errs[2] + 1e-12


@pytest.mark.parametrize("ent", ["cx", "cz", "cp"])
def test_non_adjacent_blocks_random_layouts(ent):
    """
    Unit-blocks on ANY (ctrl, targ) pair -- the reference's gradient tests use random layouts
    (test_mps_fast_dot_gradient.py:119-153, mps_dot_objective.py:380-468): the engine routes them
    through a swap network.  V|x>, V^H|y>, mps_dot and the whole fast_dot_gradient against the dense
    oracle, untruncated, n <= 6.
    """
    rng = np.random.RandomState(len(ent) + 40)
    for n in (3, 4, 5, 6):
        for trial in range(3):
            nb = int(rng.randint(1, 3 * n))
            blocks = np.zeros((2, nb), dtype=int)
            for i in range(nb):
                blocks[:, i] = rng.choice(n, size=2, replace=False)
            if trial == 0:  # back-to-back blocks on one distant pair: the inner SWAPs cancel
                blocks[:, 0] = (0, n - 1)
                if nb > 1:
                    blocks[:, 1] = (n - 1, 0)
            circ = ParametricCircuit(n, ent, blocks)
            th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
            x, y = _rand_vec(n, rng), _rand_vec(n, rng)
            ws = MpsWorkspace(circ, num_slots=5, chi_max=64, trunc_thr=1e-16)
            ws.upload(0, M.vector_to_mps(y))
            ws.upload(4, M.vector_to_mps(x))
            ws.apply(th, 4, 1)
            assert _rel(M.mps_to_vector(ws.download(1)), O.apply_v(circ, th, x)) < TOL, (n, trial, blocks)
            ws.apply(th, 0, 1, dagger=True)
            z0 = O.apply_v(circ, th, y, dagger=True)
            assert _rel(M.mps_to_vector(ws.download(1)), z0) < TOL
            g = ws.grad(th, x_slot=4, z0=1, w=2, z=3)
            assert _rel(g, O.grad_sweep(circ, th, x, z0)) < TOL, (n, trial, blocks)
            assert abs(ws.dot(2, 3) - np.vdot(O.apply_v(circ, th, x), y)) < TOL
            ws.close()


def test_mps_objective_class_reproduces_statevector_golden():
    """
    SpSurrogateObjectiveFastMpsTrotter without truncation must reproduce the golden call
    sequences recorded from the reference's STATE-VECTOR objective (same surrogate, same states).
    """
    from golden_util import load, rel
    from aqc_research_b200.model_sp_lhs.objective_lhs_sur_fast_mps_trotter import (
        SpSurrogateObjectiveFastMpsTrotter,
    )

    g = load("objective_sequences.npz")
    for c in range(int(g["num_sp"])):
        p = f"sp{c}_"
        n, layers, steps = [int(v) for v in g[p + "meta"]]
        if n > 8:
            continue
        circ = TrotterAnsatz(n, g[p + "blocks"], True)
        params = dict(num_qubits=n, max_flips=1, maxiter=10, verbose=0, enable_optim_stats=False,
                      num_simulations=1, trunc_thr=1e-16, state_prep_func=None)
        objv = SpSurrogateObjectiveFastMpsTrotter(user_parameters=params, circ=circ)
        objv.set_target(M.vector_to_mps(g[p + "target"]))
        for s in range(steps):
            th = g[p + "thetas"][s]
            assert abs(objv.objective(th) - g[p + "f"][s]) < TOL
            assert objv.max_no == int(g[p + "max_no"][s])
            assert rel(objv.gradient(th), g[p + "grad"][s]) < TOL
            assert abs(objv.weight - g[p + "weight"][s]) < TOL


def test_function_level_shims():
    from aqc_research_b200 import mps_operations as mpsop
    from aqc_research_b200.mps_dot_objective import fast_dot_gradient

    rng = np.random.RandomState(12)
    n = 5
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 1), True)
    th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    x, y = _rand_vec(n, rng), _rand_vec(n, rng)
    mx, my = M.vector_to_mps(x), M.vector_to_mps(y)
    assert abs(mpsop.mps_dot(mx, my) - np.vdot(x, y)) < TOL
    vy = mpsop.v_mul_mps(circ, th, my)
    assert _rel(mpsop.mps_to_vector(vy), O.apply_v(circ, th, y)) < TOL
    vhy = mpsop.v_dagger_mul_mps(circ, th, my)
    z0 = O.apply_v(circ, th, y, dagger=True)
    assert _rel(mpsop.mps_to_vector(vhy), z0) < TOL
    g = fast_dot_gradient(circ, th, mx, vhy, block_range=(3, 9), front_layer=False)
    assert _rel(g, O.grad_sweep(circ, th, x, z0, (3, 9), False)) < TOL


def test_full_size_properties_n50():
    """
    BASELINE.json configs[3] at full size (n = 50, chi_max = 64, Trotter ansatz depth 20) through
    size-independent properties: a physical target (Trotter-evolved Neel state, bonds stay small at
    this evolution time, so nothing is truncated), angles near the Trotter point.
      * V^H is norm preserving: <z0|z0> = 1;
      * the gradient sweep carries z0 = V^H y back to V V^H y = y and |0..> to V|neel>: both
        overlaps are checked through MPS dots;
      * the complex gradient of <V x|y> agrees with central finite differences of the overlap.
    """
    from aqc_research_b200.model_sp_lhs.trotter import trotter as trotop

    n, layers = 50, 20
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, layers), True)
    th_t = trotop.init_ansatz_to_trotter(circ, np.zeros(circ.num_thetas), evol_time=1.0, delta=1.0)
    rng = np.random.RandomState(50)
    th = th_t + 0.02 * (2 * rng.rand(circ.num_thetas) - 1)
    ws = MpsWorkspace(circ, num_slots=6, chi_max=64, trunc_thr=1e-16)
    neel = sum(1 << q for q in range(0, n, 2))
    ws.set_product(0, neel)
    ws.apply(th_t, 0, 0)  # target y = Trotter(1.0) |neel>
    assert abs(ws.dot(0, 0) - 1) < 1e-10
    assert max(M.bond_dims(ws.download(0))) < 64  # no bond saturates: the 1e-16 rule drops noise only
    idx = np.array([neel] + [neel ^ (1 << q) for q in range(n)], dtype=np.int64)
    hs = ws.objective(th, 0, 1, idx)  # slot 1 = z0 = V^H y
    assert abs(ws.dot(1, 1) - 1) < 1e-9
    assert abs(np.ravel(hs)[0]) ** 2 > 0.5  # near the Trotter point the fidelity is O(1)
    g = ws.grad(th, x_basis=neel, z0=1, w=2, z=3)  # slot 2 = V x, slot 3 = V V^H y
    assert abs(abs(ws.dot(3, 0)) - 1) < 1e-7
    f0 = ws.dot(2, 0)  # <V x | y>
    assert abs(f0 - np.ravel(hs)[0]) < 1e-7
    # finite differences of f(theta) = <V(theta) x | y> = conj(<y| V(theta) x>) for a few angles
    ws.set_product(4, neel)
    for k in (1, 3 * n + 5, 3 * n + 4 * 137 + 2, circ.num_thetas - 3):
        vals = []
        for sgn in (+1, -1):
            t2 = th.copy()
            t2[k] += sgn * 1e-4
            ws.apply(t2, 4, 5)
            vals.append(ws.dot(5, 0))
        fd = (vals[0] - vals[1]) / 2e-4
        assert abs(fd - g[k]) < 1e-6 * max(1.0, abs(g[k])), (k, fd, g[k])
    ws.close()


def _pair_runs(circ):
    """Upper bound of the SVD splits one sweep of one state makes (one per pair-run of the program)."""
    return circ.num_blocks + getattr(circ, "half_layer_num_blocks", 0)


@pytest.mark.parametrize("n,evol_time", [(12, 2.5), (14, 2.0)])
def test_truncated_results_within_the_discarded_weight_bound(n, evol_time):
    """
    trunc_thr = 1e-6 (the production setting, user_options.py:55; BASELINE config 4) against the EXACT
    state-vector oracle, with a bound stated in the weight the splits discarded:
    a split that drops the weight eps_k moves the (renormalised) state by at most sqrt(2 eps_k), so
    after K splits  ||z_trunc - z_exact|| <= sum_k sqrt(2 eps_k) <= sqrt(2 K W),  W = sum_k eps_k
    (``truncation_stats``), K <= pair-runs of the circuit.  hs_i = <s_i|z0> and every gradient entry
    0.5j <P w|z> inherit the bound (|P| = 1, unit vectors).  The physical target (Trotter-evolved Neel
    state, 10x finer steps) really is truncated at these evolution times: W > 0 is asserted.
    """
    from aqc_research_b200.model_sp_lhs.trotter import trotter as trotop

    rng = np.random.RandomState(n)
    layers = 3
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, layers), True)
    fine = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 10 * layers), True)
    th_f = trotop.init_ansatz_to_trotter(fine, np.zeros(fine.num_thetas), evol_time=evol_time, delta=1.0)
    th = trotop.init_ansatz_to_trotter(circ, np.zeros(circ.num_thetas), evol_time=evol_time, delta=1.0)
    th = th + 0.05 * (2 * rng.rand(circ.num_thetas) - 1)
    neel = sum(1 << q for q in range(0, n, 2))
    e = np.zeros(2**n, dtype=complex)
    e[neel] = 1
    # target: the Trotter-evolved Neel state, compressed once on the host to bonds <= 64; the dense
    # vector of THAT MPS is the exact target of this test
    y = M.mps_to_vector(M.vector_to_mps(O.apply_v(fine, th_f, e), chop=1e-7))
    y /= np.linalg.norm(y)
    mps_y = M.vector_to_mps(y, chop=1e-13)
    assert max(M.bond_dims(mps_y)) <= 64
    ws = MpsWorkspace(circ, num_slots=4, chi_max=64, trunc_thr=1e-6)
    ws.upload(0, mps_y)
    idx = np.array([neel] + [neel ^ (1 << q) for q in range(n)], dtype=np.int64)
    hs = np.ravel(ws.objective(th, 0, 1, idx))
    st_o = ws.truncation_stats()
    z0 = O.apply_v(circ, th, y, dagger=True)
    K = _pair_runs(circ)
    bound_o = np.sqrt(2.0 * K * st_o["discarded_weight"]) + 1e-12
    assert st_o["cap_hits"] == 0 and st_o["max_discarded"] < 1e-6  # the trunc_thr rule alone was at work
    assert np.max(np.abs(hs - z0[idx])) <= bound_o, (np.max(np.abs(hs - z0[idx])), bound_o, st_o)
    # every split renormalises; simultaneous truncations of a half-layer leave the Vidal form canonical only
    # up to the discarded weight, so the norm is 1 to that order, not to rounding
    assert abs(ws.dot(1, 1) - 1) < 10 * st_o["discarded_weight"] + 1e-9
    g = ws.grad(th, x_basis=neel, z0=1, w=2, z=3)
    st_g = ws.truncation_stats()
    g_ref = O.grad_sweep(circ, th, e, z0)
    # z starts from the truncated z0 and both swept states are truncated again
    bound_g = bound_o + np.sqrt(2.0 * 2 * K * st_g["discarded_weight"]) + 1e-12
    assert np.max(np.abs(g - g_ref)) <= bound_g, (np.max(np.abs(g - g_ref)), bound_g, st_g)
    assert st_o["discarded_weight"] + st_g["discarded_weight"] > 0  # the case does truncate
    # the bound is a meaningful one (well below the size of the results, max |g| ~ 0.6) and the actual
    # deviation is what trunc_thr = 1e-6 suggests (the gate-by-gate CPU rule gives ~5e-3 in the state)
    print(f"[mps trunc n={n}] max|dhs|={np.max(np.abs(hs - z0[idx])):.2e} (bound {bound_o:.2e}) "
          f"max|dg|={np.max(np.abs(g - g_ref)):.2e} (bound {bound_g:.2e}) {st_o} {st_g}")
    assert bound_g < 0.5 and np.max(np.abs(g - g_ref)) < 2e-2 and np.max(np.abs(hs - z0[idx])) < 2e-2
    ws.close()


def test_bond_cap_is_reported_not_silent():
    """
    chi_max is a deviation from the reference (qiskit-aer has no cap): when the cap -- not
    trunc_thr -- removes weight, the statistics say so and the no-truncation helpers raise.
    """
    from aqc_research_b200 import mps_operations as mpsop
    from aqc_research_b200.mps_engine import BondCapacityError

    n = 10
    rng = np.random.RandomState(3)
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 3), True)
    th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)  # random angles: bonds grow to 2^5 = 32
    ws = MpsWorkspace(circ, num_slots=2, chi_max=8, trunc_thr=1e-16)
    ws.set_product(0, 0)
    ws.apply(th, 0, 1)
    st = ws.truncation_stats()
    assert st["cap_hits"] > 0 and st["cap_discarded"] > 1e-6 and st["discarded_weight"] >= st["cap_discarded"]
    with pytest.raises(BondCapacityError):
        ws.check_cap("test")
    ws.close()
    ws = MpsWorkspace(circ, num_slots=2, chi_max=32, trunc_thr=1e-16)
    ws.set_product(0, 0)
    ws.apply(th, 0, 1)
    st = ws.truncation_stats()
    assert st["cap_hits"] == 0 and st["discarded_weight"] < 1e-12
    ws.check_cap("test")
    ws.close()
    with pytest.raises(BondCapacityError):
        mpsop.v_mul_mps(circ, th, M.product_state(n, 0), chi_max=8)


def test_saturated_n50_gradient_is_consistent_with_the_truncated_overlap():
    """
    n = 50, chi_max = 64, depth 20, trunc_thr = 1e-6 in the regime where bonds DO saturate (VERDICT r01
    weak #3: the bench workload).  No exact answer exists at this size; what can be checked is that the
    gradient is the derivative of the overlap the same truncated machinery computes: central finite
    differences of hs_0(theta) = <x|V^H(theta) y>_trunc, with a tolerance stated in the discarded weight.
    ||z0|| and the truncation record are printed (bench.py reports them for the timed workload).
    Two regimes: a physical one (Trotter-evolved target, evolution time 3: bonds reach the cap mildly)
    where the agreement is tight, and random angles (every bond saturates, large discarded weight) where
    only the bound -- then O(1) -- and finiteness can be asserted.
    """
    from aqc_research_b200.model_sp_lhs.trotter import trotter as trotop

    n, layers = 50, 20
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, layers), True)
    K = _pair_runs(circ)
    neel = sum(1 << q for q in range(0, n, 2))
    idx = np.array([neel], dtype=np.int64)
    rng = np.random.RandomState(5050)
    for regime in ("physical", "random"):
        if regime == "physical":
            th_t = trotop.init_ansatz_to_trotter(circ, np.zeros(circ.num_thetas), evol_time=3.0, delta=1.0)
            th = th_t + 0.01 * (2 * rng.rand(circ.num_thetas) - 1)
        else:
            th_t = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
            th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
        ws = MpsWorkspace(circ, num_slots=4, chi_max=64, trunc_thr=1e-6)
        ws.set_product(0, neel)
        ws.apply(th_t, 0, 0)  # target
        st_t = ws.truncation_stats()
        hs = np.ravel(ws.objective(th, 0, 1, idx))
        st_o = ws.truncation_stats()
        z0_norm = abs(ws.dot(1, 1))
        g = ws.grad(th, x_basis=neel, z0=1, w=2, z=3)
        st_g = ws.truncation_stats()
        assert np.all(np.isfinite(g)) and np.isfinite(hs[0])
        # every split renormalises its own bond; the half-layer's simultaneous truncations leave the Vidal
        # form canonical only up to the discarded weight, so the norm is 1 to that order: tight in the
        # physical regime, O(1) off when every bond is cut at the cap (the bench workload; reported there)
        assert abs(z0_norm - 1) < (1e-5 if regime == "physical" else 1.0)
        W = st_o["discarded_weight"] + st_g["discarded_weight"]
        bound = np.sqrt(2.0 * 3 * K * W)
        worst = 0.0
        for k in (2, 3 * n + 4 * 57 + 1, circ.num_thetas - 2):
            vals = []
            for sgn in (+1, -1):
                t2 = th.copy()
                t2[k] += sgn * 1e-3
                vals.append(np.ravel(ws.objective(t2, 0, 1, idx))[0])
            fd = (vals[0] - vals[1]) / 2e-3  # d<x|V^H y>/dtheta_k = d<V x|y>/dtheta_k
            worst = max(worst, abs(fd - g[k]))
        print(f"[mps n=50 {regime}] |hs0|^2={abs(hs[0])**2:.4f} ||z0||={z0_norm:.9f} target trunc {st_t} "
              f"objective trunc {st_o} gradient trunc {st_g} max|fd-g|={worst:.3e} bound={bound:.3e}")
        # finite differences of a truncated overlap carry the truncation noise divided by the step
        assert worst <= bound / 1e-3 + 1e-5, (regime, worst, bound)
        if regime == "physical":
            assert abs(hs[0]) ** 2 > 0.3 and worst < 1e-3 * max(1.0, np.max(np.abs(g))), (worst, hs)
        ws.close()


def test_set_product_site_and_fused_two_term_sweep():
    """
    aqc_mps_set_product_site: a product state with one qubit in superposition; one gradient sweep
    from the weighted pair of flip states equals the weighted sum of two sweeps (antilinearity).
    """
    rng = np.random.RandomState(77)
    n = 6
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    ws = MpsWorkspace(circ, num_slots=4, chi_max=64, trunc_thr=1e-16)
    index, site = 0b010101, 2  # bit 2 of the index is set
    a_same, a_flip = np.array([0.6 - 0.3j, -0.2 + 0.7j]) / np.sqrt(0.98)  # unit norm, as the caller keeps it
    ws.set_product_site(2, index, site, a_flip, a_same)  # |0> carries the flipped state here
    vec = M.mps_to_vector(ws.download(2))
    ref = np.zeros(2**n, dtype=np.complex128)
    ref[index], ref[index ^ (1 << site)] = a_same, a_flip
    assert _rel(vec, ref) < 1e-15
    with pytest.raises(Exception):
        ws.set_product_site(2, index, n, 1.0, 0.0)
    ws.upload(0, M.vector_to_mps(_rand_vec(n, rng)))
    ws.apply(th, 0, 1, dagger=True)
    g_same = ws.grad(th, x_basis=index, z0=1, w=2, z=3)
    g_flip = ws.grad(th, x_basis=index ^ (1 << site), z0=1, w=2, z=3)
    ws.set_product_site(2, index, site, a_flip, a_same)
    both = ws.grad(th, x_slot=2, z0=1, w=2, z=3)
    assert _rel(both, np.conj(a_same) * g_same + np.conj(a_flip) * g_flip) < TOL
    ws.close()


def test_rand_mps_vec():
    """rand_mps_vec (mps_operations.py:301-323): a valid, normalised MPS whose dense form is returned."""
    from aqc_research_b200 import mps_operations as mpsop

    np.random.seed(3)
    n = 6
    dense = np.zeros(2**n, dtype=np.complex128)
    mps = mpsop.rand_mps_vec(n, out_state=dense, num_layers=2)
    assert mpsop.check_mps(mps) and len(mps[0]) == n
    assert abs(np.linalg.norm(dense) - 1.0) < 1e-12
    assert _rel(M.mps_to_vector(mps), dense) < 1e-14
    assert abs(mpsop.mps_dot(mps, mps) - 1.0) < 1e-12
    other = mpsop.rand_mps_vec(n)
    assert abs(mpsop.mps_dot(mps, other) - np.vdot(dense, M.mps_to_vector(other))) < 1e-12


def test_gate_by_gate_helpers_vs_dense_oracle():
    """
    x/y/z_mul_mps, rx/ry/rz_mul_mps, cx/cz/cp_mul_mps, dot_x/y/z (mps_dot_objective.py:245-516)
    against the dense gates of the oracle (qubit k = bit k of the flat index, as mps_to_vector).
    """
    from aqc_research_b200 import mps_dot_objective as mdo

    rng = np.random.RandomState(21)
    n, ang = 5, 0.9173
    v, z = _rand_vec(n, rng), _rand_vec(n, rng)
    mv, mz = M.vector_to_mps(v), M.vector_to_mps(z)
    for q in range(n):
        for fn, g in ((mdo.x_mul_mps, O.PAULI_X), (mdo.y_mul_mps, O.PAULI_Y), (mdo.z_mul_mps, O.PAULI_Z)):
            assert _rel(M.mps_to_vector(fn(q, mv)), O.op1(v.copy(), q, g)) < TOL, (fn.__name__, q)
        for fn, mk in ((mdo.rx_mul_mps, O.rx), (mdo.ry_mul_mps, O.ry), (mdo.rz_mul_mps, O.rz)):
            assert _rel(M.mps_to_vector(fn(ang, q, mv)), O.op1(v.copy(), q, mk(ang))) < TOL, (fn.__name__, q)
        for fn, g in ((mdo.dot_x, O.PAULI_X), (mdo.dot_y, O.PAULI_Y), (mdo.dot_z, O.PAULI_Z)):
            assert abs(fn(q, mv, mz) - 0.5j * np.vdot(O.op1(v.copy(), q, g), z)) < TOL, (fn.__name__, q)
    for c, t in ((0, 1), (1, 0), (3, 2), (3, 4)):
        assert _rel(M.mps_to_vector(mdo.cx_mul_mps(0.0, c, t, mv)), O.ctrl_op(v.copy(), c, t, O.PAULI_X)) < TOL
        assert _rel(M.mps_to_vector(mdo.cz_mul_mps(0.0, c, t, mv)), O.ctrl_op(v.copy(), c, t, O.PAULI_Z)) < TOL
        assert _rel(M.mps_to_vector(mdo.cp_mul_mps(ang, c, t, mv)), O.ctrl_op(v.copy(), c, t, O.phase(ang))) < TOL
    for c, t in ((0, 2), (4, 0), (1, 4)):  # non-adjacent pairs go through the engine's swap network
        assert _rel(M.mps_to_vector(mdo.cx_mul_mps(0.0, c, t, mv)), O.ctrl_op(v.copy(), c, t, O.PAULI_X)) < TOL
        assert _rel(M.mps_to_vector(mdo.cp_mul_mps(ang, c, t, mv)), O.ctrl_op(v.copy(), c, t, O.phase(ang))) < TOL
