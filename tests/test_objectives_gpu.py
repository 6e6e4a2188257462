"""
GPU parity of the drop-in objective classes against golden call sequences recorded from the
unmodified reference (tests/golden/objective_sequences.npz): same thetas in the same order must
give the same objective values, gradients, surrogate weight and leading flip-state.
"""

import numpy as np
import pytest

from golden_util import KINDS, load, rel
from aqc_research_b200.model_sketching.sk_core import (
    BatchedSketchingObjective,
    FullRangeSketchingVectors,
    SketchingObjectiveEx,
)
from aqc_research_b200.model_sp_lhs.objective_lhs_sur_max import SpSurrogateObjectiveMax
from aqc_research_b200.parametric_circuit import ParametricCircuit, TrotterAnsatz
from oracle import sv_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _params(n, **kw):
    p = dict(num_qubits=n, max_flips=1, maxiter=10, verbose=0, enable_optim_stats=True,
             num_simulations=1, trunc_thr=1e-6, state_prep_func=None)
    p.update(kw)
    return p


def test_sur_max_sequences():
    g = load("objective_sequences.npz")
    for c in range(int(g["num_sp"])):
        p = f"sp{c}_"
        n, layers, steps = [int(v) for v in g[p + "meta"]]
        circ = TrotterAnsatz(n, g[p + "blocks"], True)
        objv = SpSurrogateObjectiveMax(user_parameters=_params(n), circ=circ, front_layer=True)
        objv.set_target(g[p + "target"])
        for s in range(steps):
            th = g[p + "thetas"][s]
            f = objv.objective(th)
            assert abs(f - g[p + "f"][s]) < TOL
            assert objv.max_no == int(g[p + "max_no"][s])
            grad = objv.gradient(th)
            assert rel(grad, g[p + "grad"][s]) < TOL
            assert abs(objv.weight - g[p + "weight"][s]) < TOL
        assert objv.statistics["num_grad_ev"] == steps


def test_sur_max_gradient_before_objective_and_partial_range():
    """gradient() first must recompute the objective (objective_base.py:715-734)."""
    g = load("objective_sequences.npz")
    p = "sp0_"
    n = int(g[p + "meta"][0])
    circ = TrotterAnsatz(n, g[p + "blocks"], True)
    objv = SpSurrogateObjectiveMax(user_parameters=_params(n), circ=circ, front_layer=True)
    objv.set_target(g[p + "target"])
    grad = objv.gradient(g[p + "thetas"][0])
    assert rel(grad, g[p + "grad"][0]) < TOL
    # partial block range without front layer: entries outside are exactly zero
    objv2 = SpSurrogateObjectiveMax(user_parameters=_params(n), circ=circ, block_range=(3, 9), front_layer=False)
    objv2.set_target(g[p + "target"])
    g2 = objv2.gradient(g[p + "thetas"][0])
    full = g[p + "grad"][0]
    lo, hi = 3 * n + 4 * 3, 3 * n + 4 * 9
    assert np.all(g2[:lo] == 0) and np.all(g2[hi:] == 0)
    assert rel(g2[lo:hi], full[lo:hi]) < TOL


def test_sur_max_neel_and_dense_handlers():
    """Basis-state preparation (Neel) and dense-state handler agree with the oracle."""
    n = 6
    np.random.seed(61)
    from aqc_research_b200 import circuit_structures as cs, utils

    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    target, th = utils.rand_state(n), utils.rand_thetas(circ.num_thetas)
    neel = sum(1 << q for q in range(0, n, 2))
    f_ref, hs_ref, grad_ref, _ = O.sur_max_value_and_grad(circ, th, target, 1.0, 0, init_index=neel)
    # in the reference max_no is picked by hysteresis; replay that rule for the oracle
    hs2 = np.abs(hs_ref) ** 2
    mx = 0
    for i in range(n + 1):
        if 1.1 * hs2[mx] < hs2[i]:
            mx = i
    f_ref, _, grad_ref, _ = O.sur_max_value_and_grad(circ, th, target, 1.0, mx, init_index=neel)
    for prep in (lambda nq: [q for q in range(0, nq, 2)], lambda nq: neel):
        objv = SpSurrogateObjectiveMax(user_parameters=_params(n, state_prep_func=prep), circ=circ, front_layer=True)
        objv.set_target(target)
        assert abs(objv.objective(th) - f_ref) < TOL and objv.max_no == mx
        assert rel(objv.gradient(th), grad_ref) < TOL
    dense = np.zeros((n + 1, 2**n), dtype=np.complex128)
    for i, k in enumerate(O.basis_state_indices(n, neel)):
        dense[i, k] = 1
    objv = SpSurrogateObjectiveMax(user_parameters=_params(n, state_prep_func=lambda nq: dense), circ=circ, front_layer=True)
    objv.set_target(target)
    assert abs(objv.objective(th) - f_ref) < TOL
    assert rel(objv.gradient(th), grad_ref) < TOL


def test_sketching_sequences_and_batch():
    g = load("objective_sequences.npz")
    for c in range(int(g["num_sk"])):
        p = f"sk{c}_"
        n, kind = [int(v) for v in g[p + "meta"]]
        circ = ParametricCircuit(n, KINDS[kind], g[p + "blocks"])
        U = g[p + "target"]
        objv = SketchingObjectiveEx(circ, FullRangeSketchingVectors(U), enable_stats=True)
        ths = g[p + "thetas"]
        for s in range(ths.shape[0]):
            f = objv.objective(ths[s])
            grad = objv.gradient(ths[s])
            assert abs(f - g[p + "f"][s]) < TOL
            assert rel(grad, g[p + "grad"][s]) < TOL
        assert objv.num_iterations == ths.shape[0]
        bo = BatchedSketchingObjective(circ, U, batch=ths.shape[0])
        fb, gb = bo.evaluate(ths)
        assert np.max(np.abs(fb - g[p + "f"])) < TOL
        assert rel(gb, g[p + "grad"]) < TOL


def test_early_stop_exception_propagates():
    """Exceptions raised by the status trackers are control flow and must pass through gradient()."""
    g = load("objective_sequences.npz")
    p = "sp2_"
    n = int(g[p + "meta"][0])
    circ = TrotterAnsatz(n, g[p + "blocks"], True)
    objv = SpSurrogateObjectiveMax(user_parameters=_params(n), circ=circ, front_layer=True)
    objv.set_target(g[p + "target"])

    class Stopper:
        def check(self, **kw):
            raise StopIteration(kw["on_stop"](kw["fobj"], kw["thetas"]))

    objv.set_status_trackers(None, Stopper())
    objv.objective(g[p + "thetas"][0])
    with pytest.raises(StopIteration):
        objv.gradient(g[p + "thetas"][0])


def test_lbfgs_run_matches_reference_trajectory():
    """
    Drop-in under SciPy's L-BFGS-B (what optimizer.py:585-590 calls through Qiskit): config C1
    (n = 5, 2nd-order TrotterAnsatz, Neel state, Trotter target, theta_0 = init_ansatz_to_trotter).
    The optimiser must follow the trajectory recorded with the unmodified reference objective.
    """
    from scipy.optimize import minimize
    from aqc_research_b200 import circuit_structures as cs
    from aqc_research_b200.model_sp_lhs.trotter import trotter as trot

    g = load("trotter_lbfgs.npz")
    n, layers, maxiter, nit, nfev = [int(v) for v in g["lb_meta"]]
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, layers), True)
    th0 = trot.init_ansatz_to_trotter(circ, np.zeros(circ.num_thetas), evol_time=float(g["lb_time"]), delta=1.0)
    assert np.array_equal(th0, g["lb_theta0"])
    # the target itself: GPU Trotter evolution with 10x finer steps == the reference's
    target = trot.trotter_state(n, evol_time=float(g["lb_time"]), num_steps=10 * layers, delta=1.0,
                                second_order=True, ini_state=trot.neel_init_state(n))
    assert rel(target, g["lb_target"]) < TOL
    objv = SpSurrogateObjectiveMax(user_parameters=_params(n, state_prep_func=trot.neel_init_state, maxiter=40),
                                   circ=circ, front_layer=True)
    objv.set_target(g["lb_target"])
    res = minimize(fun=objv.objective, x0=th0.copy(), jac=objv.gradient, method="L-BFGS-B",
                   options=dict(maxfun=5 * maxiter, maxiter=maxiter, ftol=10 * np.finfo(float).eps, eps=1e-8))
    assert (res.nit, res.nfev) == (nit, nfev)
    assert abs(res.fun - float(g["lb_fun"])) < 1e-9
    assert rel(res.x, g["lb_x"]) < 1e-7
    assert abs(objv.fidelity - float(g["lb_fidelity"])) < 1e-9


def test_early_gradient_start_and_its_fallbacks():
    """
    After two fun(theta) / jac(theta) pairs the objective starts the gradient sweep before returning
    (aqc_sv_grad_begin / _end).  Results must not depend on it: pairs, a gradient at OTHER angles
    while a sweep is in flight, objective-only stretches (early start switches off), set_target in
    between -- all against the oracle.
    """
    from aqc_research_b200 import circuit_structures as cs
    from aqc_research_b200 import utils

    n = 8
    np.random.seed(88)
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    target = utils.rand_state(n)
    objv = SpSurrogateObjectiveMax(user_parameters=_params(n), circ=circ, front_layer=True)
    objv.set_target(target)

    def check(th, do_obj=True):
        w_before, max_before = objv.weight, objv.max_no
        if do_obj:
            objv.objective(th)
        g = objv.gradient(th)
        f_ref, _, g_ref, _ = O.sur_max_value_and_grad(circ, th, target, w_before, objv.max_no)
        assert rel(g, g_ref) < TOL, (objv._early_on, max_before)

    ths = [utils.rand_thetas(circ.num_thetas) for _ in range(9)]
    for th in ths[:4]:  # the pattern is learnt after two pairs; pairs 3 and 4 use the early sweep
        check(th)
    assert objv._early_on
    objv.objective(ths[4])  # early sweep in flight for ths[4] ...
    check(ths[5], do_obj=False)  # ... but the gradient is asked at other angles
    assert not objv._early_on
    check(ths[6])
    check(ths[6])
    assert objv._early_on
    objv.objective(ths[7])
    objv.objective(ths[8])  # objective-only: the uncollected sweep switches the early start off
    assert not objv._early_on
    check(ths[8], do_obj=False)
    check(ths[0])
    check(ths[1])
    objv.objective(ths[2])  # early sweep in flight, then the target changes
    target = utils.rand_state(n)
    objv.set_target(target)
    check(ths[3])


@pytest.mark.parametrize("n", [3, 6, 11])
def test_one_submission_evaluation_with_leader_zero(n):
    """
    Near target (the leading state stays |s_0>): once the fun / jac pattern is learnt objective() enqueues the
    whole evaluation (aqc_sv_eval_begin) and gradient() collects it; values and gradients equal the oracle's.
    n = 3 runs on the legacy engine, which has no such call (the class falls back to the two-step path).
    """
    from aqc_research_b200 import circuit_structures as cs, utils

    np.random.seed(1300 + n)
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    th_star = utils.rand_thetas(circ.num_thetas)
    e0 = np.zeros(2**n, dtype=np.complex128)
    e0[0] = 1
    target = O.apply_v(circ, th_star, e0)
    objv = SpSurrogateObjectiveMax(user_parameters=_params(n), circ=circ, front_layer=True)
    objv.set_target(target)
    assert objv.workspace.can_eval == (n >= 5)
    for step in range(6):
        th = th_star + 0.02 * (2 * np.random.rand(circ.num_thetas) - 1)
        w_before = objv.weight
        f = objv.objective(th)
        g = objv.gradient(th)
        f_ref, _, g_ref, _ = O.sur_max_value_and_grad(circ, th, target, w_before, 0)
        assert objv.max_no == 0 and abs(f - f_ref) < TOL and rel(g, g_ref) < TOL, step
    assert objv._early_on


def test_set_sparse_and_fused_two_term_sweep():
    """
    aqc_sv_set_sparse writes a few basis amplitudes; one sweep started from the weighted combination
    of two flip states equals the weighted sum of the two single-state gradients (antilinearity of
    <V x|t> in x), which is what the reference adds up (objective_lhs_sur_max.py:150-186).
    """
    from aqc_research_b200 import circuit_structures as cs, utils
    from aqc_research_b200.engine import SvWorkspace

    n = 7
    np.random.seed(707)
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    th, target = utils.rand_thetas(circ.num_thetas), utils.rand_state(n)
    ws = SvWorkspace(circ, num_slots=4)
    a = np.array([0.3 - 0.2j, -1.1 + 0.7j])
    ws.set_sparse(2, [5, 96], a)
    x = ws.download(2)
    ref = np.zeros(2**n, dtype=np.complex128)
    ref[5], ref[96] = a
    assert np.array_equal(x, ref)
    with pytest.raises(Exception):
        ws.set_sparse(2, [2**n], [1.0])
    with pytest.raises(Exception):
        ws.set_sparse(2, list(range(9)), [1.0] * 9)
    ws.upload(0, target)
    ws.apply(th, 0, 1, dagger=True)
    g5 = ws.grad(th, x_basis=5, z0=1, w=2, z=3)[0]
    ws.apply(th, 0, 1, dagger=True)
    g96 = ws.grad(th, x_basis=96, z0=1, w=2, z=3)[0]
    ws.apply(th, 0, 1, dagger=True)
    ws.set_sparse(2, [5, 96], a)
    both = ws.grad(th, x_slot=2, z0=1, w=2, z=3)[0]
    assert rel(both, np.conj(a[0]) * g5 + np.conj(a[1]) * g96) < TOL
    ws.close()
