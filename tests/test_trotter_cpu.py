"""Host-side Trotter helpers against the reference's golden outputs (tests/golden/trotter_lbfgs.npz)."""

import numpy as np

from golden_util import load, rel
from aqc_research_b200 import circuit_structures as cs
from aqc_research_b200.model_sp_lhs.trotter import trotter as trot
from aqc_research_b200.parametric_circuit import TrotterAnsatz
from oracle import sv_oracle as O


def test_init_ansatz_to_trotter_matches_reference():
    g = load("trotter_lbfgs.npz")
    for c in range(int(g["num_ia"])):
        n, layers, so = [int(v) for v in g[f"ia{c}_meta"]]
        circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, layers), bool(so))
        th = trot.init_ansatz_to_trotter(circ, np.full(circ.num_thetas, 0.123), evol_time=float(g[f"ia{c}_time"]), delta=1.0)
        assert np.array_equal(th, g[f"ia{c}_thetas"])
        # the oracle applied to the Neel state reproduces the reference's Trotter-evolved state
        v = np.zeros(2**n, dtype=np.complex128)
        v[trot.basis_index(trot.neel_init_state(n))] = 1
        assert rel(O.apply_v(circ, th, v), g[f"ia{c}_state"]) < 1e-12


def test_partial_layer_range_and_helpers():
    n = 5
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 4), True)
    th = np.full(circ.num_thetas, 0.5)
    trot.init_ansatz_to_trotter(circ, th, evol_time=0.6, delta=1.0, layer_range=(2, 4))
    v2q, rng = trot.slice2q(circ, th)
    assert rng == (0, 4) and np.all(v2q[:2] == 0.5) and np.all(circ.subset1q(th) == 0.5)
    a = trot.trotter_alphas(0.3, 1.0)
    assert np.allclose(v2q[2:, :, 5], a[0]) and np.allclose(v2q[2:, :, 0], a[1]) and np.allclose(v2q[2:, :, 6], a[2])
    assert trot.neel_init_state(5) == [0, 2, 4] and trot.half_zero_circuit(4) == [2, 3]
    x, y = np.array([1, 0], dtype=complex), np.array([0.6, 0.8j])
    assert abs(trot.fidelity(x, y) - 0.36) < 1e-15


def test_hamiltonian_exact_evolution_and_global_phase():
    """
    make_hamiltonian / exact_evolution / trotter_global_phase (trotter.py:183-314 of the reference):
    Hermitian XXZ chain with the published coefficients; the Trotter circuit (= TrotterAnsatz with
    init_ansatz_to_trotter angles, evaluated with the oracle here) converges to exp(-itH)|Neel>, and
    for the first order e^{i phase} closes the gap completely.  When /root/reference is present the
    three functions are also compared with the reference itself.
    """
    n, delta, t = 4, 0.8, 0.7
    ham = trot.make_hamiltonian(n, delta)
    assert np.allclose(ham, ham.conj().T)
    # <01|H|10> = -1/2 on every bond, diagonal = -delta/4 * sum of (+1 equal, -1 different)
    assert ham[0b0001, 0b0010] == -0.5 and ham[0, 0] == -0.25 * delta * (n - 1)
    assert ham[0b0101, 0b0101] == 0.25 * delta * (n - 1)
    neel = trot.basis_index(trot.neel_init_state(n))
    exact = trot.exact_evolution(ham, neel, t)
    assert abs(np.linalg.norm(exact) - 1.0) < 1e-13
    v0 = np.zeros(2**n, dtype=np.complex128)
    v0[neel] = 1
    assert np.allclose(exact, trot.exact_evolution(ham, v0, t))
    errs = []
    for steps in (5, 10, 20):
        circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, steps), False)
        th = trot.init_ansatz_to_trotter(circ, np.zeros(circ.num_thetas), evol_time=t, delta=delta)
        state = np.exp(1j * trot.trotter_global_phase(n, steps, False)) * O.apply_v(circ, th, v0)
        errs.append(np.linalg.norm(state - exact))
    assert errs[0] > errs[1] > errs[2] and errs[2] < 2e-2  # first order: error ~ 1 / steps
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 10), True)
    th = trot.init_ansatz_to_trotter(circ, np.zeros(circ.num_thetas), evol_time=t, delta=delta)
    assert 1.0 - abs(np.vdot(exact, O.apply_v(circ, th, v0))) < 1e-6  # second order: far closer
    assert trot.trotter_global_phase(5, 3, True) == 0.25 * np.pi * (4 * 3 + 4)
    assert trot.trotter_global_phase(4, 3, True) == 0.25 * np.pi * (3 * 3 + 4)

    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from ref_loader import load_reference, reference_available

    if reference_available():
        ref = load_reference().trotter
        for nq in (3, 4, 5):
            assert np.allclose(trot.make_hamiltonian(nq, delta), ref.make_hamiltonian(nq, delta), rtol=0, atol=1e-15)
            for so in (False, True):
                assert trot.trotter_global_phase(nq, 7, so) == ref.trotter_global_phase(nq, 7, so)
        assert np.allclose(exact, ref.exact_evolution(ref.make_hamiltonian(n, delta).astype(np.complex128), v0, t),
                           rtol=0, atol=1e-15)


def test_trotter_class_properties():
    tr = trot.Trotter(num_qubits=5, evol_time=1.2, num_steps=6, delta=0.9, second_order=True)
    assert tr.evol_time == 1.2 and tr.num_trotter_steps == 6 and abs(tr.time_step - 0.2) < 1e-15
    import pytest

    with pytest.raises(NotImplementedError):
        tr.as_qcircuit(None)
    with pytest.raises(AssertionError):
        trot.Trotter(num_qubits=1, evol_time=1.0, num_steps=1, second_order=False)


def test_reference_named_state_handlers_and_predicates():
    """ThinStateHandler / GenericStateHandler / MpsStateHandler (objective_base.py:42-429) and checking.py."""
    from aqc_research_b200 import checking as chk
    from aqc_research_b200.model_sp_lhs import objective_base as ob

    n = 4
    thin = ob.ThinStateHandler(n, 1)
    assert thin.num_states == n + 1 and list(thin.state_indices) == [0, 1, 2, 4, 8]
    rng = np.random.RandomState(3)
    vec = rng.randn(2**n) + 1j * rng.randn(2**n)
    for i in range(thin.num_states):
        st = thin.init_state(i).copy()
        assert st.sum() == 1 and st[thin.state_indices[i]] == 1
        assert thin.state_dot_vector(i, vec) == np.vdot(st, vec)
    coefs = rng.randn(n + 1) + 1j * rng.randn(n + 1)
    coefs /= np.linalg.norm(coefs)
    comp = thin.init_composite_state(coefs).copy()
    assert abs(thin.composite_state_dot_vector(coefs, vec) - np.vdot(comp, vec)) < 1e-14
    c1 = coefs[1:] / np.linalg.norm(coefs[1:])
    comp = thin.init_composite_state_no_zero(c1).copy()
    assert comp[0] == 0 and abs(thin.composite_state_dot_vector_no_zero(c1, vec) - np.vdot(comp, vec)) < 1e-14
    two = ob.ThinStateHandler(n, 2)
    assert two.num_states == 1 + n + n * (n - 1) // 2
    gen = ob.GenericStateHandler(n, 1, trot.neel_init_state)
    neel = trot.basis_index(trot.neel_init_state(n))
    assert gen.num_states == n + 1 and gen.state0[neel] == 1
    assert gen.state_dot_vector(2, vec) == vec[neel ^ 0b10]
    mps = ob.MpsStateHandler(n, 1, trot.neel_init_state)
    gam, lam = mps.init_state(3)
    assert len(gam) == n and len(lam) == n - 1 and list(mps.state_indices) == [neel ^ m for m in (0, 1, 2, 4, 8)]
    bits = [int(abs(g[1][0, 0]) == 1) for g in gam]
    assert sum(b << q for q, b in enumerate(bits)) == neel ^ 0b100
    import pytest

    with pytest.raises(ValueError):
        ob.GenericStateHandler(n, 2, trot.neel_init_state)
    with pytest.raises(NotImplementedError):
        gen.init_composite_state(coefs)
    assert chk.check_permutation(np.array([2, 0, 1])) and not chk.check_permutation(np.array([0, 0, 1]))
    assert chk.int_2d(np.zeros((2, 2), dtype=np.int64)) and chk.bool_1d(np.zeros(3, dtype=bool))
    assert chk.is_complex(1j) and not chk.is_complex(1.0) and chk.none_or_type(None, int)
    a = np.zeros(4, dtype=np.complex128)
    assert chk.check_sim_complex_vecs4(a, a.copy(), a.copy(), a.copy())
    assert not chk.check_sim_complex_vecs4(a, a.copy(), a.copy(), np.zeros(5, dtype=np.complex128))
    assert chk.complex_or_float_1d(np.zeros(3)) and chk.complex_3d(np.zeros((1, 1, 1), dtype=np.complex128))
