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
