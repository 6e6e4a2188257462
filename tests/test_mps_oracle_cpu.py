"""
Pins the MPS oracle (oracle/mps_oracle.py) and the host-side format helpers
(aqc_research_b200/mps_operations.py) to the reference's own pure-NumPy MPS functions
(golden vectors: tests/golden/mps_cases.npz), and the oracle's gate-by-gate TEBD restatement to
the state-vector oracle in the untruncated limit (what the reference's test_mps.py asserts).
"""

import numpy as np
import pytest

from golden_util import load, mps_from_golden, rel
from aqc_research_b200 import circuit_structures as cs
from aqc_research_b200 import mps_operations as mpsop
from aqc_research_b200.mps_engine import pack_mps, unpack_mps
from aqc_research_b200.parametric_circuit import ParametricCircuit, TrotterAnsatz
from oracle import mps_oracle as M
from oracle import sv_oracle as O

TOL = 1e-12


def test_format_and_dot_golden():
    g = load("mps_cases.npz")
    for c in range(int(g["num_cases"])):
        p = f"m{c}_"
        a, b = mps_from_golden(g, p, "a"), mps_from_golden(g, p, "b")
        assert mpsop.check_mps(a) and mpsop.check_mps(b)
        assert rel(M.mps_to_vector(a), g[p + "vec_a"]) < TOL
        assert rel(mpsop.mps_to_vector(b), g[p + "vec_b"]) < TOL
        assert abs(M.mps_dot(a, b) - complex(g[p + "dot_ab"])) < TOL
        assert abs(M.mps_dot(a, a) - complex(g[p + "dot_aa"])) < TOL
        assert abs(np.vdot(g[p + "vec_a"], g[p + "vec_b"]) - complex(g[p + "dot_ab"])) < TOL
        # device layout round trip
        back = unpack_mps(*pack_mps(a, 64))
        assert rel(M.mps_to_vector(back), g[p + "vec_a"]) < TOL


def test_vector_mps_roundtrip_and_check():
    rng = np.random.RandomState(3)
    for n in (2, 4, 7):
        v = rng.randn(2**n) + 1j * rng.randn(2**n)
        v /= np.linalg.norm(v)
        m = M.vector_to_mps(v)
        assert mpsop.check_mps(m)
        assert rel(M.mps_to_vector(m), v) < TOL
    assert not mpsop.check_mps(([(np.zeros((1, 1)), np.zeros((1, 2)))], []))
    e = M.mps_to_vector(M.product_state(5, 0b10110))
    assert e[0b10110] == 1 and np.count_nonzero(e) == 1


@pytest.mark.parametrize("n", [3, 5])
def test_tebd_restatement_untruncated(n):
    rng = np.random.RandomState(n)
    circs = [TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)]
    circs += [ParametricCircuit(n, e, cs.create_ansatz_structure(n, "spin", "full", 7)) for e in ("cz", "cp")]
    for circ in circs:
        th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
        y = rng.randn(2**n) + 1j * rng.randn(2**n)
        y /= np.linalg.norm(y)
        my = M.vector_to_mps(y)
        assert rel(M.mps_to_vector(M.apply_v(circ, th, my)), O.apply_v(circ, th, y)) < 1e-11
        z0 = O.apply_v(circ, th, y, dagger=True)
        assert rel(M.mps_to_vector(M.apply_v(circ, th, my, dagger=True)), z0) < 1e-11
        g = M.grad_sweep(circ, th, M.product_state(n, 1), M.vector_to_mps(z0))
        e = np.zeros(2**n, dtype=complex)
        e[1] = 1
        assert rel(g, O.grad_sweep(circ, th, e, z0)) < 1e-11


def test_truncation_rule():
    s = np.array([0.9, 0.4, 0.1, 1e-3, 1e-5, 1e-18])
    s /= np.linalg.norm(s)
    keep, kept = M.truncate_rule(s, 1e-16, None)
    assert keep == 5 and np.allclose(kept, s[:5])
    keep, kept = M.truncate_rule(s, 1e-5, None)  # drops 1e-5 and 1e-3 (sum of squares ~1e-6)
    assert keep == 3 and abs(np.linalg.norm(kept) - 1) < 1e-15
    keep, kept = M.truncate_rule(s, 1e-16, 2)
    assert keep == 2 and abs(np.linalg.norm(kept) - 1) < 1e-15
