"""
Scheduler tests (CPU): the tile-pass program compiled by libaqc_b200.so, replayed gate by gate
with the oracle's primitives, must reproduce the oracle's V, V^H and gradient for every
circuit family, tile size and low-bit setting.
"""

import numpy as np
import pytest

from aqc_research_b200 import circuit_structures as cs
from aqc_research_b200 import utils
from aqc_research_b200.engine import CircuitHandle
from aqc_research_b200.parametric_circuit import ParametricCircuit, TrotterAnsatz
from oracle import sv_oracle as O
from program_sim import check_structure, parse_program, replay

TOL = 1e-12


def _rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300)


def _circuits(n):
    yield "t1", TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), False)
    yield "t2", TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    for ent in ("cx", "cz", "cp"):
        yield ent, ParametricCircuit(n, ent, utils.rand_circuit(n, 9))
    yield "spin3", ParametricCircuit(n, "cx", cs.create_ansatz_structure(n, "spin", "full", 7, 3))


@pytest.mark.parametrize("n", [2, 3, 5, 6])
@pytest.mark.parametrize("tb,low", [(2, 0), (3, 1), (4, 2), (11, 4)])
def test_program_replay_matches_oracle(n, tb, low):
    np.random.seed(1000 + 10 * n + tb)
    for name, circ in _circuits(n):
        h = CircuitHandle(circ)
        th = utils.rand_thetas(circ.num_thetas)
        x, y = utils.rand_state(n), utils.rand_state(n)
        for rev in (False, True):
            prog = parse_program(h.debug_program(0, tb, low, rev))
            check_structure(prog, n)
            (v,), _ = replay(prog, circ.entangler, th, [y], dagger=rev, grad=False)
            assert _rel(v, O.apply_v(circ, th, y, dagger=rev)) < TOL, (name, rev)
        z0 = O.apply_v(circ, th, y, dagger=True)
        prog = parse_program(h.debug_program(0, tb, low, False))
        (w, z), g = replay(prog, circ.entangler, th, [x, z0], dagger=False, grad=True)
        assert _rel(g, O.grad_sweep(circ, th, x, z0)) < TOL, name
        assert _rel(w, O.apply_v(circ, th, x)) < TOL and _rel(z, y) < 1e-10


@pytest.mark.parametrize("n,k", [(3, 2), (4, 4)])
def test_program_matrix_layout(n, k):
    """Matrix path: gates act on bits k..k+n-1 of the flat (2^n, 2^k) index."""
    np.random.seed(77 + n)
    m = 1 << k
    for ent in ("cx", "cz", "cp"):
        circ = ParametricCircuit(n, ent, utils.rand_circuit(n, 8))
        h = CircuitHandle(circ)
        th = utils.rand_thetas(circ.num_thetas)
        X = np.random.rand(2**n, m) + 1j * np.random.rand(2**n, m)
        Y = np.random.rand(2**n, m) + 1j * np.random.rand(2**n, m)
        z0 = O.apply_v(circ, th, Y.ravel(), dagger=True, ncols=m)
        for tb, low in ((4, 2), (5, 1), (11, 4)):
            prog = parse_program(h.debug_program(k, tb, low, False))
            check_structure(prog, n + k)
            _, g = replay(prog, ent, th, [X.ravel(), z0], dagger=False, grad=True)
            assert _rel(g, O.grad_sweep(circ, th, X.ravel(), z0, ncols=m)) < TOL


def test_pass_counts_large():
    """Pass counts at benchmark sizes stay far below pair-run granularity."""
    n, layers = 28, 4
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, layers), True)
    h = CircuitHandle(circ)
    prog = parse_program(h.debug_program(0, 11, 4, False))
    check_structure(prog, n)
    pair_runs = (n - 1) * layers + n // 2
    units = sum(len(u) for ps in prog for _, _, u in ps["stages"])
    assert units == n + circ.num_blocks + circ.half_layer_num_blocks
    assert len(prog) < pair_runs / 3
