"""
Dense-stage (DMMA) engine, CPU side: the program compiled for it (front gates merged, up to 5
units per stage) replayed gate by gate, its shared-memory lane tables, and a NumPy emulation of
the sweep kernel's fragment algebra + the post-processing of the stage matrices, all against
the oracle.
"""

import numpy as np
import pytest

from aqc_research_b200 import circuit_structures as cs
from aqc_research_b200 import utils
from aqc_research_b200.engine import CircuitHandle
from aqc_research_b200.parametric_circuit import ParametricCircuit, TrotterAnsatz
from oracle import sv_oracle as O
from program_sim import (
    check_structure,
    dense_bank_conflicts,
    dense_check_tables,
    dense_emulate,
    parse_program,
    replay,
)

TOL = 1e-12


def _rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300)


def _circuits(n):
    yield "t1", TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), False)
    yield "t2", TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    for ent in ("cx", "cz", "cp"):
        yield ent, ParametricCircuit(n, ent, utils.rand_circuit(n, 9))
    yield "spin3", ParametricCircuit(n, "cx", cs.create_ansatz_structure(n, "spin", "full", 7, 3))


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("n,tb,low", [(5, 5, 2), (6, 5, 1), (6, 6, 4), (7, 6, 3), (8, 11, 4)])
def test_dense_program_and_emulation(n, tb, low, fused):
    """fused = True: the production tables, consecutive stages on disjoint bit pairs run as one step."""
    np.random.seed(4000 + 10 * n + tb)
    for name, circ in _circuits(n):
        h = CircuitHandle(circ)
        th = utils.rand_thetas(circ.num_thetas)
        x, y = utils.rand_state(n), utils.rand_state(n)
        for rev in (False, True):
            prog = parse_program(h.debug_program(0, tb, low, rev, dense=True, fused=fused), dense=True)
            check_structure(prog, n)
            dense_check_tables(prog)
            units = sum(len(st[2]) for ps in prog for st in ps["stages"])
            assert units == n + circ.num_blocks + getattr(circ, "half_layer_num_blocks", 0)
            ref = O.apply_v(circ, th, y, dagger=rev)
            (v,), _ = replay(prog, circ.entangler, th, [y], dagger=rev, grad=False)
            assert _rel(v, ref) < TOL, (name, rev)
            (v,), _ = dense_emulate(prog, circ.entangler, th, [y], dagger=rev, grad=False)
            assert _rel(v, ref) < TOL, (name, rev)
        z0 = O.apply_v(circ, th, y, dagger=True)
        prog = parse_program(h.debug_program(0, tb, low, False, dense=True, fused=fused), dense=True)
        gref = O.grad_sweep(circ, th, x, z0)
        (w, z), g = replay(prog, circ.entangler, th, [x, z0], dagger=False, grad=True)
        assert _rel(g, gref) < TOL, name
        (w, z), g = dense_emulate(prog, circ.entangler, th, [x, z0], dagger=False, grad=True)
        assert _rel(g, gref) < TOL, name
        assert _rel(w, O.apply_v(circ, th, x)) < TOL and _rel(z, y) < 1e-10


def test_dense_matrix_layout():
    np.random.seed(4100)
    n, k = 4, 3
    m = 1 << k
    for ent in ("cx", "cp"):
        circ = ParametricCircuit(n, ent, utils.rand_circuit(n, 8))
        h = CircuitHandle(circ)
        th = utils.rand_thetas(circ.num_thetas)
        X = np.random.rand(2**n, m) + 1j * np.random.rand(2**n, m)
        Y = np.random.rand(2**n, m) + 1j * np.random.rand(2**n, m)
        z0 = O.apply_v(circ, th, Y.ravel(), dagger=True, ncols=m)
        prog = parse_program(h.debug_program(k, 6, 3, False, dense=True), dense=True)
        check_structure(prog, n + k)
        dense_check_tables(prog)
        _, g = dense_emulate(prog, ent, th, [X.ravel(), z0], dagger=False, grad=True)
        assert _rel(g, O.grad_sweep(circ, th, X.ravel(), z0, ncols=m)) < TOL


def test_dense_front_merge_and_bank_conflicts():
    """Benchmark-size Trotter program: no stage of its own for the front layer, no bank conflicts."""
    n, layers = 28, 4
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, layers), True)
    h = CircuitHandle(circ)
    for rev in (False, True):
        prog = parse_program(h.debug_program(0, 11, 4, rev, dense=True), dense=True)
        check_structure(prog, n)
        dense_check_tables(prog)
        pair_runs = (n - 1) * layers + n // 2
        nstages = sum(len(ps["stages"]) for ps in prog)
        assert nstages <= pair_runs + n // 4, (nstages, pair_runs)
        assert dense_bank_conflicts(prog) == (1, 1)


def test_fused_steps_at_benchmark_size():
    """Most stages of the n = 28 gradient program pair up into fused steps."""
    from program_sim import dense_steps

    n = 28
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 4), True)
    prog = parse_program(CircuitHandle(circ).debug_program(0, 11, 4, False, dense=True, fused=True), dense=True)
    dense_check_tables(prog)
    stages = sum(len(ps["stages"]) for ps in prog)
    steps = sum(len(dense_steps(ps["stages"])) for ps in prog)
    assert steps < 0.7 * stages, (steps, stages)


@pytest.mark.parametrize("n,layers,tb,low", [(12, 2, 6, 2), (13, 3, 7, 2), (14, 2, 8, 3)])
def test_tile_planner_needs_fewer_passes_and_replays_exactly(n, layers, tb, low, monkeypatch):
    """
    The planned tile sets (csrc/aqc_program.h plan_tiles) replace the greedy schedule only when they need fewer
    passes; the planned program is replayed gate by gate against the oracle, forward, reversed and with the
    gradient, like the greedy one above.  (n = 20 / 22 / 24 / 28 at the production tile sizes: 4 / 4 / 8 / 9
    instead of 5 / 6 / 13 / 20 gradient passes -- checked below without the replay.)
    """
    np.random.seed(6000 + n)
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, layers), True)
    h = CircuitHandle(circ)
    th = utils.rand_thetas(circ.num_thetas)
    x, y = utils.rand_state(n), utils.rand_state(n)
    for rev in (False, True):
        monkeypatch.setenv("AQC_TILE_PLAN", "0")
        greedy = parse_program(h.debug_program(0, tb, low, rev, dense=True, fused=True), dense=True)
        monkeypatch.setenv("AQC_TILE_PLAN", "1")
        planned = parse_program(h.debug_program(0, tb, low, rev, dense=True, fused=True), dense=True)
        assert len(planned) <= len(greedy)
        if not rev:
            assert len(planned) < len(greedy), "the planner is expected to win on a brick-wall circuit"
        check_structure(planned, n)
        dense_check_tables(planned)
        units = sum(len(st[2]) for ps in planned for st in ps["stages"])
        assert units == n + circ.num_blocks + circ.half_layer_num_blocks
        ref = O.apply_v(circ, th, y, dagger=rev)
        (v,), _ = dense_emulate(planned, circ.entangler, th, [y], dagger=rev, grad=False)
        assert _rel(v, ref) < TOL
    z0 = O.apply_v(circ, th, y, dagger=True)
    planned = parse_program(h.debug_program(0, tb, low, False, dense=True, fused=True), dense=True)
    (w, z), g = dense_emulate(planned, circ.entangler, th, [x, z0], dagger=False, grad=True)
    assert _rel(g, O.grad_sweep(circ, th, x, z0)) < TOL
    assert _rel(w, O.apply_v(circ, th, x)) < TOL and _rel(z, y) < 1e-10


def test_tile_planner_pass_counts_at_benchmark_sizes(monkeypatch):
    """Pass counts of the BASELINE configurations (host-side scheduling only, no state is touched)."""
    want = {(20, 2, 10, 2): (5, 4), (22, 2, 10, 2): (6, 4), (24, 4, 11, 3): (None, 8), (28, 4, 11, 3): (None, 9)}
    for (n, layers, tb, low), (greedy_want, planned_want) in want.items():
        circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, layers), True)
        h = CircuitHandle(circ)
        monkeypatch.setenv("AQC_TILE_PLAN", "1")
        planned = parse_program(h.debug_program(0, tb, low, False, dense=True, fused=True), dense=True)
        assert len(planned) <= planned_want, (n, len(planned))
        if greedy_want is not None:
            monkeypatch.setenv("AQC_TILE_PLAN", "0")
            greedy = parse_program(h.debug_program(0, tb, low, False, dense=True, fused=True), dense=True)
            assert len(greedy) == greedy_want


def test_tile_planner_on_generic_layouts(monkeypatch):
    """Random / spin layouts (units on arbitrary qubit pairs): planned programs stay exact, never need more passes."""
    n, tb, low = 10, 6, 2
    np.random.seed(6100)
    wins = 0
    for name, circ in _circuits(n):
        h = CircuitHandle(circ)
        th = utils.rand_thetas(circ.num_thetas)
        x, y = utils.rand_state(n), utils.rand_state(n)
        z0 = O.apply_v(circ, th, y, dagger=True)
        gref = O.grad_sweep(circ, th, x, z0)
        monkeypatch.setenv("AQC_TILE_PLAN", "0")
        greedy = parse_program(h.debug_program(0, tb, low, False, dense=True, fused=True), dense=True)
        monkeypatch.setenv("AQC_TILE_PLAN", "1")
        planned = parse_program(h.debug_program(0, tb, low, False, dense=True, fused=True), dense=True)
        assert len(planned) <= len(greedy), name
        wins += len(planned) < len(greedy)
        check_structure(planned, n)
        dense_check_tables(planned)
        (w, z), g = dense_emulate(planned, circ.entangler, th, [x, z0], dagger=False, grad=True)
        assert _rel(g, gref) < TOL, name
        assert _rel(w, O.apply_v(circ, th, x)) < TOL and _rel(z, y) < 1e-10, name
    assert wins >= 1
