"""
Host logic of SpSurrogateObjectiveMax on CPU: hysteresis of ``max_no``, weight smoothing, the ONE
fused sweep for the two weighted gradient terms, the early start of the gradient and its fallbacks.

The GPU workspace is replaced by a stand-in that answers the same calls (objective, set_sparse,
grad_begin / grad_end, grad, upload) with the NumPy oracle, so what is checked here is the Python
layer above the C-ABI -- against the golden call sequences recorded from the unmodified reference
(tests/golden/objective_sequences.npz; objective_lhs_sur_max.py:82-191).  The kernels themselves are
checked by the ``-m gpu`` tests.
"""

import numpy as np
import pytest

from golden_util import load, rel
from aqc_research_b200.engine import CircuitHandle
from aqc_research_b200.model_sp_lhs import objective_base
from aqc_research_b200.model_sp_lhs.objective_lhs_sur_max import SpSurrogateObjectiveMax
from aqc_research_b200.parametric_circuit import TrotterAnsatz
from oracle import sv_oracle as O

TOL = 1e-10


class OracleWorkspace:
    """Answers the SvWorkspace calls the objective makes, with oracle/sv_oracle.py underneath."""

    calls = []

    def __init__(self, circ, num_slots, device=0):
        self.circ = circ
        self.circuit = CircuitHandle(circ)  # host-only part of the C-ABI (no GPU needed)
        self.size = 2**circ.num_qubits
        self.slots = [np.zeros(self.size, dtype=np.complex128) for _ in range(num_slots)]
        self.pending = None

    def upload(self, slot, data, batch_index=-1):
        self.slots[slot] = np.array(data, dtype=np.complex128).ravel()

    def objective(self, thetas, target, z0, indices):
        self.pending = None  # a new call drops an uncollected sweep
        self.slots[z0] = O.apply_v(self.circ, np.asarray(thetas), self.slots[target], dagger=True)
        OracleWorkspace.calls.append("objective")
        return self.slots[z0][np.asarray(indices)][None, :]

    def set_sparse(self, slot, indices, amplitudes):
        assert 1 <= len(indices) <= 8
        v = np.zeros(self.size, dtype=np.complex128)
        for i, a in zip(indices, amplitudes):
            v[int(i)] = a
        self.slots[slot] = v
        OracleWorkspace.calls.append("set_sparse")

    def grad_begin(self, thetas, *, z0, w, z, x_slot=-1, x_basis=0):
        if x_slot >= 0:
            x = self.slots[x_slot].copy()
        else:
            x = np.zeros(self.size, dtype=np.complex128)
            x[int(x_basis)] = 1.0
        self.pending = O.grad_sweep(self.circ, np.asarray(thetas), x, self.slots[z0])
        OracleWorkspace.calls.append("sweep")

    def eval_begin(self, thetas, target, z0, indices, *, x_basis, w, z):
        hs = self.objective(thetas, target, z0, indices)
        self.grad_begin(thetas, z0=z0, w=w, z=z, x_basis=x_basis)
        OracleWorkspace.calls.append("fused")
        return hs

    def grad_end(self):
        assert self.pending is not None, "no gradient sweep in flight"
        out, self.pending = self.pending, None
        return out[None, :]

    def grad(self, thetas, **kw):
        self.grad_begin(thetas, **kw)
        return self.grad_end()

    def close(self):
        pass


@pytest.fixture()
def fake_gpu(monkeypatch):
    monkeypatch.setattr(objective_base, "SvWorkspace", OracleWorkspace)
    OracleWorkspace.calls = []
    yield OracleWorkspace


def _params(n, **kw):
    p = dict(num_qubits=n, max_flips=1, maxiter=10, verbose=0, enable_optim_stats=False,
             num_simulations=1, trunc_thr=1e-6, state_prep_func=None)
    p.update(kw)
    return p


def test_golden_sequences_with_one_sweep_per_evaluation(fake_gpu):
    g = load("objective_sequences.npz")
    seen_two_term = False
    for c in range(int(g["num_sp"])):
        p = f"sp{c}_"
        n, _, steps = [int(v) for v in g[p + "meta"]]
        if n > 8:
            continue
        circ = TrotterAnsatz(n, g[p + "blocks"], True)
        objv = SpSurrogateObjectiveMax(user_parameters=_params(n), circ=circ, front_layer=True)
        objv.set_target(g[p + "target"])
        for s in range(steps):
            th = g[p + "thetas"][s]
            fake_gpu.calls.clear()
            assert abs(objv.objective(th) - g[p + "f"][s]) < TOL
            assert objv.max_no == int(g[p + "max_no"][s])
            grad = objv.gradient(th)
            assert rel(grad, g[p + "grad"][s]) < TOL
            assert abs(objv.weight - g[p + "weight"][s]) < TOL
            # one V^H sweep and ONE gradient sweep, also when two weighted terms are needed
            assert fake_gpu.calls.count("objective") == 1 and fake_gpu.calls.count("sweep") == 1
            if objv.max_no != 0:
                seen_two_term = True
                assert "set_sparse" in fake_gpu.calls
    assert seen_two_term


def test_dense_state_handler_uses_one_sweep_too(fake_gpu):
    g = load("objective_sequences.npz")
    p = "sp3_"  # n = 7, max_no = 7 throughout
    n, _, steps = [int(v) for v in g[p + "meta"]]
    circ = TrotterAnsatz(n, g[p + "blocks"], True)
    dense = np.zeros((n + 1, 2**n), dtype=np.complex128)
    for i, k in enumerate(O.basis_state_indices(n, 0)):
        dense[i, k] = 1
    objv = SpSurrogateObjectiveMax(user_parameters=_params(n, state_prep_func=lambda nq: dense), circ=circ,
                                   front_layer=True)
    objv.set_target(g[p + "target"])
    # the dense handler takes its overlaps through vdot
    OracleWorkspace.vdot = lambda self, a, b: np.array([np.vdot(self.slots[a], self.slots[b])])
    OracleWorkspace.apply = lambda self, th, src, dst, dagger=False: self.slots.__setitem__(
        dst, O.apply_v(self.circ, np.asarray(th), self.slots[src], dagger=dagger))
    for s in range(steps):
        th = g[p + "thetas"][s]
        fake_gpu.calls.clear()
        assert abs(objv.objective(th) - g[p + "f"][s]) < TOL
        assert rel(objv.gradient(th), g[p + "grad"][s]) < TOL
        assert fake_gpu.calls.count("sweep") == 1
        assert abs(objv.weight - g[p + "weight"][s]) < TOL


def test_early_start_follows_the_fun_jac_pattern(fake_gpu):
    g = load("objective_sequences.npz")
    p = "sp2_"  # n = 4, five steps, max_no != 0
    n, _, steps = [int(v) for v in g[p + "meta"]]
    circ = TrotterAnsatz(n, g[p + "blocks"], True)
    objv = SpSurrogateObjectiveMax(user_parameters=_params(n), circ=circ, front_layer=True)
    objv.set_target(g[p + "target"])
    order = []
    for s in range(steps):
        th = g[p + "thetas"][s]
        fake_gpu.calls.clear()
        objv.objective(th)
        order.append("sweep" in fake_gpu.calls)  # was the sweep enqueued by objective() already?
        assert rel(objv.gradient(th), g[p + "grad"][s]) < TOL
        assert fake_gpu.calls.count("sweep") == 1
    assert order == [False, False] + [True] * (steps - 2)  # learnt after two fun/jac pairs
    # a gradient asked at OTHER angles while an early sweep is in flight: recomputed, early start off
    th_a, th_b = g[p + "thetas"][0], g[p + "thetas"][1]
    objv.objective(th_a)
    w_before = objv.weight
    grad_b = objv.gradient(th_b)
    _, _, ref_b, _ = O.sur_max_value_and_grad(circ, th_b, g[p + "target"], w_before, objv.max_no)
    assert rel(grad_b, ref_b) < TOL
    assert not objv._early_on


def test_fused_evaluation_when_the_leader_is_state_zero(fake_gpu):
    """
    Leading state |s_0> and the fun / jac pattern learnt: objective() enqueues the WHOLE evaluation
    (aqc_sv_eval_begin: V^H sweep, gather, gradient sweep) and gradient() only collects it -- one sweep per
    evaluation, results equal the oracle's; a leader change drops the speculative sweep and takes the
    two-term sweep instead.
    """
    from aqc_research_b200 import circuit_structures as cs

    rng = np.random.RandomState(12)
    n = 5
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    th_star = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    e0 = np.zeros(2**n, dtype=np.complex128)
    e0[0] = 1
    target = O.apply_v(circ, th_star, e0)  # near target: the leader is |0>
    objv = SpSurrogateObjectiveMax(user_parameters=_params(n), circ=circ, front_layer=True)
    objv.set_target(target)
    fused_seen = 0
    for step in range(6):
        th = th_star + 0.02 * (2 * rng.rand(circ.num_thetas) - 1)
        w_before = objv.weight
        fake_gpu.calls.clear()
        f = objv.objective(th)
        fused = "fused" in fake_gpu.calls
        g = objv.gradient(th)
        f_ref, _, g_ref, _ = O.sur_max_value_and_grad(circ, th, target, w_before, 0)
        assert objv.max_no == 0 and abs(f - f_ref) < TOL and rel(g, g_ref) < TOL
        assert fake_gpu.calls.count("sweep") == 1 and fused == (step >= 2)
        fused_seen += fused
    assert fused_seen == 4
    # another target: the leader changes inside a fused objective() -> the speculative sweep is not used
    y = rng.rand(2**n) + 1j * rng.rand(2**n)
    y /= np.linalg.norm(y)
    objv.set_target(y)
    objv._early_on, objv._early_hits = True, 2
    th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    w_before = objv.weight
    fake_gpu.calls.clear()
    f = objv.objective(th)
    g = objv.gradient(th)
    assert objv.max_no != 0 and "fused" in fake_gpu.calls and "set_sparse" in fake_gpu.calls
    f_ref, _, g_ref, _ = O.sur_max_value_and_grad(circ, th, y, w_before, objv.max_no)
    assert abs(f - f_ref) < TOL and rel(g, g_ref) < TOL


class OracleMpsWorkspace:
    """Stand-in for MpsWorkspace: states kept as dense vectors (no truncation), oracle underneath."""

    calls = []

    def __init__(self, circ, num_slots, chi_max=64, trunc_thr=1e-16, device=0):
        self.circ = circ
        self.size = 2**circ.num_qubits
        self.slots = [np.zeros(self.size, dtype=np.complex128) for _ in range(num_slots)]

    def upload(self, slot, mps):
        from oracle import mps_oracle as M

        self.slots[slot] = M.mps_to_vector(mps)

    def objective(self, thetas, target, z0, indices):
        self.slots[z0] = O.apply_v(self.circ, np.asarray(thetas), self.slots[target], dagger=True)
        return self.slots[z0][np.asarray(indices)]

    def set_product_site(self, slot, index, site, amp0, amp1):
        v = np.zeros(self.size, dtype=np.complex128)
        v[index & ~(1 << site)] = amp0  # qubit `site` in |0>
        v[index | (1 << site)] = amp1   # ... in |1>
        assert abs(abs(amp0) ** 2 + abs(amp1) ** 2 - 1.0) < 1e-14  # the caller keeps the pair normalised
        self.slots[slot] = v
        OracleMpsWorkspace.calls.append("set_product_site")

    def grad(self, thetas, *, z0, w, z, x_slot=-1, x_basis=0):
        if x_slot >= 0:
            x = self.slots[x_slot].copy()
        else:
            x = np.zeros(self.size, dtype=np.complex128)
            x[int(x_basis)] = 1.0
        OracleMpsWorkspace.calls.append("sweep")
        return O.grad_sweep(self.circ, np.asarray(thetas), x, self.slots[z0])

    def close(self):
        pass


def test_mps_objective_two_terms_from_one_product_state(monkeypatch):
    """
    SpSurrogateObjectiveFastMpsTrotter: the weighted pair (s_0, X_i s_0) is one product state with
    qubit i in superposition; which amplitude sits on |0> depends on the bit of s_0 (Neel state:
    both cases occur).  Golden values: the reference's state-vector objective on the same states.
    """
    from aqc_research_b200.model_sp_lhs import objective_lhs_sur_fast_mps_trotter as mod
    from oracle import mps_oracle as M

    monkeypatch.setattr(mod, "MpsWorkspace", OracleMpsWorkspace)
    OracleMpsWorkspace.calls = []
    g = load("objective_sequences.npz")
    for p in ("sp2_", "sp3_"):
        n, _, steps = [int(v) for v in g[p + "meta"]]
        circ = TrotterAnsatz(n, g[p + "blocks"], True)
        objv = mod.SpSurrogateObjectiveFastMpsTrotter(
            user_parameters=_params(n, trunc_thr=1e-16), circ=circ)
        objv.set_target(M.vector_to_mps(g[p + "target"]))
        for s in range(steps):
            th = g[p + "thetas"][s]
            OracleMpsWorkspace.calls.clear()
            assert abs(objv.objective(th) - g[p + "f"][s]) < TOL
            assert objv.max_no == int(g[p + "max_no"][s]) and objv.max_no != 0
            assert rel(objv.gradient(th), g[p + "grad"][s]) < TOL
            assert OracleMpsWorkspace.calls == ["set_product_site", "sweep"]
    # Neel preparation: flipped qubits with bit 1 (even sites) and bit 0 (odd sites) in s_0
    n = 4
    np.random.seed(5)
    from aqc_research_b200 import circuit_structures as cs, utils

    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    target, neel = utils.rand_state(n), 0b0101
    objv = mod.SpSurrogateObjectiveFastMpsTrotter(
        user_parameters=_params(n, trunc_thr=1e-16, state_prep_func=lambda nq: neel), circ=circ)
    objv.set_target(M.vector_to_mps(target))
    seen = set()
    for _ in range(12):
        th = utils.rand_thetas(circ.num_thetas)
        w_before = objv.weight
        objv.objective(th)
        grad = objv.gradient(th)
        _, _, ref, _ = O.sur_max_value_and_grad(circ, th, target, w_before, objv.max_no, init_index=neel)
        assert rel(grad, ref) < TOL
        if objv.max_no:
            seen.add((neel >> (objv.max_no - 1)) & 1)
    assert seen == {0, 1}


def test_mps_gate_helpers_host_logic(monkeypatch):
    """
    The gate-by-gate MPS helpers (mps_dot_objective.py:245-516 of the reference) are tiny circuits
    pushed through the engine: angles, cancelling blocks, the factor i of the Pauli gates.  Checked
    here with the dense stand-in; tests/test_mps_gpu.py repeats it on the GPU engine.
    """
    from aqc_research_b200 import mps_dot_objective as mdo
    from aqc_research_b200 import mps_operations as mpsop
    from oracle import mps_oracle as M

    class Dense(OracleMpsWorkspace):
        def apply(self, thetas, src, dst, dagger=False):
            self.slots[dst] = O.apply_v(self.circ, np.asarray(thetas), self.slots[src], dagger=dagger)

        def download(self, slot):
            return M.vector_to_mps(self.slots[slot])

        def dot(self, a, b):
            return complex(np.vdot(self.slots[a], self.slots[b]))

        def check_cap(self, what):  # the dense stand-in never truncates
            pass

    monkeypatch.setattr(mpsop, "MpsWorkspace", Dense)
    rng = np.random.RandomState(21)
    n, ang = 4, 0.9173
    v = rng.randn(2**n) + 1j * rng.randn(2**n)
    v /= np.linalg.norm(v)
    mv = M.vector_to_mps(v)
    for q in range(n):
        for fn, g in ((mdo.x_mul_mps, O.PAULI_X), (mdo.y_mul_mps, O.PAULI_Y), (mdo.z_mul_mps, O.PAULI_Z)):
            assert rel(M.mps_to_vector(fn(q, mv)), O.op1(v.copy(), q, g)) < TOL, (fn.__name__, q)
        for fn, mk in ((mdo.rx_mul_mps, O.rx), (mdo.ry_mul_mps, O.ry), (mdo.rz_mul_mps, O.rz)):
            assert rel(M.mps_to_vector(fn(ang, q, mv)), O.op1(v.copy(), q, mk(ang))) < TOL, (fn.__name__, q)
    for c, t in ((0, 1), (1, 0), (3, 2)):
        assert rel(M.mps_to_vector(mdo.cx_mul_mps(0.0, c, t, mv)), O.ctrl_op(v.copy(), c, t, O.PAULI_X)) < TOL
        assert rel(M.mps_to_vector(mdo.cz_mul_mps(0.0, c, t, mv)), O.ctrl_op(v.copy(), c, t, O.PAULI_Z)) < TOL
        assert rel(M.mps_to_vector(mdo.cp_mul_mps(ang, c, t, mv)), O.ctrl_op(v.copy(), c, t, O.phase(ang))) < TOL


class OracleMatWorkspace:
    """Stand-in for the matrix-path SvWorkspace used by SketchingObjectiveEx (dense oracle underneath)."""

    def __init__(self, circ, num_slots, device=0, log2_cols=0, batch=1, as_generic=False):
        self.circ, self.ncols = circ, 1 << log2_cols
        self.size = (2**circ.num_qubits) * self.ncols
        self.slots = [np.zeros(self.size, dtype=np.complex128) for _ in range(num_slots)]
        self.sweeps = 0

    def upload(self, slot, data, batch_index=-1):
        self.slots[slot] = np.array(data, dtype=np.complex128).ravel()

    def set_identity(self, slot):
        self.slots[slot] = np.eye(2**self.circ.num_qubits, dtype=np.complex128).ravel()

    def apply(self, thetas, src, dst, dagger=False):
        self.slots[dst] = O.apply_v(self.circ, np.asarray(thetas), self.slots[src], dagger=dagger, ncols=self.ncols)

    def vdot(self, a, b):
        return np.array([np.vdot(self.slots[a], self.slots[b])])

    def grad(self, thetas, *, z0, w, z, x_slot=-1, x_basis=0):
        self.sweeps += 1
        return O.grad_sweep(self.circ, np.asarray(thetas), self.slots[x_slot], self.slots[z0], ncols=self.ncols)[None, :]

    def close(self):
        pass


def test_sketching_objective_host_logic(monkeypatch):
    """
    SketchingObjectiveEx + FullRangeSketchingVectors (sk_core.py:94-326): f = 1 - Re Tr(V^H U) / d,
    g = -Re(grad) / d, the (f, g) cache keyed on theta, best-point bookkeeping -- against the golden
    sequences of the reference, with the oracle standing in for the GPU workspace.
    """
    from golden_util import KINDS
    from aqc_research_b200.model_sketching import sk_core
    from aqc_research_b200.parametric_circuit import ParametricCircuit

    monkeypatch.setattr(sk_core, "SvWorkspace", OracleMatWorkspace)
    g = load("objective_sequences.npz")
    for c in range(int(g["num_sk"])):
        p = f"sk{c}_"
        n, kind = [int(v) for v in g[p + "meta"]]
        circ = ParametricCircuit(n, KINDS[kind], g[p + "blocks"])
        objv = sk_core.SketchingObjectiveEx(circ, sk_core.FullRangeSketchingVectors(g[p + "target"]), enable_stats=True)
        ths = g[p + "thetas"]
        for s in range(ths.shape[0]):
            assert abs(objv.objective(ths[s]) - g[p + "f"][s]) < TOL
            before = objv.workspace.sweeps
            assert rel(objv.gradient(ths[s]), g[p + "grad"][s]) < TOL
            assert objv.workspace.sweeps == before  # same theta: the cached gradient is returned
        assert objv.num_iterations == ths.shape[0]
        grad = objv.gradient(ths[0])  # gradient first at other angles: objective is recomputed
        assert rel(grad, g[p + "grad"][0]) < TOL and objv.num_iterations == ths.shape[0] + 1
        res = objv.optim_results
        best = int(np.argmin(g[p + "f"]))
        assert abs(res["cost"] - g[p + "f"][best]) < TOL and np.array_equal(res["thetas"], ths[best])
        assert objv.statistics["nit"] == ths.shape[0] + 1


def test_structure_change_rebuilds_the_workspace_and_drops_cached_results(fake_gpu):
    """
    insert_unit_blocks / update_structure between calls: the next objective OR gradient
    call must run on the new layout -- also a gradient() on the angles of the last objective() -- and
    a device-generated random target must survive the rebuild.
    """
    from aqc_research_b200 import circuit_structures as cs
    from aqc_research_b200.parametric_circuit import ParametricCircuit

    rng = np.random.RandomState(3)
    n = 5
    circ = ParametricCircuit(n, "cx", cs.create_ansatz_structure(n, "spin", "full", 8))
    target = rng.rand(2**n) + 1j * rng.rand(2**n)
    target /= np.linalg.norm(target)
    objv = SpSurrogateObjectiveMax(user_parameters=_params(n), circ=circ, front_layer=True)
    objv.set_target(target)
    th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    objv.objective(th)
    objv.gradient(th)
    first_ws = objv.workspace
    th2, new_idx = circ.insert_unit_blocks(4, np.array([[0, 3], [1, 2]]), th)
    th2[new_idx] = 0.3
    # gradient straight away on a vector whose size changed: the objective is recomputed on the new layout
    g = objv.gradient(th2)
    assert objv.workspace is not first_ws and objv.workspace.circ.num_blocks == 10
    fresh = SpSurrogateObjectiveMax(user_parameters=_params(n), circ=circ, front_layer=True)
    fresh.set_target(target)
    fresh.objective(th2)
    assert rel(objv._hs2, fresh._hs2) < TOL  # |<s_i|V^H t>|^2 of the NEW circuit
    # same block count, different layout, same angles: the cached objective must not be reused
    blocks = circ.blocks.copy()
    blocks[:, 0] = blocks[::-1, 0]
    circ.update_structure(blocks)
    fake_gpu.calls.clear()
    objv.gradient(th2)
    assert "objective" in fake_gpu.calls and objv.workspace.circ is circ
    assert np.all(np.isfinite(g))
