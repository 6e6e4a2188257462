"""
GPU parity tests (state-vector and matrix paths): CUDA engine through the C-ABI vs the CPU
oracle on identical seeded inputs.  Tolerance: 1e-10 relative, norm-wise (north star).
"""

import os

import numpy as np
import pytest

from aqc_research_b200 import circuit_structures as cs
from aqc_research_b200 import core_op_matrix as cpm
from aqc_research_b200 import core_operations as cop
from aqc_research_b200 import utils
from aqc_research_b200.engine import SvWorkspace
from aqc_research_b200.parametric_circuit import ParametricCircuit, TrotterAnsatz
from oracle import sv_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-10  # relative, norm-wise


def _rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300)


def _circuits(n, layers=2, depth=9):
    yield "t1", TrotterAnsatz(n, cs.make_trotter_like_circuit(n, layers), False)
    yield "t2", TrotterAnsatz(n, cs.make_trotter_like_circuit(n, layers), True)
    for ent in ("cx", "cz", "cp"):
        yield ent, ParametricCircuit(n, ent, utils.rand_circuit(n, depth))


@pytest.mark.parametrize("n", [2, 3, 4, 5, 7, 10, 12, 13])
def test_apply_and_grad_vs_oracle(n):
    np.random.seed(4242 + n)
    for name, circ in _circuits(n):
        th = utils.rand_thetas(circ.num_thetas)
        x, y = utils.rand_state(n), utils.rand_state(n)
        out = np.zeros_like(y)
        ws_unused = np.zeros((3, y.size), dtype=np.complex128)
        v = cop.v_mul_vec(circ, th, y, out, ws_unused).copy()
        assert _rel(v, O.apply_v(circ, th, y)) < TOL, (name, "V")
        z0 = cop.v_dagger_mul_vec(circ, th, y, out, ws_unused).copy()
        assert _rel(z0, O.apply_v(circ, th, y, dagger=True)) < TOL, (name, "VH")
        g = cop.grad_of_dot_product(circ, th, x, z0, ws_unused)
        assert _rel(g, O.grad_sweep(circ, th, x, z0)) < TOL, (name, "grad")
        if circ.num_blocks > 4:
            br = (2, circ.num_blocks - 1)
            g = cop.grad_of_dot_product(circ, th, x, z0, ws_unused, block_range=br, front_layer=False)
            assert _rel(g, O.grad_sweep(circ, th, x, z0, br, False)) < TOL, (name, "partial grad")
    cop.clear_workspace_cache()


def test_round_trip_and_aliasing():
    """V V^H v = v (test_core_operations.py:252-281), with out aliasing vec."""
    np.random.seed(5)
    n = 9
    for name, circ in _circuits(n, layers=3):
        th = utils.rand_thetas(circ.num_thetas)
        v = utils.rand_state(n)
        buf = v.copy()
        cop.v_dagger_mul_vec(circ, th, buf, buf)
        cop.v_mul_vec(circ, th, buf, buf)
        assert _rel(buf, v) < TOL, name
    cop.clear_workspace_cache()


@pytest.mark.parametrize("n,m", [(2, 2), (3, 5), (4, 16), (5, 7), (7, 128)])
def test_matrix_path_vs_oracle(n, m):
    np.random.seed(99 + n + m)
    for ent in ("cx", "cz", "cp"):
        circ = ParametricCircuit(n, ent, utils.rand_circuit(n, 12))
        th = utils.rand_thetas(circ.num_thetas)
        X = np.random.rand(2**n, m) + 1j * np.random.rand(2**n, m)
        Y = np.random.rand(2**n, m) + 1j * np.random.rand(2**n, m)
        a = cpm.v_mul_mat(circ, th, Y.copy())
        assert _rel(a, O.apply_v(circ, th, Y.ravel(), False, m).reshape(Y.shape)) < TOL
        z0 = cpm.v_dagger_mul_mat(circ, th, Y.copy())
        z0_ref = O.apply_v(circ, th, Y.ravel(), True, m)
        assert _rel(z0, z0_ref.reshape(Y.shape)) < TOL
        xw, zw = X.copy(), z0.copy()
        g = cpm.grad_of_matrix_dot_product(circ, th, xw, zw)
        assert _rel(g, O.grad_sweep(circ, th, X.ravel(), z0.ravel(), ncols=m)) < TOL
    cop.clear_workspace_cache()


def test_batched_thetas():
    """batch independent angle sets in one launch sequence == one at a time."""
    np.random.seed(31)
    n, batch = 6, 5
    circ = ParametricCircuit(n, "cx", cs.create_ansatz_structure(n, "cyclic_spin", "full", 24))
    ws = SvWorkspace(circ, num_slots=4, batch=batch)
    ths = np.stack([utils.rand_thetas(circ.num_thetas) for _ in range(batch)])
    y, x = utils.rand_state(n), utils.rand_state(n)
    ws.upload(0, y)
    ws.upload(3, x)
    ws.apply(ths, 0, 1, dagger=True)
    g = ws.grad(ths, x_slot=3, z0=1, w=2, z=0)
    for b in range(batch):
        z0 = O.apply_v(circ, ths[b], y, dagger=True)
        assert _rel(g[b], O.grad_sweep(circ, ths[b], x, z0)) < TOL
    ws.close()


def test_objective_gather_basis_and_vdot():
    np.random.seed(8)
    n = 11
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    ws = SvWorkspace(circ, num_slots=4)
    th = utils.rand_thetas(circ.num_thetas)
    y = utils.rand_state(n)
    ws.upload(0, y)
    idx = O.basis_state_indices(n, init_index=0b10101010101 & (2**n - 1))
    hs = ws.objective(th, 0, 1, idx)[0]
    z0 = O.apply_v(circ, th, y, dagger=True)
    assert _rel(hs, z0[idx]) < TOL
    g = ws.grad(th, x_basis=int(idx[3]), z0=1, w=2, z=3)[0]
    e = np.zeros(2**n, dtype=np.complex128)
    e[idx[3]] = 1
    assert _rel(g, O.grad_sweep(circ, th, e, z0)) < TOL
    # slot 1 (z0) must be intact after the sweep that wrote into slots 2, 3
    assert _rel(ws.download(1), z0) < TOL
    ws.set_basis(2, 5)
    assert ws.gather(2, [4, 5, 6])[0].tolist() == [0, 1, 0]
    ws.fill_random(2, 123)
    r = ws.download(2)
    assert abs(np.linalg.norm(r) - 1) < 1e-12 and r.real.min() >= 0 and r.imag.min() >= 0
    ws.close()


@pytest.mark.parametrize("n", [6, 12, 15])
def test_fused_evaluation_vs_oracle(n):
    """
    aqc_sv_eval_begin / aqc_sv_eval_hs / aqc_sv_grad_end: the whole evaluation enqueued at once equals the
    two separate calls and the oracle; an evaluation whose gradient is never collected is dropped cleanly.
    """
    np.random.seed(600 + n)
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    ws = SvWorkspace(circ, num_slots=4)
    y = utils.rand_state(n)
    ws.upload(0, y)
    idx = O.basis_state_indices(n, init_index=5 % 2**n)
    for rep in range(3):
        th = utils.rand_thetas(circ.num_thetas)
        hs = ws.eval_begin(th, 0, 1, idx, x_basis=int(idx[0]), w=2, z=3)[0]
        if rep == 1:
            continue  # not collected: the next call drops the sweep in flight
        g = ws.grad_end()[0]
        z0 = O.apply_v(circ, th, y, dagger=True)
        e = np.zeros(2**n, dtype=np.complex128)
        e[idx[0]] = 1
        assert _rel(hs, z0[idx]) < TOL and _rel(g, O.grad_sweep(circ, th, e, z0)) < TOL
        o_ms, g_ms = ws.eval_times()
        assert o_ms > 0 and g_ms > 0
    # and the plain calls still agree afterwards
    th = utils.rand_thetas(circ.num_thetas)
    hs = ws.objective(th, 0, 1, idx)[0]
    g = ws.grad(th, x_basis=int(idx[0]), z0=1, w=2, z=3)[0]
    z0 = O.apply_v(circ, th, y, dagger=True)
    e = np.zeros(2**n, dtype=np.complex128)
    e[idx[0]] = 1
    assert _rel(hs, z0[idx]) < TOL and _rel(g, O.grad_sweep(circ, th, e, z0)) < TOL
    ws.close()


@pytest.mark.parametrize("n", [16, 20])
def test_mid_size_vs_oracle(n):
    """Multi-pass programs with many tiles (sizes the NumPy oracle finishes in seconds)."""
    np.random.seed(n)
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 1), True)
    th = utils.rand_thetas(circ.num_thetas)
    y = utils.rand_state(n)
    ws = SvWorkspace(circ, num_slots=4)
    ws.upload(0, y)
    idx = O.basis_state_indices(n)
    hs = ws.objective(th, 0, 1, idx)[0]
    z0 = O.apply_v(circ, th, y, dagger=True)
    assert _rel(hs, z0[idx]) < TOL
    assert _rel(ws.download(1), z0) < TOL
    g = ws.grad(th, x_basis=0, z0=1, w=2, z=3)[0]
    e = np.zeros(2**n, dtype=np.complex128)
    e[0] = 1
    assert _rel(g, O.grad_sweep(circ, th, e, z0)) < TOL
    ws.close()


def test_large_properties():
    """n = 26: size-independent properties (V V^H = I, <w|z> invariance)."""
    n = 26
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    np.random.seed(26)
    th = utils.rand_thetas(circ.num_thetas)
    ws = SvWorkspace(circ, num_slots=4)
    ws.fill_random(0, 7)
    idx = O.basis_state_indices(n)
    hs = ws.objective(th, 0, 1, idx)[0]
    nrm = ws.vdot(1, 1)[0]
    assert abs(nrm - 1) < 1e-10  # unitarity
    g = ws.grad(th, x_basis=0, z0=1, w=2, z=3)[0]
    # V V^H y = y and <V e0 | y> = <e0 | V^H y> = hs[0] through the apply path
    ws.apply(th, 1, 3, dagger=False)
    assert abs(ws.vdot(3, 0)[0] - 1) < 1e-10
    ws.set_basis(2, 0)
    ws.apply(th, 2, 2, dagger=False)
    assert abs(ws.vdot(2, 0)[0] - hs[0]) < 1e-10
    # gradient sum rule: the front-layer Rz derivatives of qubit q obey
    # d/dt0 + ... ; cheap global check: the gradient is finite and not identically zero
    assert np.all(np.isfinite(g)) and np.linalg.norm(g) > 0
    ws.close()


def test_full_size_properties_n28():
    """
    BASELINE.json configs[4] at full size (n = 28, 4 layers, 2nd-order Trotter) through
    size-independent properties: V^H is norm preserving, the gradient sweep ends with
    (w, z) = (V e0, V V^H y = y), <V e0|y> = hs[0], and the complex gradient of <V e0|y> agrees with
    central finite differences for angles of the front layer, a full layer and the re-used half layer.
    """
    n = 28
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 4), True)
    np.random.seed(28)
    th = utils.rand_thetas(circ.num_thetas)
    ws = SvWorkspace(circ, num_slots=5)
    ws.fill_random(0, 11)
    idx = O.basis_state_indices(n)
    hs = ws.objective(th, 0, 1, idx)[0]
    assert abs(ws.vdot(1, 1)[0] - 1) < 1e-10
    g = ws.grad(th, x_basis=0, z0=1, w=2, z=3)[0]
    if os.environ.get("AQC_ENGINE", "dense") != "scaled":  # the scale-free engine leaves rescaled work states
        assert abs(ws.vdot(3, 0)[0] - 1) < 1e-10  # z came back to the target
        assert abs(ws.vdot(2, 0)[0] - hs[0]) < 1e-10  # <V e0 | y> = <e0 | V^H y>
    for k in (2, 3 * n + 1, 3 * n + 4 * 100 + 2, circ.num_thetas - 1):
        vals = []
        for sgn in (+1, -1):
            t2 = th.copy()
            t2[k] += sgn * 1e-4
            ws.set_basis(4, 0)
            ws.apply(t2, 4, 4, dagger=False)
            vals.append(ws.vdot(4, 0)[0])
        fd = (vals[0] - vals[1]) / 2e-4
        assert abs(fd - g[k]) < 1e-7 * max(1.0, abs(g[k])) + 1e-9, (k, fd, g[k])
    ws.close()


def test_circuit_transform_and_trotter_class():
    """
    ansatz_to_numpy_fast / ansatz_to_numpy_trotter (circuit_transform.py:273-390) against the dense
    oracle; Trotter.as_vector / as_mps (trotter.py:97-163) against the oracle's Trotter circuit.
    """
    from aqc_research_b200 import circuit_structures as cs
    from aqc_research_b200 import circuit_transform as ctr
    from aqc_research_b200.model_sp_lhs.trotter import trotter as trot
    from aqc_research_b200.parametric_circuit import ParametricCircuit, TrotterAnsatz
    from oracle import mps_oracle as M
    from oracle import sv_oracle as O

    rng = np.random.RandomState(77)
    n = 4
    circ = ParametricCircuit(n, "cx", cs.create_ansatz_structure(n, "spin", "full", 7))
    th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    eye = np.eye(2**n, dtype=np.complex128)
    ref_mat = O.apply_v(circ, th, eye.copy().ravel(), ncols=2**n).reshape(2**n, 2**n)
    assert np.linalg.norm(ctr.ansatz_to_numpy_fast(circ, th) - ref_mat) < 1e-12
    tro = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 2), True)
    th = np.pi * (2 * rng.rand(tro.num_thetas) - 1)
    cols = np.stack([O.apply_v(tro, th, eye[:, k].copy()) for k in range(2**n)], axis=1)
    got = ctr.ansatz_to_numpy_trotter(tro, th)
    assert np.linalg.norm(got - cols) < 1e-12 and np.linalg.norm(got.conj().T @ got - eye) < 1e-12
    with pytest.raises(ValueError):
        ctr.ansatz_to_numpy_fast(tro, th)
    with pytest.raises(NotImplementedError):
        ctr.ansatz_to_qcircuit(circ, th)
    # Trotter class: the TrotterAnsatz at the init_ansatz_to_trotter angles
    tr = trot.Trotter(num_qubits=5, evol_time=0.9, num_steps=3, delta=1.1, second_order=True)
    neel = trot.neel_init_state(5)
    c5 = TrotterAnsatz(5, cs.make_trotter_like_circuit(5, 3), True)
    t5 = trot.init_ansatz_to_trotter(c5, np.zeros(c5.num_thetas), evol_time=0.9, delta=1.1)
    v0 = np.zeros(32, dtype=np.complex128)
    v0[trot.basis_index(neel)] = 1
    want = O.apply_v(c5, t5, v0)
    assert np.linalg.norm(tr.as_vector(neel) - want) < 1e-12
    assert np.linalg.norm(tr.as_vector(v0) - want) < 1e-12
    dense = np.zeros(32, dtype=np.complex128)
    mps = tr.as_mps(neel, out_state=dense)
    assert np.linalg.norm(dense - want) < 1e-10 and np.linalg.norm(M.mps_to_vector(mps) - want) < 1e-10
    exact = trot.exact_evolution(trot.make_hamiltonian(5, 1.1), neel, 0.9)
    assert 1.0 - abs(np.vdot(exact, want)) < 1e-3


@pytest.mark.parametrize("ent,m", [("cx", 8), ("cz", 5), ("cp", 8)])
def test_matrix_gradient_equals_parameter_shift(ent, m):
    """
    grad_of_matrix_dot_product against the EXACT parameter-shift gradient of f = <V X, Y>_F, the
    reference's strongest check of the matrix path (test_core_op_matrix.py:115-140, 305-336): every
    rotation angle enters as cos / sin of theta/2 (shift pi, scale 1/4), the CPhase angle as e^{i phi}
    (shift pi/2, scale 1/2).  Nothing but v_mul_mat is used for f.
    """
    rng = np.random.RandomState(len(ent) * 100 + m)
    n = 3
    circ = ParametricCircuit(n, ent, cs.create_ansatz_structure(n, "spin", "full", 5))
    th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    x = np.ascontiguousarray(rng.randn(2**n, m) + 1j * rng.randn(2**n, m))
    y = np.ascontiguousarray(rng.randn(2**n, m) + 1j * rng.randn(2**n, m))

    def fobj(angles):
        return np.vdot(cpm.v_mul_mat(circ, angles, x.copy()), y)

    grad = cpm.grad_of_matrix_dot_product(circ, th, x.copy(), cpm.v_dagger_mul_mat(circ, th, y.copy()))
    tpb = 5 if ent == "cp" else 4
    shifted = np.zeros(circ.num_thetas, dtype=np.complex128)
    for k in range(circ.num_thetas):
        is_phase = ent == "cp" and k >= 3 * n and (k - 3 * n) % tpb == 4
        shift, scale = (0.5 * np.pi, 0.5) if is_phase else (np.pi, 0.25)
        tp, tm = th.copy(), th.copy()
        tp[k] += shift
        tm[k] -= shift
        shifted[k] = scale * (fobj(tp) - fobj(tm))
    assert _rel(grad, shifted) < 1e-10
