"""
Generates the golden fixtures ``tests/golden/*.npz`` by running the UNMODIFIED reference
(/root/reference, imported read-only through ref_loader.py) on seeded inputs.
Build-container only:  python tests/golden/make_golden.py
The fixtures hold inputs AND reference outputs, so tests need neither the reference nor the
global NumPy RNG stream to reproduce them.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

R = load_reference()


def make_circuit(kind, n, depth_or_layers):
    if kind in ("trotter1", "trotter2"):
        blocks = R.cs.make_trotter_like_circuit(n, depth_or_layers)
        return R.pc.TrotterAnsatz(n, blocks, kind == "trotter2"), blocks
    blocks = R.utils.rand_circuit(n, depth_or_layers)
    return R.pc.ParametricCircuit(n, kind, blocks), blocks


def sv_cases():
    """v_mul_vec, v_dagger_mul_vec, grad_of_dot_product (full and partial)."""
    out = {}
    case = 0
    np.random.seed(0x0696969)  # seed used by the reference's own tests
    for n in (2, 3, 4, 5, 6):
        for kind in ("cx", "cz", "cp", "trotter1", "trotter2"):
            size = 2 if kind.startswith("trotter") else 8
            circ, blocks = make_circuit(kind, n, size)
            th = R.utils.rand_thetas(circ.num_thetas)
            x, y = R.utils.rand_state(n), R.utils.rand_state(n)
            ws = np.zeros((3, 2**n), dtype=np.complex128)
            buf = np.zeros(2**n, dtype=np.complex128)
            vy = R.cop.v_mul_vec(circ, th, y, buf, ws).copy()
            vhy = R.cop.v_dagger_mul_vec(circ, th, y, buf, ws).copy()
            grad = R.cop.grad_of_dot_product(circ, th, x, vhy, ws)
            br = (1, circ.num_blocks - 1)
            gpart = R.cop.grad_of_dot_product(circ, th, x, vhy, ws, block_range=br, front_layer=False)
            pre = f"c{case}_"
            out[pre + "meta"] = np.array([n, ["cx", "cz", "cp", "trotter1", "trotter2"].index(kind), br[0], br[1]])
            out[pre + "blocks"] = blocks
            out[pre + "thetas"], out[pre + "x"], out[pre + "y"] = th, x, y
            out[pre + "v_y"], out[pre + "vh_y"], out[pre + "grad"], out[pre + "grad_part"] = vy, vhy, grad, gpart
            case += 1
    out["num_cases"] = np.array(case)
    np.savez_compressed(os.path.join(HERE, "sv_cases.npz"), **out)
    print("sv_cases:", case)


def mat_cases():
    """v_mul_mat, v_dagger_mul_mat, grad_of_matrix_dot_product on rectangular matrices."""
    out = {}
    case = 0
    np.random.seed(0x696969)
    for n, m in ((2, 3), (3, 8), (4, 5), (5, 16)):
        for kind in ("cx", "cz", "cp"):
            circ, blocks = make_circuit(kind, n, 10)
            th = R.utils.rand_thetas(circ.num_thetas)
            X = np.random.rand(2**n, m) + 1j * np.random.rand(2**n, m)
            Y = np.random.rand(2**n, m) + 1j * np.random.rand(2**n, m)
            ws = np.zeros_like(X)
            vy = R.cpm.v_mul_mat(circ, th, Y.copy(), ws).copy()
            vhy = R.cpm.v_dagger_mul_mat(circ, th, Y.copy(), ws).copy()
            grad = R.cpm.grad_of_matrix_dot_product(circ, th, X.copy(), vhy.copy(), ws)
            pre = f"c{case}_"
            out[pre + "meta"] = np.array([n, ["cx", "cz", "cp"].index(kind), m])
            out[pre + "blocks"] = blocks
            out[pre + "thetas"], out[pre + "x"], out[pre + "y"] = th, X, Y
            out[pre + "v_y"], out[pre + "vh_y"], out[pre + "grad"] = vy, vhy, grad
            case += 1
    out["num_cases"] = np.array(case)
    np.savez_compressed(os.path.join(HERE, "mat_cases.npz"), **out)
    print("mat_cases:", case)


def objective_sequences():
    """
    Stateful call sequences of SpSurrogateObjectiveMax (ThinStateHandler) and of
    SketchingObjectiveEx + FullRangeSketchingVectors: objective(), gradient() at several thetas.
    Includes the SURVEY section 8(c) sanity configuration (seed 1234, n = 5 and 12).
    """
    out = {}
    case = 0
    for n, layers, seed, steps in ((5, 2, 1234, 4), (12, 2, 1234, 2), (4, 1, 77, 5), (7, 3, 5, 3)):
        np.random.seed(seed)
        target = R.utils.rand_state(n)
        blocks = R.cs.make_trotter_like_circuit(n, layers)
        circ = R.pc.TrotterAnsatz(n, blocks, True)
        th = R.utils.rand_thetas(circ.num_thetas)
        params = dict(num_qubits=n, max_flips=1, maxiter=10, verbose=0, enable_optim_stats=False,
                      num_simulations=1, trunc_thr=1e-6, state_prep_func=None)
        objv = R.sur_max.SpSurrogateObjectiveMax(user_parameters=params, circ=circ, front_layer=True)
        objv.set_target(target)
        ths, fs, gs, ws_, mx, hs = [], [], [], [], [], []
        for s in range(steps):
            f = objv.objective(th)
            hs.append(objv._hs.copy())
            mx.append(objv._max_no)
            g = objv.gradient(th)
            ths.append(th.copy()); fs.append(f); gs.append(g.copy()); ws_.append(objv._weight)
            th = th - 0.3 * g + 0.01 * np.cos(np.arange(th.size) + s)  # deterministic next point
        pre = f"sp{case}_"
        out[pre + "meta"] = np.array([n, layers, steps])
        out[pre + "blocks"], out[pre + "target"] = blocks, target
        out[pre + "thetas"], out[pre + "f"], out[pre + "grad"] = np.array(ths), np.array(fs), np.array(gs)
        out[pre + "weight"], out[pre + "max_no"], out[pre + "hs"] = np.array(ws_), np.array(mx), np.array(hs)
        case += 1
    out["num_sp"] = np.array(case)

    case = 0
    from scipy.stats import unitary_group
    for n, depth, ent, seed in ((3, 12, "cx", 11), (4, 20, "cz", 12), (5, 25, "cp", 13), (5, 30, "cx", 14)):
        np.random.seed(seed)
        U = unitary_group.rvs(2**n, random_state=seed).astype(np.complex128)
        blocks = R.cs.create_ansatz_structure(n, "cyclic_spin", "full", depth)
        circ = R.pc.ParametricCircuit(n, ent, blocks)
        objv = R.sk_core.SketchingObjectiveEx(circ, R.sk_core.FullRangeSketchingVectors(U))
        th = R.utils.rand_thetas(circ.num_thetas)
        ths, fs, gs = [], [], []
        for s in range(3):
            f = objv.objective(th)
            g = objv.gradient(th)
            ths.append(th.copy()); fs.append(f); gs.append(g.copy())
            th = th - 0.5 * g
        pre = f"sk{case}_"
        out[pre + "meta"] = np.array([n, ["cx", "cz", "cp"].index(ent)])
        out[pre + "blocks"], out[pre + "target"] = blocks, U
        out[pre + "thetas"], out[pre + "f"], out[pre + "grad"] = np.array(ths), np.array(fs), np.array(gs)
        case += 1
    out["num_sk"] = np.array(case)
    np.savez_compressed(os.path.join(HERE, "objective_sequences.npz"), **out)
    print("objective sequences written")


def mps_cases():
    """mps_to_vector / mps_dot of the reference (pure NumPy) on fixed MPS tuples (Vidal form)."""
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import mps_oracle as M  # only used to GENERATE input tuples (format), not outputs

    out = {}
    rng = np.random.RandomState(2718)
    case = 0
    for n, chi in ((2, 2), (3, 2), (5, 4), (7, 3), (9, 8)):
        m1, m2 = M.random_mps(n, chi, rng), M.random_mps(n, chi, rng)
        assert R.mpsop.check_mps(m1) and R.mpsop.check_mps(m2)
        pre = f"m{case}_"
        out[pre + "n"] = np.array(n)
        for tag, m in (("a", m1), ("b", m2)):
            for k in range(n):
                out[pre + f"{tag}_g0_{k}"], out[pre + f"{tag}_g1_{k}"] = m[0][k]
                if k < n - 1:
                    out[pre + f"{tag}_l_{k}"] = m[1][k]
        out[pre + "vec_a"] = R.mpsop.mps_to_vector(m1)
        out[pre + "vec_b"] = R.mpsop.mps_to_vector(m2)
        out[pre + "dot_ab"] = np.array(R.mpsop.mps_dot(m1, m2))
        out[pre + "dot_aa"] = np.array(R.mpsop.mps_dot(m1, m1))
        case += 1
    out["num_cases"] = np.array(case)
    np.savez_compressed(os.path.join(HERE, "mps_cases.npz"), **out)
    print("mps_cases:", case)


def trotter_and_lbfgs():
    """
    init_ansatz_to_trotter angles of the reference, and a full SciPy L-BFGS-B run on the
    reference's SpSurrogateObjectiveMax (config C1 of SURVEY 8(d): n = 5, 2nd-order TrotterAnsatz,
    Neel initial state, Trotter target with 10x finer steps, theta_0 = init_ansatz_to_trotter).
    The reference builds its Neel-state handler through Qiskit (absent here); for an X-only
    preparation the states are basis vectors, so a duck-typed handler with the same indices is
    swapped in.
    """
    from scipy.optimize import minimize

    out = {}
    case = 0
    for n, layers, so, t_evol in ((4, 2, True, 0.8), (5, 3, False, 1.0), (6, 2, True, 1.2)):
        circ = R.pc.TrotterAnsatz(n, R.cs.make_trotter_like_circuit(n, layers), so)
        th = R.trotter.init_ansatz_to_trotter(circ, np.full(circ.num_thetas, 0.123), evol_time=t_evol, delta=1.0)
        out[f"ia{case}_meta"] = np.array([n, layers, int(so)])
        out[f"ia{case}_time"] = np.array(t_evol)
        out[f"ia{case}_thetas"] = th
        v = np.zeros(2**n, dtype=np.complex128)
        v[sum(1 << q for q in range(0, n, 2))] = 1
        out[f"ia{case}_state"] = R.cop.v_mul_vec(circ, th, v, np.zeros_like(v), np.zeros((2, 2**n), dtype=np.complex128))
        case += 1
    out["num_ia"] = np.array(case)

    n, layers, t_evol, delta = 5, 2, 1.2, 1.0
    neel = sum(1 << q for q in range(0, n, 2))
    fine = R.pc.TrotterAnsatz(n, R.cs.make_trotter_like_circuit(n, 10 * layers), True)
    th_f = R.trotter.init_ansatz_to_trotter(fine, np.zeros(fine.num_thetas), evol_time=t_evol, delta=delta)
    e = np.zeros(2**n, dtype=np.complex128)
    e[neel] = 1
    target = R.cop.v_mul_vec(fine, th_f, e, np.zeros_like(e), np.zeros((2, 2**n), dtype=np.complex128)).copy()
    circ = R.pc.TrotterAnsatz(n, R.cs.make_trotter_like_circuit(n, layers), True)
    th0 = R.trotter.init_ansatz_to_trotter(circ, np.zeros(circ.num_thetas), evol_time=t_evol, delta=delta)
    params = dict(num_qubits=n, max_flips=1, maxiter=40, verbose=0, enable_optim_stats=False,
                  num_simulations=1, trunc_thr=1e-6, state_prep_func=None)
    objv = R.sur_max.SpSurrogateObjectiveMax(user_parameters=params, circ=circ, front_layer=True)

    class NeelHandler:
        num_states = n + 1
        idx = [neel] + [neel ^ (1 << q) for q in range(n)]

        def init_state(self, i):
            v = np.zeros(2**n, dtype=np.complex128)
            v[self.idx[i]] = 1
            return v

        @property
        def state0(self):
            return self.init_state(0)

        def state_dot_vector(self, i, vec):
            return vec[self.idx[i]]

    objv._state_handler = NeelHandler()
    objv.set_target(target)
    maxiter = 12
    res = minimize(fun=objv.objective, x0=th0.copy(), jac=objv.gradient, method="L-BFGS-B",
                   options=dict(maxfun=5 * maxiter, maxiter=maxiter, ftol=10 * np.finfo(float).eps, eps=1e-8))
    out["lb_meta"] = np.array([n, layers, maxiter, res.nit, res.nfev])
    out["lb_time"] = np.array(t_evol)
    out["lb_target"], out["lb_theta0"], out["lb_x"] = target, th0, res.x
    out["lb_fun"], out["lb_fidelity"] = np.array(res.fun), np.array(objv.fidelity)
    np.savez_compressed(os.path.join(HERE, "trotter_lbfgs.npz"), **out)
    print("trotter + lbfgs: nit", res.nit, "nfev", res.nfev, "fun", res.fun, "fidelity", objv.fidelity)


def cd_cases():
    """coord_descent_single_sweep (core_op_matrix.py:765-917): several consecutive sweeps."""
    from scipy.stats import unitary_group

    out = {}
    case = 0
    np.random.seed(0x0C0D)
    for n, nblocks in ((2, 3), (3, 7), (4, 12), (5, 20)):
        for kind in ("cx", "cz"):
            circ, blocks = make_circuit(kind, n, nblocks)
            target = unitary_group.rvs(2**n, random_state=100 * n + len(kind) + case).astype(np.complex128)
            th = R.utils.rand_thetas(circ.num_thetas)
            ws = np.zeros((3, 2**n, 2**n), dtype=np.complex128)
            ths, fs = [th.copy()], []
            for _ in range(4):
                f = R.cpm.coord_descent_single_sweep(circ, th, target, ws)
                ths.append(th.copy())
                fs.append(f)
            pre = f"c{case}_"
            out[pre + "meta"] = np.array([n, ["cx", "cz"].index(kind)])
            out[pre + "blocks"] = blocks
            out[pre + "target"] = target
            out[pre + "thetas"] = np.array(ths)  # [5][T]: start + after each sweep
            out[pre + "fobj"] = np.array(fs)
            case += 1
    out["num_cases"] = np.array(case)
    np.savez_compressed(os.path.join(HERE, "cd_cases.npz"), **out)
    print("cd_cases:", case)


def sketch_cases():
    """rand / alt / eigen sketching generators + SketchingObjectiveEx (sk_core.py:167-222, 329-464)."""
    from scipy.stats import unitary_group

    out = {}
    case = 0
    for n, m, kind, ent in ((3, 2, "rand", "cx"), (4, 4, "alt", "cz"), (4, 8, "eigen", "cx"), (5, 4, "rand", "cp"),
                            (5, 8, "alt", "cx"), (5, 16, "eigen", "cz"), (6, 8, "eigen", "cx")):
        seed = 4242 + case
        np.random.seed(seed)
        circ, blocks = make_circuit(ent, n, 3 * n)
        target = unitary_group.rvs(2**n, random_state=seed).astype(np.complex128)
        ths = [R.utils.rand_thetas(circ.num_thetas) for _ in range(3)]
        np.random.seed(seed + 1)  # the stream the generator sees (the test re-seeds the same way)
        gen = R.sk_core.skvecs_generator(kind, m, target)
        objv = R.sk_core.SketchingObjectiveEx(circ, gen)
        fs, gs = [], []
        for th in ths:
            f, g = objv.objective_and_gradient(th)
            fs.append(f)
            gs.append(g.copy())
        pre = f"c{case}_"
        out[pre + "meta"] = np.array([n, m, ["rand", "alt", "eigen"].index(kind), ["cx", "cz", "cp"].index(ent), seed + 1])
        out[pre + "blocks"] = blocks
        out[pre + "target"] = target
        out[pre + "thetas"] = np.array(ths)
        out[pre + "f"] = np.array(fs)
        out[pre + "grad"] = np.array(gs)
        case += 1
    out["num_cases"] = np.array(case)
    np.savez_compressed(os.path.join(HERE, "sketch_cases.npz"), **out)
    print("sketch_cases:", case)


def primitive_cases():
    """
    Gate-by-gate primitives of the reference (core_operations.py:46-603, core_op_matrix.py:32-477,
    elementary_operations.py:39-291) on seeded inputs: every function, every qubit position.
    """
    cop, cpm, eo = R.cop, R.cpm, R.eo
    rng = np.random.RandomState(0xA11)
    out = {}
    n = 4
    dim = 2**n

    def rvec():
        return (rng.randn(dim) + 1j * rng.randn(dim)).astype(np.complex128)

    g = (rng.randn(2, 2) + 1j * rng.randn(2, 2)).astype(np.complex128)
    cm, tm, gm = [(rng.randn(2, 2) + 1j * rng.randn(2, 2)).astype(np.complex128) for _ in range(3)]
    ang = 0.7321
    out["n"], out["gate"], out["angle"] = np.array(n), g, np.array(ang)
    out["c_mat"], out["t_mat"], out["g_mat"] = cm, tm, gm
    v0, z0 = rvec(), rvec()
    out["vec"], out["zvec"] = v0, z0
    tmp = np.zeros(dim, dtype=np.complex128)
    for pos in range(n):
        out[f"gate2x2_{pos}"] = cop.gate2x2_mul_vec(n, pos, g, v0.copy(), tmp.copy(), True).copy()
        o = np.zeros(dim, dtype=np.complex128)
        cop.gate2x2_mul_vec(n, pos, g, v0.copy(), o, False)
        out[f"gate2x2_out_{pos}"] = o
        out[f"proj00_{pos}"] = cop.proj00_mul_vec(n, pos, v0.copy()).copy()
        out[f"proj11_{pos}"] = cop.proj11_mul_vec(n, pos, v0.copy()).copy()
        for nm in ("rx", "ry", "rz"):
            out[f"{nm}_{pos}"] = getattr(cop, nm + "_mul_vec")(n, pos, ang, v0.copy(), tmp.copy()).copy()
        for nm in ("dot_x", "dot_y", "dot_z"):
            out[f"{nm}_{pos}"] = np.array(getattr(cop, nm)(n, pos, v0.copy(), z0.copy(), tmp.copy()))
    for c in range(n):
        for t in range(n):
            if c == t:
                continue
            for nm in ("cx", "cz", "cp"):
                out[f"{nm}_{c}{t}"] = getattr(cop, nm + "_mul_vec")(n, c, t, ang, v0.copy(), tmp.copy()).copy()
            o = np.zeros(dim, dtype=np.complex128)
            out[f"dcp_{c}{t}"] = cop.derv_cphase_mul_vec(n, c, t, ang, v0.copy(), o).copy()
            for dag in (False, True):
                ws = np.zeros((2, dim), dtype=np.complex128)
                out[f"block_{c}{t}_{int(dag)}"] = cop.block_mul_vec(n, c, t, cm, tm, gm, v0.copy(), ws, dag).copy()
            out[f"np_block_{c}{t}"] = eo.np_block_matrix(n, c, t, cm, tm, gm)
            out[f"np_cx_{c}{t}"] = eo.np_cx_matrix(n, c, t)
    for nm in ("np_rx", "np_ry", "np_rz", "np_phase"):
        out[nm] = getattr(eo, nm)(ang)
    out["np_x"], out["np_z"] = eo.np_x(), eo.np_z()
    # matrices: square and ragged (m = 5 columns)
    for m in (dim, 5):
        m0 = (rng.randn(dim, m) + 1j * rng.randn(dim, m)).astype(np.complex128)
        zm = (rng.randn(dim, m) + 1j * rng.randn(dim, m)).astype(np.complex128)
        out[f"mat_{m}"], out[f"zmat_{m}"] = m0, zm
        ws = np.zeros(dim * m, dtype=np.complex128)
        for q in range(n):
            out[f"m{m}_gate2x2_{q}"] = cpm.gate2x2_mul_mat(q, g, m0.copy(), ws.copy()).copy()
            for nm in ("rx", "ry", "rz"):
                out[f"m{m}_{nm}_{q}"] = getattr(cpm, nm + "_mul_mat")(ang, q, m0.copy(), ws.copy()).copy()
            for nm in ("x", "y", "z"):
                out[f"m{m}_{nm}dot_{q}"] = np.array(getattr(cpm, nm + "_dot_mat")(q, m0.copy(), zm.copy(), ws.copy()))
        for c in range(n):
            for t in range(n):
                if c == t:
                    continue
                for nm in ("cx", "cz", "cp"):
                    out[f"m{m}_{nm}_{c}{t}"] = getattr(cpm, nm + "_mul_mat")(c, t, ang, m0.copy(), ws.copy()).copy()
                out[f"m{m}_dcp_{c}{t}"] = np.array(cpm.derv_cphase(c, t, m0.copy(), zm.copy(), ws.copy()))
    np.savez_compressed(os.path.join(HERE, "primitive_cases.npz"), **out)
    print("primitive_cases:", len(out))


if __name__ == "__main__":
    if len(sys.argv) > 1:  # e.g. `make_golden.py primitive_cases`
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    primitive_cases()
    sv_cases()
    mat_cases()
    objective_sequences()
    mps_cases()
    trotter_and_lbfgs()
    cd_cases()
    sketch_cases()
