"""
GPU parity of the sketching-vector generators (rand / alt / eigen; csrc/aqc_sketch.cuh: DMMA GEMM,
Cholesky-QR, column gathers) driving SketchingObjectiveEx, against golden outputs of the
reference classes (tests/golden/sketch_cases.npz) and against the oracle; through the C-ABI.
"""

import numpy as np
import pytest

from golden_util import load, rel
from oracle import sv_oracle as O
from aqc_research_b200 import circuit_structures as cs
from aqc_research_b200.engine import SvWorkspace
from aqc_research_b200.model_sketching import sk_core
from aqc_research_b200.parametric_circuit import ParametricCircuit

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_generators_golden():
    g = load("sketch_cases.npz")
    for c in range(int(g["num_cases"])):
        p = f"c{c}_"
        n, m, kind, ent, seed = [int(v) for v in g[p + "meta"]]
        circ = ParametricCircuit(n, ["cx", "cz", "cp"][ent], g[p + "blocks"])
        np.random.seed(seed)  # the generators draw from the global RNG like the reference
        gen = sk_core.skvecs_generator(["rand", "alt", "eigen"][kind], m, g[p + "target"])
        objv = sk_core.SketchingObjectiveEx(circ, gen)
        for s, th in enumerate(g[p + "thetas"]):
            f, grad = objv.objective_and_gradient(th)
            assert abs(f - g[p + "f"][s]) < TOL, (c, s, f, g[p + "f"][s])
            assert rel(grad, g[p + "grad"][s]) < TOL, (c, s)


def test_gemm_and_orthonormalize_properties():
    """U X, U^H X against NumPy; orthonormalised columns: X^H X = I and span(X) = span(A)."""
    n, m = 8, 16
    d = 1 << n
    rng = np.random.RandomState(5)
    q, r = np.linalg.qr(rng.randn(d, d) + 1j * rng.randn(d, d))
    u = np.ascontiguousarray(q * (np.diag(r) / np.abs(np.diag(r))))
    circ = ParametricCircuit(n, "cx", cs.create_ansatz_structure(n, "spin", "full", 4))
    ws = SvWorkspace(circ, num_slots=4, log2_cols=4, as_generic=True)
    ws.set_dense_target(u)
    a = rng.rand(d, m) + 1j * rng.rand(d, m)
    # ill-conditioned on purpose: two nearly parallel columns (cond ~ 1e9)
    a[:, 3] = a[:, 2] + 1e-9 * a[:, 3]
    ws.upload(0, a)
    ws.target_matmul(0, 1)
    assert rel(ws.download(1).reshape(d, m), u @ a) < 1e-13
    ws.target_matmul(0, 1, conj_transpose=True)
    assert rel(ws.download(1).reshape(d, m), u.conj().T @ a) < 1e-13
    ws.orthonormalize(0, 1)
    x = ws.download(0).reshape(d, m)
    assert np.linalg.norm(x.conj().T @ x - np.eye(m)) < 1e-13
    qa, _ = np.linalg.qr(a)
    assert np.linalg.norm(x - qa @ (qa.conj().T @ x)) < 1e-6  # same span (limited by cond(A) * eps)
    ws.close()


def test_standalone_generate_and_factory():
    n, m = 4, 4
    rng = np.random.RandomState(9)
    q, _ = np.linalg.qr(rng.randn(16, 16) + 1j * rng.randn(16, 16))
    np.random.seed(1)
    gen = sk_core.skvecs_generator("rand", m, q)
    x, y = gen.generate()
    assert np.linalg.norm(x.conj().T @ x - np.eye(m)) < 1e-13 and rel(y, q @ x) < 1e-13
    np.random.seed(1)
    x_ref, _ = np.linalg.qr(np.random.rand(16, m) + 1j * np.random.rand(16, m))
    assert np.linalg.norm(x - x_ref @ (x_ref.conj().T @ x)) < 1e-12
    assert isinstance(sk_core.skvecs_generator("alt", 16, q), sk_core.FullRangeSketchingVectors)
    with pytest.raises(ValueError):
        sk_core.skvecs_generator("nope", m, q)
