"""
GPU parity of EVERY engine / tile configuration against the CPU oracle:
the configuration a workspace picks depends on the state size (nbits > 21: 3 / 3 forced low bits (3 / 2 at 22),
2^11 / 2^12 tiles), so the production settings of n >= 23 are forced here
at sizes the oracle finishes in seconds, and n = 24 is compared directly with the C oracle.
Tolerance 1e-10 relative, norm-wise (north star).
"""

import numpy as np
import pytest

from aqc_research_b200 import circuit_structures as cs
from aqc_research_b200 import utils
from aqc_research_b200.engine import SvWorkspace
from aqc_research_b200.parametric_circuit import ParametricCircuit, TrotterAnsatz
from oracle import c_oracle as C
from oracle import sv_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-10

LARGE = {"AQC_TILE_LOW_BITS": "3", "AQC_TILE_LOW_BITS_APPLY": "3"}  # what nbits > 22 selects (nbits = 22: 3 / 2)
CONFIGS = {
    "default": {},  # what the workspace picks for the state size (nbits <= 21: 2^10 / 2^11 tiles, 2 / 1 low bits)
    "large-state-tiles": dict(LARGE, AQC_TILE_BITS_GRAD="11", AQC_TILE_BITS_APPLY="12"),
    "256-byte-runs": dict(AQC_TILE_LOW_BITS="4", AQC_TILE_LOW_BITS_APPLY="3", AQC_TILE_BITS_GRAD="11",
                          AQC_TILE_BITS_APPLY="12"),
    "n22-tiles": dict(AQC_TILE_LOW_BITS="3", AQC_TILE_LOW_BITS_APPLY="2", AQC_TILE_BITS_GRAD="11",
                      AQC_TILE_BITS_APPLY="12"),  # what nbits = 22 selects
    "greedy-schedule": {"AQC_TILE_PLAN": "0"},  # the tile planner off
    "plain-launches": {"AQC_PDL": "0"},  # no programmatic dependent launch
    "small-tiles": {"AQC_TILE_BITS_GRAD": "8", "AQC_TILE_BITS_APPLY": "9"},  # generic (not unrolled) tile copies
    "no-fused-steps": {"AQC_DENSE_PAIRS": "0"},
    "legacy": {"AQC_ENGINE": "legacy"},
}


def _rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-300)


def _check(circ, n, seed, oracle=O):
    rng = np.random.RandomState(seed)
    th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    y = rng.rand(2**n) + 1j * rng.rand(2**n)
    y /= np.linalg.norm(y)
    ws = SvWorkspace(circ, num_slots=4)
    try:
        ws.upload(0, y)
        idx = O.basis_state_indices(n)
        hs = ws.objective(th, 0, 1, idx)[0]
        z0 = oracle.apply_v(circ, th, y, dagger=True)
        assert _rel(hs, z0[idx]) < TOL
        assert _rel(ws.download(1), z0) < TOL
        k = int(idx[2])
        g = ws.grad(th, x_basis=k, z0=1, w=2, z=3)[0]
        e = np.zeros(2**n, dtype=np.complex128)
        e[k] = 1
        ref = oracle.grad_sweep(circ, th, e, z0)
        ref = ref[0] if isinstance(ref, tuple) else ref
        assert _rel(g, ref) < TOL
        # forward apply, in place
        ws.apply(th, 0, 0, dagger=False)
        assert _rel(ws.download(0), oracle.apply_v(circ, th, y)) < TOL
    finally:
        ws.close()


@pytest.mark.parametrize("cfg", sorted(CONFIGS))
@pytest.mark.parametrize("n", [13, 16, 20])
def test_every_configuration_vs_oracle(cfg, n, monkeypatch):
    for k, v in CONFIGS[cfg].items():
        monkeypatch.setenv(k, v)
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 1 if n > 13 else 2), True)
    _check(circ, n, 100 + n)
    if n == 13:
        np.random.seed(7)
        for ent in ("cz", "cp"):
            _check(ParametricCircuit(n, ent, utils.rand_circuit(n, 14)), n, 200 + len(ent))


def test_n24_production_configuration_vs_c_oracle():
    """n = 24 (> 2^22 amplitudes: the large-state tile configuration chosen by the workspace itself)."""
    C.use_all_cores()
    n = 24
    circ = TrotterAnsatz(n, cs.make_trotter_like_circuit(n, 1), True)
    _check(circ, n, 24, oracle=C)


@pytest.mark.parametrize("nblocks", [1000, 3300])
def test_very_long_circuits(nblocks):
    """
    The prologue / epilogue kernels keep the (cos, sin) table of all angles in shared memory: more than
    3 072 angles need the opt-in size (> 48 KiB), more than 12 800 fall back to a table in device memory.
    (SURVEY C3 stress case: 7 qubits, 2 870 blocks, 11 501 angles.)
    """
    n = 5
    np.random.seed(nblocks)
    circ = ParametricCircuit(n, "cx", utils.rand_circuit(n, nblocks))
    assert circ.num_thetas * 16 > (48 * 1024 if nblocks == 1000 else 200 * 1024)
    _check(circ, n, 300 + nblocks)
