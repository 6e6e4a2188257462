"""Helpers to read the golden fixtures (tests/golden/*.npz) and rebuild their circuits."""

import os
import numpy as np
from aqc_research_b200.parametric_circuit import ParametricCircuit, TrotterAnsatz

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KINDS = ["cx", "cz", "cp", "trotter1", "trotter2"]


def load(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name))


def circuit_from(kind: str, n: int, blocks: np.ndarray):
    if kind == "trotter1":
        return TrotterAnsatz(n, blocks, False)
    if kind == "trotter2":
        return TrotterAnsatz(n, blocks, True)
    return ParametricCircuit(n, kind, blocks)


def rel(a, b) -> float:
    """Norm-wise relative difference (reference: test/utils_for_testing.py:23-44)."""
    a, b = np.ravel(a), np.ravel(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def mps_from_golden(g, prefix: str, tag: str):
    """Rebuilds a QiskitMPS tuple stored by make_golden.mps_cases()."""
    n = int(g[prefix + "n"])
    gam = [(g[f"{prefix}{tag}_g0_{k}"], g[f"{prefix}{tag}_g1_{k}"]) for k in range(n)]
    lam = [g[f"{prefix}{tag}_l_{k}"] for k in range(n - 1)]
    return gam, lam
