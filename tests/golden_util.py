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
