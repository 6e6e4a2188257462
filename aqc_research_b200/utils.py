"""
Small RNG helpers with the reference's distributions (reference: aqc_research/utils.py:51-79).
They draw from NumPy's GLOBAL generator in the same order as the reference so that fixed seeds
reproduce the reference's test inputs.
"""

import numpy as np


def num_qubits_from_size(size: int) -> int:
    n = int(round(np.log2(float(max(size, 1)))))
    if size != 2**n:
        raise ValueError("'size' argument is not a power of 2 value")
    return n


def rand_circuit(num_qubits: int, depth: int) -> np.ndarray:
    """Random (2, depth) block layout: two distinct qubits per block."""
    assert num_qubits >= 2 and depth >= 0
    cols = np.tile(np.arange(num_qubits).reshape(num_qubits, 1), depth)
    for i in range(depth):
        np.random.shuffle(cols[:, i])
    return cols[0:2, :].copy()


def rand_thetas(num_thetas: int) -> np.ndarray:
    """Angles ~ U(-pi, pi)."""
    assert num_thetas > 0
    return np.pi * (2 * np.random.rand(num_thetas) - 1)


def rand_state(num_qubits: int) -> np.ndarray:
    """Normalised state with re, im ~ U[0, 1)."""
    assert num_qubits >= 2
    dim = 2**num_qubits
    state = np.random.rand(dim) + 1j * np.random.rand(dim)
    state /= np.linalg.norm(state)
    return state


def zero_state(num_qubits: int) -> np.ndarray:
    """The state |0...0> as a dense complex128 vector (utils.py:82-89)."""
    assert isinstance(num_qubits, (int, np.integer)) and num_qubits >= 2
    state = np.zeros(2**num_qubits, dtype=np.complex128)
    state[0] = 1
    return state


def num_cpus() -> int:
    """Number of host CPUs, 1 if it cannot be determined (utils.py:42-48)."""
    import os  # pylint: disable=import-outside-toplevel

    count = os.cpu_count()
    return int(count) if isinstance(count, int) and count > 0 else 1
