"""
Unit-block layouts (host-side index generation only).
Reference: aqc_research/circuit_structures.py:31-349 -- same function names and outputs.
"""

from typing import List
import numpy as np

_LAYOUTS = ("spin", "line", "cyclic_spin", "cyclic_line")
_CONNECTIVITY = ("full", "line")


def circuit_layout_list() -> List[str]:
    return list(_LAYOUTS)


def circuit_connectivity_list() -> List[str]:
    return list(_CONNECTIVITY)


def lower_limit(num_qubits: int) -> int:
    """Number of unit-blocks that guarantees exact compilation: ceil((4^n - 3n - 1) / 4)."""
    return int(-((-(4**num_qubits - 3 * num_qubits - 1)) // 4))


def num_blocks_per_layer(num_qubits: int, circuit_layout: str) -> int:
    assert circuit_layout in _LAYOUTS and num_qubits >= 2
    return num_qubits if circuit_layout.startswith("cyclic_") else num_qubits - 1


def fraction_of_lower_bound(depth_fraction: float, num_qubits: int, circuit_layout: str) -> int:
    """Number of layers whose total depth is ``depth_fraction`` of ``lower_limit`` (ceil)."""
    if circuit_layout not in _LAYOUTS:
        raise ValueError(f"'circuit_layout' must be one of {list(_LAYOUTS)}")
    if not 0 < depth_fraction <= 1:
        raise ValueError("expects: 0 < depth_fraction <= 1")
    bpl = num_blocks_per_layer(num_qubits, circuit_layout)
    depth = int(round(depth_fraction * lower_limit(num_qubits)))
    return int(max(1, -(-depth // bpl)))


def _pairs_spin(n: int):
    """Endless brick-wall sequence (0,1),(2,3),...,(1,2),(3,4),..."""
    while True:
        for start in (0, 1):
            for k in range(start, n - 1, 2):
                yield k, k + 1


def _layout_pairs(n: int, depth: int, layout: str) -> np.ndarray:
    out = np.zeros((2, depth), dtype=int)
    if layout == "spin":
        gen = _pairs_spin(n)
        for i in range(depth):
            out[:, i] = next(gen)
    elif layout == "line":
        pos = 0
        for i in range(depth):
            if pos % n == n - 1:  # never connect last and first qubits
                pos += 1
            out[:, i] = pos % n, (pos + 1) % n
            pos += 1
    elif layout == "cyclic_spin":
        even = n % 2 == 0
        for i in range(depth):
            off = (i // (n // 2)) % 2 if even else 0
            out[:, i] = (2 * i + off) % n, (2 * i + off + 1) % n
    elif layout == "cyclic_line":
        idx = np.arange(depth)
        out[0], out[1] = idx % n, (idx + 1) % n
    else:
        raise ValueError(f"Unknown type of circuit layout, expects one of {list(_LAYOUTS)}, got {layout}")
    return out


def create_ansatz_structure(
    num_qubits: int,
    layout: str = "spin",
    connectivity: str = "full",
    depth: int = 0,
    block_repeat: int = 1,
    logger=None,
) -> np.ndarray:
    """(2, depth*block_repeat) array: row 0 control qubits, row 1 target qubits."""
    if num_qubits < 2:
        raise ValueError("Number of qubits must be greater or equal to 2")
    if layout not in _LAYOUTS:
        raise ValueError(f"Unknown type of circuit layout, expects one of {list(_LAYOUTS)}, got {layout}")
    if connectivity not in _CONNECTIVITY:
        raise ValueError(f"layout '{layout}' assumes 'line' or 'full' connectivity, got {connectivity}")
    if not 1 <= block_repeat <= 3:
        raise ValueError("'block_repeat' argument must be equal 1, 2 or 3")
    if depth <= 0:
        depth = lower_limit(num_qubits)
        if logger:
            logger.warning(f"choosing the maximum number of 2-qubit unit blocks: {depth}")
    blocks = _layout_pairs(num_qubits, depth, layout)
    if block_repeat > 1:
        blocks = np.repeat(blocks, block_repeat, axis=1)
    return blocks


def make_trotter_like_circuit(
    num_qubits: int, num_layers: int, *, connectivity: str = "full", verbose: bool = False
) -> np.ndarray:
    """
    ``num_layers`` brick-wall layers of triplets: pair (k, k+1) gives blocks
    (k+1 -> k), (k -> k+1), (k+1 -> k).
    """
    if num_qubits < 2:
        raise ValueError("number of qubits must be greater or equal to 2")
    if connectivity not in _CONNECTIVITY:
        raise ValueError("expects 'full' or 'line' connectivity")
    if num_layers < 0:
        raise ValueError("expects non-negative number of layers")
    if num_layers == 0:
        return np.zeros((2, 0), dtype=int)
    pairs = _layout_pairs(num_qubits, num_layers * (num_qubits - 1), "spin")
    lo, hi = pairs[0], pairs[1]
    blocks = np.empty((2, 3 * pairs.shape[1]), dtype=int)
    blocks[0, 0::3], blocks[1, 0::3] = hi, lo
    blocks[0, 1::3], blocks[1, 1::3] = lo, hi
    blocks[0, 2::3], blocks[1, 2::3] = hi, lo
    return blocks
