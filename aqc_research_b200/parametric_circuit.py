"""
Host-side circuit description with the reference's public surface
(reference: aqc_research/parametric_circuit.py:24-466).  Pure Python/NumPy: the object
only describes WHERE unit-blocks sit; all arithmetic happens in the CUDA engine, which
receives the structure through ``aqc_circuit_create`` (include/aqc_b200.h).

Unit block (control c, target t), angles th[0..tpb-1]:
    E(c->t) ; Ry_c(th0) Rz_c(th1) ; Ry_t(th2) Rs_t(th3)       Rs = Rx for cx, Rz for cz/cp
theta vector = [3 per qubit front layer | tpb per block], tpb = 5 for "cp" (th4 = phase).
"""

from typing import Optional, Tuple, Union
import numpy as np
from . import checking as chk

_ENTANGLERS = ("cx", "cz", "cp")


class ParametricCircuit:
    """Generic ansatz: front layer of Rz-Ry-Rz gates followed by 2-qubit unit-blocks."""

    def __init__(
        self,
        num_qubits: int,
        entangler: str,
        blocks: np.ndarray,
        name: Optional[str] = None,
        power: Optional[int] = 1,
    ):
        if entangler not in _ENTANGLERS:
            raise ValueError(f"entangler must be one of {_ENTANGLERS}, got {entangler}")
        self.check_block_layout(num_qubits, blocks)
        if not chk.is_int(power, power >= 1):
            raise ValueError("expects circuit power (V^p) to be integer and p >= 1")
        self._n = int(num_qubits)
        self._entangler = entangler
        self._blocks = np.array(blocks, dtype=int)
        self._name = name if isinstance(name, str) else ""
        self._power = int(power)

    # -- structure -------------------------------------------------------------------
    def update_structure(self, blocks: np.ndarray):
        """Replaces the block layout (angles are owned by the caller)."""
        self.check_block_layout(self._n, blocks)
        self._blocks = np.array(blocks, dtype=int)

    def check_block_layout(self, num_qubits: int, blocks: np.ndarray):
        """Raises ValueError unless ``blocks`` is a valid generic layout."""
        if not chk.block_structure(num_qubits, blocks):
            raise ValueError("not a valid structure of unit-blocks")

    @property
    def name(self) -> str:
        return self._name

    @property
    def num_qubits(self) -> int:
        return self._n

    @property
    def dimension(self) -> int:
        return 1 << self._n

    @property
    def num_blocks(self) -> int:
        return int(self._blocks.shape[1])

    @property
    def tpb(self) -> int:
        """Angles per unit-block."""
        return 5 if self._entangler == "cp" else 4

    @property
    def num_thetas(self) -> int:
        return 3 * self._n + self.tpb * self.num_blocks

    @property
    def blocks(self) -> np.ndarray:
        return self._blocks

    @property
    def entangler(self) -> str:
        return self._entangler

    @property
    def circuit_power(self) -> int:
        return self._power

    @property
    def num_layers(self) -> int:
        raise NotImplementedError("there are no layers in generic ansatz")

    @property
    def bpl(self) -> int:
        raise NotImplementedError("there are no layers in generic ansatz")

    # -- views on parameter / gradient vectors ------------------------------------------
    def subset1q(self, vec: np.ndarray) -> np.ndarray:
        """View (num_qubits, 3) of the front-layer entries of a theta-sized vector."""
        assert isinstance(vec, np.ndarray) and vec.shape == (self.num_thetas,)
        return vec[: 3 * self._n].reshape(-1, 3)

    def subset2q(self, vec: np.ndarray) -> np.ndarray:
        """View (num_blocks, tpb) of the unit-block entries of a theta-sized vector."""
        assert isinstance(vec, np.ndarray) and vec.shape == (self.num_thetas,)
        return vec[3 * self._n :].reshape(-1, self.tpb)

    def insert_unit_blocks(
        self,
        pos: int,
        extra_blocks: np.ndarray,
        thetas: Optional[np.ndarray] = None,
    ) -> Union[Tuple[np.ndarray, np.ndarray], Tuple[None, None]]:
        """
        Inserts ``extra_blocks`` before block ``pos`` (append if pos == num_blocks).  If
        ``thetas`` is given, returns (thetas expanded with zeros at the new entries, indices of
        those entries); otherwise (None, None).
        """
        self.check_block_layout(self._n, extra_blocks)
        assert chk.is_int(pos, 0 <= pos <= self.num_blocks)
        assert thetas is None or chk.float_1d(thetas, thetas.size == self.num_thetas)
        first = 3 * self._n + pos * self.tpb
        count = self.tpb * extra_blocks.shape[1]
        self._blocks = np.concatenate(
            (self._blocks[:, :pos], np.asarray(extra_blocks, dtype=int), self._blocks[:, pos:]),
            axis=1,
        )
        if thetas is None:
            return None, None
        grown = np.concatenate((thetas[:first], np.zeros(count, thetas.dtype), thetas[first:]))
        assert grown.size == self.num_thetas
        return grown, np.arange(first, first + count, dtype=int)


class TrotterAnsatz(ParametricCircuit):
    """
    Layers of block triplets on adjacent qubit pairs (cx entangler).  Triplet on (k, k+1):
    blocks (k+1 -> k), (k -> k+1), (k+1 -> k).  With ``second_order`` a trailing half-layer
    of 3*(n//2) blocks re-using the angles of the leading half-layer is IMPLIED (it is not
    stored in ``blocks``; the numeric core appends it and sums its derivatives into the
    leading half-layer's entries).
    """

    def __init__(
        self, num_qubits: int, blocks: np.ndarray, second_order: bool, name: Optional[str] = None
    ):
        assert isinstance(second_order, bool)
        self._second_order = second_order
        super().__init__(num_qubits, "cx", blocks, name)

    @property
    def is_second_order(self) -> bool:
        return self._second_order

    @property
    def half_layer_num_blocks(self) -> int:
        return 3 * (self.num_qubits // 2) if self._second_order else 0

    @property
    def bpl(self) -> int:
        return 3 * (self.num_qubits - 1)

    @property
    def num_layers(self) -> int:
        return self.num_blocks // self.bpl

    def insert_unit_blocks(self, pos, extra_blocks, thetas=None):
        assert chk.is_int(pos, 0 <= pos <= self.num_blocks)
        if pos % (3 * (self.num_qubits - 1)) != 0:
            raise ValueError("position of blocks insertion must be aligned at layer boundary")
        return super().insert_unit_blocks(pos, extra_blocks, thetas)

    def check_block_layout(self, num_qubits: int, blocks: np.ndarray):
        super().check_block_layout(num_qubits, blocks)
        nb = blocks.shape[1]
        if nb == 0:
            return
        if nb % (3 * (num_qubits - 1)) != 0:
            raise ValueError("not a valid Trotterized block layout")
        first, mid, last = blocks[:, 0::3], blocks[:, 1::3], blocks[:, 2::3]
        ok = (
            np.array_equal(first, last)
            and np.array_equal(first[0], mid[1])
            and np.array_equal(first[1], mid[0])
            and np.array_equal(first[0], first[1] + 1)
        )
        if not ok:
            raise ValueError("not a valid Trotterized block layout")
        if self._second_order:
            k = np.arange(num_qubits // 2)
            if not (np.array_equal(mid[0, : k.size], 2 * k) and np.array_equal(mid[1, : k.size], 2 * k + 1)):
                raise ValueError("unexpected layout of the leading half-layer")


def is_parametric_circuit(circ) -> bool:
    """
    Structural test used at every boundary of this package instead of ``isinstance``: a circuit built
    by the reference's own ``aqc_research.parametric_circuit`` (parametric_circuit.py:37-70) passes as
    well as ours -- the engine reads only these attributes.
    """
    return all(hasattr(circ, a) for a in ("num_qubits", "entangler", "blocks", "num_blocks", "num_thetas"))


def is_trotter_ansatz(circ) -> bool:
    """A circuit with the Trotterized layout (parametric_circuit.py:267-423), ours or the reference's."""
    return is_parametric_circuit(circ) and hasattr(circ, "is_second_order") and hasattr(
        circ, "half_layer_num_blocks"
    )


def layer_to_block_range(
    circ: ParametricCircuit, layer_range: Union[Tuple[int, int], None]
) -> Tuple[int, int]:
    """[from, to) range of layers -> [from, to) range of unit-blocks (None = everything)."""
    assert isinstance(circ, ParametricCircuit)
    if layer_range is None:
        return 0, circ.num_blocks
    assert chk.is_tuple(layer_range, len(layer_range) == 2)
    lo, hi = layer_range
    assert 0 <= lo < hi <= circ.num_layers
    return lo * circ.bpl, hi * circ.bpl


def first_layer_included(circ: ParametricCircuit, layer_range: Union[Tuple[int, int], None]) -> bool:
    """True if the layer range starts at layer 0 (or is None)."""
    assert isinstance(circ, ParametricCircuit)
    if layer_range is None:
        return True
    assert chk.is_tuple(layer_range, len(layer_range) == 2)
    assert 0 <= layer_range[0] < layer_range[1] <= circ.num_layers
    return layer_range[0] == 0
