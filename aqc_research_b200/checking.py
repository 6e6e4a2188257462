"""
Argument predicates used by the host-side mirror of the reference API
(reference: aqc_research/checking.py -- same names where callers rely on them).
All predicates return bool so they can be used inside ``assert``.
"""

import numbers
import numpy as np

_COMPLEX = (np.complex128,)
_FLOAT = (np.float64,)


def is_int(x, cond: bool = True) -> bool:
    return isinstance(x, (int, np.integer)) and not isinstance(x, bool) and bool(cond)


def is_float(x, cond: bool = True) -> bool:
    return isinstance(x, (float, np.floating)) and bool(cond)


def is_bool(x) -> bool:
    return isinstance(x, (bool, np.bool_))


def is_str(x, cond: bool = True) -> bool:
    return isinstance(x, str) and bool(cond)


def is_tuple(x, cond: bool = True) -> bool:
    return isinstance(x, tuple) and bool(cond)


def is_list(x, cond: bool = True) -> bool:
    return isinstance(x, list) and bool(cond)


def is_dict(x, cond: bool = True) -> bool:
    return isinstance(x, dict) and bool(cond)


def is_number(x) -> bool:
    return isinstance(x, numbers.Number)


def _arr(x, kinds, ndim, cond) -> bool:
    return (
        isinstance(x, np.ndarray)
        and x.dtype.type in kinds
        and (ndim is None or x.ndim == ndim)
        and bool(cond)
    )


def float_1d(x, cond: bool = True) -> bool:
    return _arr(x, _FLOAT, 1, cond)


def float_2d(x, cond: bool = True) -> bool:
    return _arr(x, _FLOAT, 2, cond)


def complex_1d(x, cond: bool = True) -> bool:
    return _arr(x, _COMPLEX, 1, cond)


def complex_2d(x, cond: bool = True) -> bool:
    return _arr(x, _COMPLEX, 2, cond)


def complex_array(x, cond: bool = True) -> bool:
    return _arr(x, _COMPLEX, None, cond)


def complex_2d_square(x, cond: bool = True) -> bool:
    return complex_2d(x, cond) and x.shape[0] == x.shape[1]


def complex_or_float_2d(x, cond: bool = True) -> bool:
    return _arr(x, _COMPLEX + _FLOAT, 2, cond)


def block_structure(num_qubits: int, blocks) -> bool:
    """True if ``blocks`` is a valid (2, depth) integer unit-block layout."""
    return (
        is_int(num_qubits, num_qubits >= 2)
        and isinstance(blocks, np.ndarray)
        and np.issubdtype(blocks.dtype, np.integer)
        and blocks.ndim == 2
        and blocks.shape[0] == 2
        and bool(np.all((blocks >= 0) & (blocks < num_qubits)))
        and bool(np.all(blocks[0] != blocks[1]))
    )


def no_overlap(a: np.ndarray, b: np.ndarray) -> bool:
    return not np.may_share_memory(a, b)


def contiguous_c128(*arrays) -> bool:
    """All arrays are C-contiguous complex128."""
    return all(
        isinstance(a, np.ndarray) and a.dtype == np.complex128 and a.flags.c_contiguous
        for a in arrays
    )


# -- remaining predicates of the reference's checking.py (same names and meaning) ------------------
_INT = (np.int32, np.int64)


def is_complex(x, cond: bool = True) -> bool:
    return isinstance(x, (complex, np.complexfloating)) and bool(cond)


def complex_or_float_1d(x, cond: bool = True) -> bool:
    return _arr(x, _COMPLEX + _FLOAT + (np.float32, np.complex64), 1, cond)


def complex_3d(x, cond: bool = True) -> bool:
    return _arr(x, _COMPLEX + (np.complex64,), 3, cond)


def int_1d(x, cond: bool = True) -> bool:
    return _arr(x, _INT, 1, cond)


def int_2d(x, cond: bool = True) -> bool:
    return _arr(x, _INT, 2, cond)


def bool_1d(x, cond: bool = True) -> bool:
    return _arr(x, (np.bool_,), 1, cond)


def check_sim_complex_vecs4(a, b, c, d) -> bool:
    """Four complex128 vectors of one shape, each contiguous (checking.py:176-195)."""
    vecs = (a, b, c, d)
    return all(complex_1d(v) and v.flags.c_contiguous for v in vecs) and all(v.shape == a.shape for v in vecs)


def check_permutation(x) -> bool:
    """A 1D integer array holding every number of 0..size-1 once (checking.py:213-222)."""
    return int_1d(x) and bool(np.array_equal(np.sort(x), np.arange(x.size)))


def none_or_type(entity, entity_type) -> bool:
    return entity is None or isinstance(entity, entity_type)
