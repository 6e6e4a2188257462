"""
CPU checks of the C-ABI boundary: the shared library loads, exports every symbol that
include/aqc_b200.h declares, validates arguments, and refuses to compute without a GPU
(there is no CPU fallback).
"""

import ctypes as ct
import os
import re

import numpy as np
import pytest

from aqc_research_b200 import _lib
from aqc_research_b200 import circuit_structures as cs
from aqc_research_b200.engine import CircuitHandle, SvWorkspace
from aqc_research_b200.parametric_circuit import ParametricCircuit, TrotterAnsatz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "aqc_b200.h"), encoding="utf-8").read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aqc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in aqc_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert lib.aqc_version() >= 100


def test_circuit_validation_errors():
    blocks = np.array([[0, 1], [1, 1]])  # control == target in block 1
    with pytest.raises(ValueError):
        ParametricCircuit(3, "cx", blocks)
    lib = _lib.load()
    bad = np.ascontiguousarray(blocks, dtype=np.int32)
    h = ct.c_void_p()
    rc = lib.aqc_circuit_create(3, 0, bad.ctypes.data_as(_lib.c_int32_p), 2, 0, ct.byref(h))
    assert rc == -1 and b"valid" in lib.aqc_last_error()
    # Trotter layout is validated by the library too
    ok = np.ascontiguousarray(cs.make_trotter_like_circuit(4, 1), dtype=np.int32)
    assert lib.aqc_circuit_create(4, 0, ok.ctypes.data_as(_lib.c_int32_p), ok.shape[1], 2, ct.byref(h)) == 0
    assert lib.aqc_circuit_num_thetas(h) == 3 * 4 + 4 * ok.shape[1]
    lib.aqc_circuit_destroy(h)
    swapped = ok[::-1].copy()
    assert lib.aqc_circuit_create(4, 0, swapped.ctypes.data_as(_lib.c_int32_p), ok.shape[1], 2, ct.byref(h)) == -1
    with pytest.raises(ValueError):
        TrotterAnsatz(4, ok[::-1].copy(), True)


def test_no_cpu_fallback():
    """Without a CUDA device the workspace cannot be created: compute calls fail loudly."""
    if _lib.device_count() > 0:
        pytest.skip("a GPU is visible")
    circ = TrotterAnsatz(4, cs.make_trotter_like_circuit(4, 1), True)
    CircuitHandle(circ)  # host-only structure is fine
    with pytest.raises(_lib.AqcError, match="no CUDA device"):
        SvWorkspace(circ, num_slots=3)
    from aqc_research_b200 import core_operations as cop

    v = np.zeros(16, dtype=np.complex128)
    with pytest.raises(_lib.AqcError):
        cop.v_mul_vec(circ, np.zeros(circ.num_thetas), v, v.copy())
