"""
Host logic of the gate-by-gate drop-ins on CPU: strides, control modes, argument order and in-place
behaviour, with a NumPy restatement of what csrc/aqc_prim.cu computes standing in for the GPU call.
Same golden assertions as tests/test_primitives_gpu.py (which checks the kernels themselves).
"""

import numpy as np
import pytest

import test_primitives_gpu as gpu_tests
from aqc_research_b200 import _prim


def _transform(flat, op):
    st, sc, mode, g = int(op[0]), int(op[1]), int(op[2]), np.asarray(op[3], dtype=np.complex128)
    out = flat.copy()
    idx = np.arange(flat.size)
    i0 = idx[(idx // st) % 2 == 0]
    i1 = i0 + st
    a, b = flat[i0], flat[i1]
    with np.errstate(invalid="ignore"):
        na = (0 if g[0, 0] == 0 else g[0, 0] * a) + (0 if g[0, 1] == 0 else g[0, 1] * b)
        nb = (0 if g[1, 0] == 0 else g[1, 0] * a) + (0 if g[1, 1] == 0 else g[1, 1] * b)
    if mode != 0:
        on = (i0 // sc) % 2 == 1
        na = np.where(on, na, 0 if mode == 2 else a)
        nb = np.where(on, nb, 0 if mode == 2 else b)
    out[i0], out[i1] = na, nb
    return out


def _apply_gates(arr, ops, device=0):
    flat = arr.reshape(-1)
    for op in ops:
        flat[:] = _transform(flat, op)
    return arr


def _gate_vdot(w, z, op, device=0):
    return complex(np.vdot(_transform(w.reshape(-1), op), z.reshape(-1)))


@pytest.fixture(autouse=True)
def numpy_kernels(monkeypatch):
    monkeypatch.setattr(_prim, "apply_gates", _apply_gates)
    monkeypatch.setattr(_prim, "gate_vdot", _gate_vdot)


def test_vector_primitives_host_logic():
    gpu_tests.test_vector_primitives_match_the_reference()
    gpu_tests.test_projector_does_not_spread_non_finite_values()
    gpu_tests.test_bad_arguments_raise()
    gpu_tests.test_degenerate_gates_and_blocks_vs_kronecker()


@pytest.mark.parametrize("m", [16, 5])
def test_matrix_primitives_host_logic(m):
    gpu_tests.test_matrix_primitives_match_the_reference(m)
