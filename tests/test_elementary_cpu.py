"""
elementary_operations against golden outputs of the unmodified reference
(elementary_operations.py:39-291): gate conventions and the dense unit-block / CX matrices.
"""

import numpy as np

from golden_util import load
from aqc_research_b200 import elementary_operations as eo


def test_gate_matrices_and_block_matrices():
    g = load("primitive_cases.npz")
    n, ang = int(g["n"]), float(g["angle"])
    for nm in ("np_rx", "np_ry", "np_rz", "np_phase"):
        assert np.allclose(getattr(eo, nm)(ang), g[nm], rtol=0, atol=1e-15), nm
    out = np.empty((2, 2), dtype=np.complex128)
    for nm in ("rx", "ry", "rz"):
        assert getattr(eo, "make_" + nm)(ang, out) is out
        assert np.allclose(out, g["np_" + nm], rtol=0, atol=1e-15)
    assert np.array_equal(eo.np_x(), g["np_x"]) and np.array_equal(eo.np_z(), g["np_z"])
    assert np.allclose(eo.np_y(), 1j * eo.np_x() @ eo.np_z())
    for c in range(n):
        for t in range(n):
            if c == t:
                continue
            blk = eo.np_block_matrix(n, c, t, g["c_mat"], g["t_mat"], g["g_mat"])
            assert np.allclose(blk, g[f"np_block_{c}{t}"], rtol=0, atol=1e-13), (c, t)
            assert np.array_equal(eo.np_cx_matrix(n, c, t), g[f"np_cx_{c}{t}"])
