"""
GPU generation of the time-evolution targets: incremental Trotter
evolution of an MPS (generate_all_mps_targets, target_states.py:135-231) against the dense
state-vector evolution, and the cache files.
"""

import types

import numpy as np
import pytest

from aqc_research_b200.model_sp_lhs.trotter import target_states as ts
from aqc_research_b200.model_sp_lhs.trotter import trotter as trotop
from aqc_research_b200.mps_operations import mps_to_vector

pytestmark = pytest.mark.gpu
# "no truncation" in the reference means trunc_thr = 1e-16 on the SUM OF SQUARED Schmidt values
# (mps_operations.py:26-30): components up to 1e-8 in norm may be dropped at every split, so MPS and
# dense evolution agree to ~sqrt(thr) x (number of splits that drop something), not to 1e-10.
TOL = 5e-7


def _opts(tmp, n):
    o = types.SimpleNamespace()
    o.trotter_steps = np.array([2, 4, 6])
    o.evol_times = np.round(np.array([0.4, 0.8, 1.2]), 3)
    o.trunc_thr_target = 1e-16
    o.delta = 1.0
    o.ini_state_func = (trotop.neel_init_state,)
    o.result_dir = str(tmp)
    o.num_qubits, o.use_mps, o.second_order_trotter, o.targets_file = n, True, True, ""
    return o


@pytest.mark.parametrize("second_order", [False, True])
def test_incremental_mps_targets_match_dense_evolution(tmp_path, second_order):
    n = 6
    opts = _opts(tmp_path, n)
    targets = ts.generate_all_mps_targets(opts=opts, num_qubits=n, second_order=second_order)
    assert ts.TargetMpsState.check_cached_data(opts, n, targets)
    ini = trotop.basis_index(trotop.neel_init_state(n))
    vec_gt = np.zeros(2**n, dtype=np.complex128)
    vec_gt[ini] = 1
    vec = vec_gt.copy()
    for i, tg in enumerate(targets):
        dt, steps = 0.4, 2  # equal intervals, uniform stepping
        kw = dict(evol_time=dt, delta=opts.delta, second_order=second_order)
        vec_gt = trotop.trotter_state(n, num_steps=steps * ts.precise_multiplier(), ini_state=vec_gt, **kw)
        vec = trotop.trotter_state(n, num_steps=steps, ini_state=vec, **kw)
        assert np.linalg.norm(mps_to_vector(tg.t1_gt) - vec_gt) < TOL, i
        assert np.linalg.norm(mps_to_vector(tg.t1) - vec) < TOL, i
        assert 1 - abs(np.vdot(mps_to_vector(tg.t1), vec)) ** 2 < 1e-12, i
        assert trotop.fidelity(tg.t1_gt, tg.t1) > 0.99


def test_cache_files(tmp_path):
    n = 5
    opts = _opts(tmp_path, n)
    first = ts.get_target_states(opts)  # computes and stores target_mps_states_n5.pkl
    again = ts.get_target_states(opts)  # loads
    assert len(first) == len(again) == 3
    for a, b in zip(first, again):
        assert np.linalg.norm(mps_to_vector(a.t1) - mps_to_vector(b.t1)) == 0.0
    opts.use_mps = False
    dense = ts.get_target_states(opts)
    for a, d in zip(first, dense):
        # one circuit over [0, t] vs the concatenation of incremental circuits: the same product
        # formula up to a GLOBAL PHASE (each ansatz triplet carries a constant phase and the
        # concatenation has more half-layer triplets); every consumer uses |<.|.>|^2
        assert 1 - abs(np.vdot(mps_to_vector(a.t1), d.t1)) ** 2 < 1e-12
        assert 1 - abs(np.vdot(mps_to_vector(a.t1_gt), d.t1_gt)) ** 2 < 1e-12
