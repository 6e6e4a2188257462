"""
On-disk interoperability of the cached target files: lists of
TargetMpsState / TargetClassicState pickled under the reference's module path
(aqc_research/model_sp_lhs/trotter/target_states.py:44-132, 285-371, 274-275, 509-510).
"""

import os
import pickle
import types

import numpy as np
import pytest

from aqc_research_b200.model_sp_lhs.trotter import target_states as ts
from aqc_research_b200.model_sp_lhs.trotter import trotter as trotop

REF = "/root/reference"


def _opts(tmp):
    o = types.SimpleNamespace()
    o.trotter_steps = np.array([2, 4, 6])
    o.evol_times = np.round(np.array([0.4, 0.8, 1.2]), 3)
    o.trunc_thr_target = 1e-16
    o.delta = 1.0
    o.ini_state_func = (trotop.neel_init_state,)
    o.result_dir = str(tmp)
    return o


def _fake_targets(opts, mps: bool):
    out = []
    for i, (s, t) in enumerate(zip(opts.trotter_steps, opts.evol_times)):
        if mps:
            a, b = ts.product_mps(4, [0, 2]), ts.product_mps(4, [0, 2])
            out.append(ts.TargetMpsState(opts=opts, num_qubits=4, num_trot_steps=s, evol_time=t, my_id=i,
                                         t1_gt=a, t1=b, second_order=True))
        else:
            v = np.zeros(16, dtype=np.complex128)
            v[5] = 1
            out.append(ts.TargetClassicState(opts=opts, num_qubits=4, num_trot_steps=s, evol_time=t, my_id=i,
                                             t1_gt=v, t1=v.copy(), second_order=False))
    return out


@pytest.mark.parametrize("mps", [False, True])
def test_round_trip_and_reference_module_path(tmp_path, mps):
    opts = _opts(tmp_path)
    data = _fake_targets(opts, mps)
    path = os.path.join(tmp_path, "targets.pkl")
    ts.save_targets(data, path)
    raw = open(path, "rb").read()
    assert ts.REFERENCE_MODULE.encode() in raw and b"aqc_research_b200" not in raw
    back = ts.load_targets(path)
    cls = ts.TargetMpsState if mps else ts.TargetClassicState
    assert all(type(d) is cls for d in back) and cls.check_cached_data(opts, 4, back)
    assert cls.__module__ == ts.__name__  # the alias is undone after the dump
    assert back[1].num_trot_steps == 4 and back[2].evol_time == 1.2
    # a changed option invalidates the cache, as in the reference
    opts.delta = 2.0
    assert not cls.check_cached_data(opts, 4, back)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not present (build container only)")
def test_files_open_in_the_reference_and_back(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from ref_loader import load_reference

    load_reference()
    import aqc_research.model_sp_lhs.trotter.target_states as rts  # the unmodified reference module

    opts = _opts(tmp_path)
    ours = os.path.join(tmp_path, "ours.pkl")
    ts.save_targets(_fake_targets(opts, False), ours)
    with open(ours, "rb") as fld:
        got = pickle.load(fld)  # what the reference does (target_states.py:481-482)
    assert all(type(d) is rts.TargetClassicState for d in got)
    assert rts.TargetClassicState.check_cached_data(opts, 4, got)
    theirs = os.path.join(tmp_path, "theirs.pkl")
    with open(theirs, "wb") as fld:
        pickle.dump(got, fld)  # what the reference writes (:509-510)
    back = ts.load_targets(theirs)
    assert all(type(d) is ts.TargetClassicState for d in back)
    assert ts.TargetClassicState.check_cached_data(opts, 4, back)
