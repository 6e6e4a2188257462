"""
Multi-rank (world_size 2 and 4, gloo, CPU) tests of the global-qubit sharding driver: the real
epoch planner and layouts of libaqc_b200.so, the exchange protocol and the partial-sum
reductions, with a NumPy replay standing in for the CUDA kernels (tests/sharded_sim.py).
"""

import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _launch(world, extra, port):
    cmd = [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
        "--master-addr", "127.0.0.1", "--master-port", str(port),
        os.path.join(HERE, "sharded_worker.py"),
    ] + extra
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, check=False)


# (2, 11): 10 local bits against 6-bit tiles -- the per-epoch programs go through the tile planner
@pytest.mark.parametrize("world,qubits", [(2, 7), (4, 8), (2, 11)])
def test_sharded_driver_gloo(world, qubits):
    res = _launch(world, ["--sim", "--qubits", str(qubits)], 29500 + world + qubits)
    assert res.returncode == 0 and "SHARDED_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.parametrize("world,qubits", [(2, 7), (4, 8)])
def test_sharded_driver_fused_push_protocol_gloo(world, qubits):
    """
    The slot pool and the landing slots of the FUSED layout switch (sharded.py ``_run`` with push) with
    the NumPy backend emulating the delivery: five slots must suffice for the gradient sweep, pushed
    slots must never alias the slots an epoch works on, results equal the oracle's.
    """
    res = _launch(world, ["--sim", "--sim-push", "--qubits", str(qubits)], 29520 + world)
    assert res.returncode == 0 and "SHARDED_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
    assert "p2p=True" in res.stdout


def test_locate_is_a_bijection():
    import numpy as np
    from aqc_research_b200.sharded import locate
    from sharded_sim import shard_of

    n, g = 7, 2
    vec = np.arange(2**n)
    seen = set()
    for idx in range(2**n):
        r, off = locate(idx, n, g)
        assert shard_of(vec, n, g, r)[off] == idx
        seen.add((r, off))
    assert len(seen) == 2**n
