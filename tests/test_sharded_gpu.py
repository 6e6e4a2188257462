"""
GPU test of the sharded state vector: one process per GPU (torchrun, NCCL), CUDA backend,
layout switches by the peer-memory exchange kernel (CUDA IPC) and by isend/irecv; every rank
checks objective / gradient against the single-process oracle.  Needs >= 2 GPUs.
"""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    from aqc_research_b200 import _lib

    return _lib.device_count()


def _launch(world, extra, port):
    cmd = [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
        "--master-addr", "127.0.0.1", "--master-port", str(port),
        os.path.join(HERE, "sharded_worker.py"), "--gpu",
    ] + extra
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, check=False)


@pytest.mark.parametrize("extra", [[], ["--no-p2p"]])
def test_sharded_two_gpus(extra):
    if _ngpu() < 2:
        pytest.skip("needs at least 2 GPUs")
    res = _launch(2, ["--qubits", "14"] + extra, 29711 + len(extra))
    assert res.returncode == 0 and "SHARDED_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_sharded_four_gpus():
    if _ngpu() < 4:
        pytest.skip("needs at least 4 GPUs")
    res = _launch(4, ["--qubits", "16"], 29731)
    assert res.returncode == 0 and "SHARDED_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
