"""
Worker of the sharded-state tests; run under torchrun with world_size 2 or 4:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 ... sharded_worker.py --sim
--sim : CPU, gloo, NumPy replay backend      (tests/test_sharded_cpu.py)
--gpu : one GPU per rank, NCCL, CUDA backend (tests/test_sharded_gpu.py; --no-p2p forces send/recv)
Every rank checks the sharded objective / gradient against the single-process oracle.
"""

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sim", action="store_true")
    ap.add_argument("--sim-push", action="store_true", help="--sim with the fused-push protocol emulated")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--no-p2p", action="store_true")
    ap.add_argument("--qubits", type=int, default=8)
    ap.add_argument("--layers", type=int, default=2)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpu:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    else:
        dist.init_process_group("gloo")

    from aqc_research_b200 import circuit_structures as cs
    from aqc_research_b200.parametric_circuit import ParametricCircuit, TrotterAnsatz
    from aqc_research_b200.sharded import DistComm, GpuShardBackend, ShardedStateVector
    from oracle import sv_oracle as O
    from sharded_sim import SimShardBackend, shard_of

    n, g = args.qubits, world.bit_length() - 1
    comm = DistComm()
    rng = np.random.RandomState(99)  # same stream on every rank
    circuits = [
        ("trotter2", TrotterAnsatz(n, cs.make_trotter_like_circuit(n, args.layers), True)),
        ("trotter1-deep", TrotterAnsatz(n, cs.make_trotter_like_circuit(n, n), False)),
        ("spin-cz", ParametricCircuit(n, "cz", cs.create_ansatz_structure(n, "spin", "full", 3 * (n - 1)))),
        ("line-cp", ParametricCircuit(n, "cp", cs.create_ansatz_structure(n, "line", "full", 2 * n))),
    ]
    worst = 0.0
    for name, circ in circuits:
        th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
        y = rng.randn(2**n) + 1j * rng.randn(2**n)
        y /= np.linalg.norm(y)
        if args.gpu:
            be = GpuShardBackend(circ, g, rank, local_rank, num_slots=5)
        else:
            be = SimShardBackend(circ, g, rank, num_slots=5)
            if args.sim_push:
                be.transpose = comm.transpose_chunks
        sv = ShardedStateVector(circ, comm, be, use_p2p=not args.no_p2p)
        assert sv.push == (args.sim_push or (args.gpu and not args.no_p2p and os.environ.get("AQC_SHARD_PUSH", "1") != "0"))
        be.upload(sv.slot["target"], shard_of(y, n, g, rank))
        idx = O.basis_state_indices(n, init_index=(1 << (n - 1)) | 1)
        hs = sv.objective(th, idx)
        z0 = O.apply_v(circ, th, y, dagger=True)
        err_hs = np.linalg.norm(hs - z0[idx]) / np.linalg.norm(z0[idx])
        err_z0 = np.linalg.norm(be.download(sv.slot["z0"]) - shard_of(z0, n, g, rank)) / np.linalg.norm(z0) * np.sqrt(world)
        nrm = abs(sv.vdot("z0", "z0") - 1.0)  # (the gradient sweep consumes z0)
        xb = int(idx[2])
        grad = sv.grad(th, xb, keep_states=True)
        e = np.zeros(2**n, dtype=complex)
        e[xb] = 1
        gref = O.grad_sweep(circ, th, e, z0)
        err_g = np.linalg.norm(grad - gref) / np.linalg.norm(gref)
        # z ends as V V^H y = y, w as V e_x -- both back in layout A
        err_z = np.linalg.norm(be.download(sv.slot["z"]) - shard_of(y, n, g, rank)) * np.sqrt(world)
        err_w = np.linalg.norm(be.download(sv.slot["w"]) - shard_of(O.apply_v(circ, th, e), n, g, rank)) * np.sqrt(world)
        errs = dict(hs=err_hs, z0=err_z0, grad=err_g, norm=nrm)
        errs.update(z=err_z, w=err_w)
        worst = max(worst, max(errs.values()))
        if rank == 0:
            print(f"[{name}] world={world} p2p={sv.p2p} epochs(grad,dag)=({be.num_epochs(0)},{be.num_epochs(2)}) "
                  + " ".join(f"{k}={v:.2e}" for k, v in errs.items()), flush=True)
        sv.close()
    # synthetic target generated per rank is rank-count independent (GPU only)
    if args.gpu:
        circ = circuits[0][1]
        be = GpuShardBackend(circ, g, rank, local_rank, num_slots=5)
        sv = ShardedStateVector(circ, comm, be, use_p2p=not args.no_p2p)
        sv.set_target_random(1234)
        nrm = abs(sv.vdot("target", "target") - 1.0)
        probe = sv.amplitudes("target", [0, 1, 2**n - 1, 2 ** (n - 1) + 3])
        if rank == 0:
            print(f"[random target] norm err {nrm:.2e} probe {np.round(probe, 6)}", flush=True)
        worst = max(worst, nrm)
        sv.close()
    ok = worst < 1e-10
    if rank == 0:
        print("SHARDED_OK" if ok else f"SHARDED_FAIL worst={worst:.3e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
