#!/bin/bash
# 8-GPU validation: sharded tests on 4 GPUs, the driver's `bench.py --gpus 8` line, and BASELINE config 5
# (n = 34 over 8 GPUs, 2^31 amplitudes per GPU) with the fused push and with the separate pull kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/c6_summary.txt; : > $S
nvidia-smi -L | wc -l | tee -a $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -k 10 400 $TR --master-port 29811 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/c6_bench_8gpu.json 2> gpurun_out/c6_bench_8gpu.err; echo "bench --gpus 8 rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import json
try:
    d = json.loads([l for l in open("gpurun_out/c6_bench_8gpu.json") if l.startswith("{")][-1])
    print("headline", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
    x = d["extra_workloads"]
    print("svshard", json.dumps(x["svshard"]))
    print("parity", json.dumps(x["svshard_parity"]))
except Exception as ex:
    print("no line", ex)
PY
for push in 1 0; do
  AQC_SHARD_PUSH=$push timeout -k 10 500 $TR --master-port 2982$push bench.py --gpus 8 --workload svshard --shard-qubits 31 --steps 2 --warmup 1 > gpurun_out/c6_svshard31_push$push.json 2> gpurun_out/c6_svshard31_push$push.err; echo "svshard 2^31 push=$push rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads([l for l in open("gpurun_out/c6_svshard31_push$push.json") if l.startswith("{")][-1])
    print("push=$push n", d["config"]["num_qubits"], "evals/s", round(d["value"], 4), d["kernel_ms"], {k: d["sharded"].get(k) for k in ("fused_push", "epochs", "vector_switches_per_eval", "nvlink_bytes_sent_per_gpu_per_eval", "nvlink_gbs_per_gpu_over_whole_step")})
except Exception as ex:
    print("no line", ex)
PY
done
for f in gpurun_out/c6_*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $f | tail -6; done | tail -40
