#!/bin/bash
# 2-GPU check of the fused NVLink push: sharded parity tests, bench --gpus 2 (svshard + parity self-test), A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/c5_summary.txt; : > $S
nvidia-smi -L | tee -a $S
timeout -k 5 90 python scripts/gpu_smoke.py 6 10 13 > gpurun_out/c5_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $S; tail -2 gpurun_out/c5_smoke.log | tee -a $S
if ! grep -q "smoke ok" gpurun_out/c5_smoke.log; then echo "smoke failed: stop" | tee -a $S; exit 1; fi
timeout -k 10 400 python -m pytest tests/test_sharded_gpu.py -x -q -k "two_gpus" > gpurun_out/c5_shard_tests.log 2>&1; echo "sharded tests (push) rc=$?" | tee -a $S; tail -15 gpurun_out/c5_shard_tests.log | tee -a $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -k 10 400 $TR --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --shard-qubits 28 > gpurun_out/c5_bench_2gpu.json 2> gpurun_out/c5_bench_2gpu.err; echo "bench --gpus 2 rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import json
try:
    d = json.loads([l for l in open("gpurun_out/c5_bench_2gpu.json") if l.startswith("{")][-1])
    print("headline", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
    x = d["extra_workloads"]
    print("svshard", json.dumps(x["svshard"]))
    print("parity", json.dumps(x["svshard_parity"]))
except Exception as ex:
    print("no line", ex)
PY
for push in 1 0; do
  AQC_SHARD_PUSH=$push timeout -k 10 300 $TR --master-port 2962$push bench.py --gpus 2 --workload svshard --shard-qubits 29 --steps 3 --warmup 1 > gpurun_out/c5_svshard29_push$push.json 2> gpurun_out/c5_svshard29_push$push.err; echo "svshard 2^29 push=$push rc=$?" | tee -a $S
  python - <<PY | tee -a $S
import json
try:
    d = json.loads([l for l in open("gpurun_out/c5_svshard29_push$push.json") if l.startswith("{")][-1])
    print("push=$push n", d["config"]["num_qubits"], "evals/s", round(d["value"], 4), d["kernel_ms"], {k: d["sharded"].get(k) for k in ("fused_push", "vector_switches_per_eval", "nvlink_gbs_per_gpu_over_whole_step")})
except Exception as ex:
    print("no line", ex)
PY
done
for f in gpurun_out/c5_*.err; do echo "== $f"; tail -4 $f; done | tail -40
