#!/bin/bash
# Multi-GPU validation (gpurun --gpus N -- 'bash scripts/gpu_validate_multi.sh N [Q]'): the sharded parity
# tests that fit N GPUs, the driver's `bench.py --gpus N` line (replica headline + one-state extras with the
# parity self-test) and `--workload svshard` at 2^Q amplitudes per GPU with the fused push and with the
# stand-alone pull kernel.
N=${1:-2}; Q=${2:-29}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/valm_summary.txt; : > $S
timeout -k 10 500 python -m pytest tests/test_sharded_gpu.py -q > gpurun_out/valm_shard_tests.log 2>&1; echo "sharded tests rc=$?" | tee -a $S; tail -3 gpurun_out/valm_shard_tests.log | tee -a $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -k 10 500 $TR --master-port 29611 bench.py --gpus $N > gpurun_out/valm_bench_${N}gpu.json 2> gpurun_out/valm_bench_${N}gpu.err; echo "bench --gpus $N rc=$?" | tee -a $S
for push in 1 0; do
  AQC_SHARD_PUSH=$push timeout -k 10 500 $TR --master-port 2962$push bench.py --gpus $N --workload svshard --shard-qubits $Q --steps 2 --warmup 1 > gpurun_out/valm_svshard${Q}_push$push.json 2> gpurun_out/valm_svshard${Q}_push$push.err; echo "svshard 2^$Q push=$push rc=$?" | tee -a $S
done
python - <<PY | tee -a $S
import json
def last(p):
    return json.loads([l for l in open(p) if l.startswith("{")][-1])
try:
    d = last("gpurun_out/valm_bench_${N}gpu.json")
    print("headline", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
    print("svshard", json.dumps(d["extra_workloads"]["svshard"]))
    print("parity", json.dumps(d["extra_workloads"]["svshard_parity"]))
    for push in (1, 0):
        x = last("gpurun_out/valm_svshard${Q}_push%d.json" % push)
        print("push=%d n" % push, x["config"]["num_qubits"], "evals/s", round(x["value"], 4), x["kernel_ms"])
except Exception as ex:
    print("no line", repr(ex))
PY
