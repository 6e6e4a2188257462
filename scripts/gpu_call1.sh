#!/bin/bash
# first GPU check of the persistent sweep kernel: parity tests, then A/B bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout -k 10 900 python -m pytest tests/test_sv_configs_gpu.py -x -q > gpurun_out/c1_cfg_tests.log 2>&1
echo "cfg tests rc=$?" | tee -a gpurun_out/c1_summary.txt
tail -5 gpurun_out/c1_cfg_tests.log
timeout -k 10 900 python -m pytest tests/test_sv_gpu.py tests/test_objectives_gpu.py tests/test_sketching_gpu.py tests/test_coord_descent_gpu.py -x -q > gpurun_out/c1_sv_tests.log 2>&1
echo "sv tests rc=$?" | tee -a gpurun_out/c1_summary.txt
tail -5 gpurun_out/c1_sv_tests.log
for wl in sv20 sv12 sv28; do
  for st in 1 0; do
    AQC_STREAM=$st timeout -k 10 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/c1_bench_${wl}_stream${st}.json 2> gpurun_out/c1_bench_${wl}_stream${st}.err
    echo "bench $wl stream=$st rc=$?" | tee -a gpurun_out/c1_summary.txt
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/c1_bench_${wl}_stream${st}.json"))
    print("${wl} stream=${st}", "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), d["kernel_ms"], "launches", d["gpu_launches"], d["config"]["tile_passes"])
except Exception as ex:
    print("no line", ex)
PY
  done
done
