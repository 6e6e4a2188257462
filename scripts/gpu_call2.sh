#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/c2_summary.txt; : > $S
timeout -k 5 90 python scripts/gpu_smoke.py 6 10 13 > gpurun_out/c2_smoke.log 2>&1; echo "smoke default rc=$?" | tee -a $S; tail -4 gpurun_out/c2_smoke.log | tee -a $S
AQC_STREAM_REBAL=1 timeout -k 5 60 python scripts/gpu_smoke.py 10 > gpurun_out/c2_smoke_rebal.log 2>&1; echo "smoke rebal rc=$?" | tee -a $S; tail -2 gpurun_out/c2_smoke_rebal.log | tee -a $S
if ! grep -q "smoke ok" gpurun_out/c2_smoke.log; then echo "default smoke failed: stop" | tee -a $S; exit 1; fi
timeout -k 10 600 python -m pytest tests/test_sv_configs_gpu.py -x -q > gpurun_out/c2_cfg_tests.log 2>&1; echo "cfg tests rc=$?" | tee -a $S; tail -3 gpurun_out/c2_cfg_tests.log | tee -a $S
timeout -k 10 600 python -m pytest tests/test_sv_gpu.py tests/test_objectives_gpu.py tests/test_sketching_gpu.py tests/test_coord_descent_gpu.py -x -q > gpurun_out/c2_sv_tests.log 2>&1; echo "sv tests rc=$?" | tee -a $S; tail -3 gpurun_out/c2_sv_tests.log | tee -a $S
for wl in sv20 sv12 sv28; do
  for st in 1 0; do
    AQC_STREAM=$st timeout -k 10 240 python bench.py --workload $wl --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/c2_bench_${wl}_stream${st}.json 2> gpurun_out/c2_bench_${wl}_stream${st}.err
    echo "bench $wl stream=$st rc=$?" | tee -a $S
    python - <<PY | tee -a $S
import json
try:
    d = json.load(open("gpurun_out/c2_bench_${wl}_stream${st}.json"))
    print("${wl} stream=${st}", "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), d["kernel_ms"], "launches", d["gpu_launches"], d["config"]["tile_passes"])
except Exception as ex:
    print("no line", ex)
PY
  done
done
