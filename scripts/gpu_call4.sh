#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/c4_summary.txt; : > $S
timeout -k 5 90 python scripts/gpu_smoke.py 6 10 13 > gpurun_out/c4_smoke.log 2>&1; echo "smoke default rc=$?" | tee -a $S; tail -4 gpurun_out/c4_smoke.log | tee -a $S
if ! grep -q "smoke ok" gpurun_out/c4_smoke.log; then echo "default smoke failed: stop" | tee -a $S; exit 1; fi
timeout -k 10 400 python -m pytest tests/test_sv_configs_gpu.py -x -q -k "stream" > gpurun_out/c4_cfg_tests.log 2>&1; echo "cfg tests rc=$?" | tee -a $S; tail -3 gpurun_out/c4_cfg_tests.log | tee -a $S
run() { # name, workload, env...
  name=$1; wl=$2; shift 2
  env "$@" timeout -k 10 240 python bench.py --workload $wl --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/c4_bench_${name}.json 2> gpurun_out/c4_bench_${name}.err
  python - <<PY | tee -a $S
import json
try:
    d = json.load(open("gpurun_out/c4_bench_${name}.json"))
    print("${name}", "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), d["kernel_ms"], "launches", d["gpu_launches"], d["details"]["tile_passes"], "frac", round(d["roofline"]["frac"], 3))
except Exception as ex:
    print("${name}: no line", ex)
PY
}
run sv28_stream_nbuf1 sv28 AQC_STREAM=1
run sv28_perpass sv28 AQC_STREAM=0
run sv24_stream_nbuf1 sv24 AQC_STREAM=1
run sv24_perpass sv24 AQC_STREAM=0
run sv20_stream_nbuf1 sv20 AQC_STREAM=1 AQC_STREAM_NBUF=1 AQC_TILE_BITS_GRAD=11 AQC_TILE_BITS_APPLY=12
run sv20_stream sv20 AQC_STREAM=1
