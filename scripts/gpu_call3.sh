#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/c3_summary.txt; : > $S
timeout -k 5 90 python scripts/gpu_smoke.py 6 10 13 > gpurun_out/c3_smoke.log 2>&1; echo "smoke default rc=$?" | tee -a $S; tail -4 gpurun_out/c3_smoke.log | tee -a $S
if ! grep -q "smoke ok" gpurun_out/c3_smoke.log; then echo "default smoke failed: stop" | tee -a $S; exit 1; fi
timeout -k 10 500 python -m pytest tests/test_sv_configs_gpu.py -x -q -k "not n24" > gpurun_out/c3_cfg_tests.log 2>&1; echo "cfg tests rc=$?" | tee -a $S; tail -3 gpurun_out/c3_cfg_tests.log | tee -a $S
timeout -k 10 600 python -m pytest tests/test_sv_gpu.py tests/test_objectives_gpu.py tests/test_sketching_gpu.py tests/test_coord_descent_gpu.py tests/test_mps_gpu.py -x -q > gpurun_out/c3_sv_tests.log 2>&1; echo "sv+mps tests rc=$?" | tee -a $S; tail -5 gpurun_out/c3_sv_tests.log | tee -a $S
run() { # name, workload, env...
  name=$1; wl=$2; shift 2
  env "$@" timeout -k 10 240 python bench.py --workload $wl --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/c3_bench_${name}.json 2> gpurun_out/c3_bench_${name}.err
  python - <<PY | tee -a $S
import json
try:
    d = json.load(open("gpurun_out/c3_bench_${name}.json"))
    print("${name}", "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), d["kernel_ms"], "launches", d["gpu_launches"], d["details"]["tile_passes"])
except Exception as ex:
    print("${name}: no line", ex)
PY
}
run sv20_stream sv20 AQC_STREAM=1
run sv20_perpass sv20 AQC_STREAM=0
run sv20_stream_rebal sv20 AQC_STREAM=1 AQC_STREAM_REBAL=1
run sv20_stream_nocoop sv20 AQC_STREAM=1 AQC_STREAM_COOP=0
run sv12_stream sv12 AQC_STREAM=1
run sv12_perpass sv12 AQC_STREAM=0
run sv28_stream sv28 AQC_STREAM=1
run sv28_perpass sv28 AQC_STREAM=0
run sv28_perpass_tb10 sv28 AQC_STREAM=0 AQC_TILE_BITS_GRAD=10 AQC_TILE_BITS_APPLY=11
run sv28_stream_rebal sv28 AQC_STREAM=1 AQC_STREAM_REBAL=1
# ncu: the stream kernel on sv24 (two launches per evaluation: V^H sweep, gradient sweep)
timeout -k 10 200 python bench.py --workload sv24 --steps 1 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/c3_plain_sv24.json 2> gpurun_out/c3_plain_sv24.err &&
timeout -k 10 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:dense_stream_kernel -s 6 -c 2 -o gpurun_out/c3_prof_sv24 -f python bench.py --workload sv24 --steps 1 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/c3_ncu_sv24.log 2>&1
echo "ncu rc=$?" | tee -a $S
