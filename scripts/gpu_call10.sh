#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/c10_summary.txt; : > $S
timeout -k 5 90 python scripts/gpu_smoke.py 6 10 13 > gpurun_out/c10_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $S
if ! grep -q "smoke ok" gpurun_out/c10_smoke.log; then tail -5 gpurun_out/c10_smoke.log; echo "smoke failed: stop" | tee -a $S; exit 1; fi
timeout -k 10 600 python -m pytest tests/test_sv_configs_gpu.py tests/test_sv_gpu.py tests/test_objectives_gpu.py -q -k "not n24" > gpurun_out/c10_tests.log 2>&1; echo "sv tests rc=$?" | tee -a $S; tail -3 gpurun_out/c10_tests.log | tee -a $S
run() { name=$1; wl=$2; shift 2
  env "$@" timeout -k 10 120 python bench.py --workload $wl --steps 200 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/c10_bench_${name}.json 2> gpurun_out/c10_bench_${name}.err
  python - <<PY | tee -a $S
import json
try:
    d = json.load(open("gpurun_out/c10_bench_${name}.json"))
    print("${name}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["kernel_ms"], "launches/step", d["gpu_launches"] / d["steps"], d["details"]["tile_passes"])
except Exception as ex:
    print("${name}: no line", ex)
PY
}
run sv20_ctas4 sv20 AQC_GRAD_CTAS=4
run sv20_ctas3 sv20 AQC_GRAD_CTAS=3
run sv16_ctas4 sv16 AQC_GRAD_CTAS=4
run sv16_ctas3 sv16 AQC_GRAD_CTAS=3
run sv12_ctas4 sv12 AQC_GRAD_CTAS=4
run sv12_ctas3 sv12 AQC_GRAD_CTAS=3
run sv22_ctas4 sv22 AQC_GRAD_CTAS=4
run sv22_ctas3 sv22 AQC_GRAD_CTAS=3
timeout -k 10 120 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/c10_plain_sv20.json 2> gpurun_out/c10_plain_sv20.err &&
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c10_launches_sv20.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/c10_ncu_launches.log 2>&1
echo "ncu launch list rc=$?" | tee -a $S
