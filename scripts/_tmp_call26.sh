#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/c26_summary.txt; : > $S
timeout -k 5 90 python scripts/gpu_smoke.py 6 10 13 > gpurun_out/c26_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $S
if ! grep -q "smoke ok" gpurun_out/c26_smoke.log; then tail -8 gpurun_out/c26_smoke.log; exit 1; fi
run() { name=$1; wl=$2; st=$3; shift 3
  env "$@" timeout -k 10 300 python bench.py --workload $wl --steps $st --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/c26_bench_${name}.json 2> gpurun_out/c26_bench_${name}.err
  python - <<PY | tee -a $S
import json
try:
    d = json.load(open("gpurun_out/c26_bench_${name}.json"))
    print("${name}", "value", round(d["value"], 2), d["kernel_ms"], d["details"]["tile_passes"])
except Exception as ex:
    print("${name}: no line", ex)
PY
}
run sv22_base sv22 100 A=1
run sv22_g11 sv22 100 AQC_TILE_BITS_GRAD=11
run sv22_g11_low3 sv22 100 AQC_TILE_BITS_GRAD=11 AQC_TILE_LOW_BITS=3
run sv22_a12 sv22 100 AQC_TILE_BITS_APPLY=12 AQC_TILE_LOW_BITS_APPLY=2
run sv22_a12_low3 sv22 100 AQC_TILE_BITS_APPLY=12 AQC_TILE_LOW_BITS_APPLY=3
run sv24_g10 sv24 30 AQC_TILE_BITS_GRAD=10
run sv24_a11 sv24 30 AQC_TILE_BITS_APPLY=11
