#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/c9_summary.txt; : > $S
timeout -k 5 90 python scripts/gpu_smoke.py 6 10 13 > gpurun_out/c9_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $S
if ! grep -q "smoke ok" gpurun_out/c9_smoke.log; then tail -5 gpurun_out/c9_smoke.log; echo "smoke failed: stop" | tee -a $S; exit 1; fi
timeout -k 10 1500 python -m pytest tests -m gpu -q -s > gpurun_out/c9_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $S; grep -E "passed|failed|error" gpurun_out/c9_gpu_tests.log | tail -3 | tee -a $S; grep -E "^\[mps" gpurun_out/c9_gpu_tests.log | tee -a $S; grep -E "^FAILED|^ERROR" gpurun_out/c9_gpu_tests.log | tee -a $S
timeout -k 10 600 python bench.py > gpurun_out/c9_bench_default.json 2> gpurun_out/c9_bench_default.err; echo "bench default rc=$?" | tee -a $S
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c9_bench_reference.json 2> gpurun_out/c9_bench_reference.err; echo "bench reference rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import json
try:
    d = json.loads([l for l in open("gpurun_out/c9_bench_default.json") if l.startswith("{")][-1])
    print("headline value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 4), "launches", d["gpu_launches"], "steps", d["steps"], "roofline frac", round(d["roofline"]["frac"], 3), d["kernel_ms"], d["clocks"])
    for k, v in d.get("extra_workloads", {}).items():
        if isinstance(v, dict):
            print(" ", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "e2e_value", "unit", "ms_per_step", "kernel_ms", "wall_s", "evaluations", "z0_norm", "ms_per_sweep_batch", "ms_per_eval")})
        else:
            print(" ", k, v)
    r = json.loads([l for l in open("gpurun_out/c9_bench_reference.json") if l.startswith("{")][-1])
    print("reference arm", round(r["value"], 3), r["cpu_baseline"]["cores"], "same config:", r["config"] == d["config"])
except Exception as ex:
    print("no line", repr(ex))
PY
run() { name=$1; wl=$2; shift 2
  env "$@" timeout -k 10 120 python bench.py --workload $wl --steps 100 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/c9_bench_${name}.json 2> gpurun_out/c9_bench_${name}.err
  python - <<PY | tee -a $S
import json
try:
    d = json.load(open("gpurun_out/c9_bench_${name}.json"))
    print("${name}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["kernel_ms"], "launches/step", d["gpu_launches"] / d["steps"], d["details"]["tile_passes"])
except Exception as ex:
    print("${name}: no line", ex)
PY
}
run sv12 sv12 A=1

run sv16 sv16 A=1



timeout -k 10 120 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/c9_plain_sv20.json 2> gpurun_out/c9_plain_sv20.err &&
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c9_launches_sv20.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/c9_ncu_launches.log 2>&1
echo "ncu launch list rc=$?" | tee -a $S
