#!/bin/bash
# One-GPU validation run used at the end of a round (under gpurun): smoke, the whole `-m gpu` suite, the
# driver's default bench line and the reference arm, the ncu launch list of the headline workload and the
# `--set full` captures of the gradient tile pass (condensed by scripts/ncu_traffic.py here afterwards).
# Everything lands in gpurun_out/val_*.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/val_summary.txt; : > $S
timeout -k 5 90 python scripts/gpu_smoke.py 6 10 13 > gpurun_out/val_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $S
if ! grep -q "smoke ok" gpurun_out/val_smoke.log; then tail -5 gpurun_out/val_smoke.log; echo "smoke failed: stop" | tee -a $S; exit 1; fi
timeout -k 10 1500 python -m pytest tests -m gpu -q -s > gpurun_out/val_gpu_tests.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $S
grep -E "passed|failed|error" gpurun_out/val_gpu_tests.log | tail -3 | tee -a $S; grep -E "^\[mps|^FAILED|^ERROR" gpurun_out/val_gpu_tests.log | tee -a $S
timeout -k 10 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_entry_smoke.log 2>&1; echo "__graft_entry__.smoke rc=$?" | tee -a $S; tail -1 gpurun_out/val_entry_smoke.log | tee -a $S
timeout -k 10 600 python bench.py > gpurun_out/val_bench_default.json 2> gpurun_out/val_bench_default.err; echo "bench default rc=$?" | tee -a $S
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/val_bench_reference.json 2> gpurun_out/val_bench_reference.err; echo "bench reference rc=$?" | tee -a $S
python - <<'PY' | tee -a $S
import json
try:
    d = json.loads([l for l in open("gpurun_out/val_bench_default.json") if l.startswith("{")][-1])
    print("headline value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 4), "launches", d["gpu_launches"], "steps", d["steps"], "roofline frac", round(d["roofline"]["frac"], 3), d["kernel_ms"], d["clocks"])
    for k, v in d.get("extra_workloads", {}).items():
        if isinstance(v, dict):
            print(" ", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "e2e_value", "unit", "ms_per_step", "kernel_ms", "wall_s", "evaluations", "z0_norm", "ms_per_sweep_batch", "ms_per_eval")})
        else:
            print(" ", k, v)
    r = json.loads([l for l in open("gpurun_out/val_bench_reference.json") if l.startswith("{")][-1])
    print("reference arm", round(r["value"], 3), r["cpu_baseline"]["cores"], "same config:", r["config"] == d["config"])
except Exception as ex:
    print("no line", repr(ex))
PY
for spec in sv12:200 sv16:200 sv22:100 sv24:30 sv28:5; do
  wl=${spec%%:*}; st=${spec##*:}
  timeout -k 10 200 python bench.py --workload $wl --steps $st --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/val_bench_$wl.json 2> gpurun_out/val_bench_$wl.err
  python - <<PY | tee -a $S
import json
try:
    d = json.load(open("gpurun_out/val_bench_$wl.json"))
    print("$wl", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["kernel_ms"], "launches/step", d["gpu_launches"] / d["steps"], "frac", round(d["roofline"]["frac"], 3))
except Exception as ex:
    print("$wl: no line", ex)
PY
done
if [ -n "$VAL_SKIP_NCU" ]; then echo "ncu passes skipped (VAL_SKIP_NCU)" | tee -a $S; exit 0; fi
timeout -k 10 120 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/val_plain_sv20.json 2> gpurun_out/val_plain_sv20.err &&
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/val_launches_sv20.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/val_ncu_launches.log 2>&1
echo "ncu launch list rc=$?" | tee -a $S
timeout -k 10 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:dense_pass_kernel<.int.2>" -s 20 -c 5 -o gpurun_out/val_prof_sv20 -f python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/val_ncu_sv20.log 2>&1
echo "ncu sv20 rc=$?" | tee -a $S
timeout -k 10 120 python bench.py --workload sv28 --steps 1 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/val_plain_sv28.json 2> gpurun_out/val_plain_sv28.err &&
timeout -k 10 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:dense_pass_kernel<.int.2>" -s 25 -c 3 -o gpurun_out/val_prof_sv28 -f python bench.py --workload sv28 --steps 1 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/val_ncu_sv28.log 2>&1
echo "ncu sv28 rc=$?" | tee -a $S
