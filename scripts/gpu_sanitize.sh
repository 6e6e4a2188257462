#!/bin/bash
# compute-sanitizer passes over the small-state paths (one GPU; never on a multi-rank command).
# memcheck: state-vector smoke (dense engine, n = 6/10/13), the objective-class tests (legacy + dense +
# one-submission evaluation), the small MPS tests (swap network, splits).  racecheck: the smoke only.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/san_summary.txt; : > $S
timeout -k 5 90 python scripts/gpu_smoke.py 6 10 13 > gpurun_out/san_plain.log 2>&1; echo "plain smoke rc=$?" | tee -a $S
grep -q "smoke ok" gpurun_out/san_plain.log || { tail -5 gpurun_out/san_plain.log; exit 1; }
CS="compute-sanitizer --error-exitcode 99 --print-limit 20"
timeout -k 10 600 $CS --tool memcheck python scripts/gpu_smoke.py 6 10 13 > gpurun_out/san_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?" | tee -a $S
grep "ERROR SUMMARY" gpurun_out/san_memcheck_smoke.log | tee -a $S
timeout -k 10 900 $CS --tool memcheck python -m pytest tests/test_objectives_gpu.py -q -x > gpurun_out/san_memcheck_objectives.log 2>&1; echo "memcheck objectives rc=$?" | tee -a $S
grep "ERROR SUMMARY\|passed\|failed" gpurun_out/san_memcheck_objectives.log | tee -a $S
timeout -k 10 900 $CS --tool memcheck python -m pytest tests/test_mps_gpu.py -q -x -k "roundtrip or non_adjacent or truncation_behaviour or set_product_site or gate_by_gate" > gpurun_out/san_memcheck_mps.log 2>&1; echo "memcheck mps rc=$?" | tee -a $S
grep "ERROR SUMMARY\|passed\|failed" gpurun_out/san_memcheck_mps.log | tee -a $S
timeout -k 10 900 $CS --tool racecheck python scripts/gpu_smoke.py 6 10 13 > gpurun_out/san_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?" | tee -a $S
grep "RACECHECK SUMMARY\|ERROR SUMMARY" gpurun_out/san_racecheck_smoke.log | tee -a $S
timeout -k 10 600 $CS --tool synccheck python scripts/gpu_smoke.py 6 10 > gpurun_out/san_synccheck_smoke.log 2>&1; echo "synccheck smoke rc=$?" | tee -a $S
grep "ERROR SUMMARY" gpurun_out/san_synccheck_smoke.log | tee -a $S
