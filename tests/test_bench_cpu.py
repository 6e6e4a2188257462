"""
bench.py contract checks that need no GPU: the reference arm (CPU port of the reference algorithm
on the host cores) prints one JSON line with the agreed keys, and the roofline block of the GPU
arm carries bound / achieved / peak / unit / frac / traffic plus the true-DRAM and FP64 figures.
"""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_line():
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "sv12",
         "--steps", "1", "--warmup", "0"],
        capture_output=True, text=True, check=True, timeout=300,
    ).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "objective+gradient evals/sec" and d["unit"] == "evals/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["config"]["workload"] == "sv12" and d["config"]["num_qubits"] == 12
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_roofline_block_keys_and_arithmetic():
    import bench

    n, layers = 28, 4
    P = bench.pair_runs(n, layers)
    assert P == 27 * 4 + 14  # (n - 1) L + n // 2
    r = bench.roofline_block("sv28", n, P, passes_grad=20, passes_dag=13, stages_grad=128, stages_dag=122,
                             grad_s=0.121, obj_s=0.0425)
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "dram", "fp64"):
        assert key in r
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    # algorithmic bytes of one gradient sweep = 4 vectors x 16 B x 2^n x P, spread over its launches
    assert abs(r["algorithmic_bytes_per_launch"] * 20 - 4 * 16 * 2.0**n * P) < 1
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["dram"]["bytes_per_launch"] - 4 * 16 * 2.0**n) < 1
    assert 0 < r["fp64"]["frac_eval"] < 1 and r["traffic"] == 17.12e9
