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
    r = bench.roofline_block("sv28", n, P, stages_grad=128, stages_dag=122, grad_kernel_s=0.121,
                             obj_kernel_s=0.0425, step_s=0.165, grad_launches=1)
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "hbm_pairrun", "eval_frac_fp64"):
        assert key in r
    # the pair-run byte figure exceeds the HBM peak, so the line names the FP64 pipe as the roof
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s"
    flops = 128 * 2.0**n / 32 * 6 * 512
    assert abs(r["achieved"] - flops / 0.121 / 1e12) < 1e-9 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert 0 < r["frac"] < 1 and 30 < r["peak"] < 45
    h = r["hbm_pairrun"]
    assert abs(h["algorithmic_bytes_per_sweep"] - 4 * 16 * 2.0**n * P) < 1 and h["frac"] > 1
    # whole-step figures are taken on the step time, not on the kernel time
    assert abs(r["eval_frac_hbm_pairrun"] - 6 * 16 * 2.0**n * P / 0.165 / 1e9 / h["peak_gbs"]) < 1e-9
    tr = bench.ncu_traffic("sv28")
    assert r["traffic"] == (tr["dram_bytes_per_launch"] if tr else None)


def test_both_arms_describe_the_same_config():
    import bench

    a = bench.workload_config("sv20", 20, 2)
    assert a["workload"] == "sv20" and a["num_qubits"] == 20 and a["pair_runs"] == 48 and a["gate_units"] == 164
    assert set(a) == {"workload", "num_qubits", "layers", "ansatz", "num_thetas", "pair_runs", "gate_units",
                      "target", "parallelism"}
