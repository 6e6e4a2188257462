#!/usr/bin/env python
"""
bench.py -- objective + gradient evaluations per second of the state-vector ASP hot path
(BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sv20|sv24|sv28|sv12]
                  [--impl b200|reference]

A "step" is one objective(theta) followed by one gradient(theta) at the same theta (one V^H
sweep + hs gather, then one forward w/z gradient sweep; the leading flip-state is |0>, i.e. the
single-term gradient of SURVEY.md section 8(d)), on a 2nd-order TrotterAnsatz with a synthetic
"near" target V(theta*)|0>, theta = theta* + small perturbation.  Both halves are enqueued in one go
(aqc_sv_eval_begin, what the objective class does once it has seen scipy's fun / jac pattern); the
amplitudes and the gradient are both read back inside the timed region.

  value      evals/s with everything resident in HBM, timed with CUDA events on the engine's
             stream (aqc_sv_timer_*), max over ranks.  N > 1: independent evaluations, one per
             GPU (multistart / time horizons -- the path shards with no collective), weak scaling.
  e2e        the same step through the public objective class (host thetas in, f and gradient
             out), wall clock, copies included.
  roofline   dominant kernel (gradient tile pass): algorithmic bytes at pair-run granularity
             (4 * 16 * 2^n * P bytes per sweep) / its measured device time, against
             the measured HBM peak of MEASURED_PEAKS.json.
  cpu_baseline  the C/OpenMP oracle port (oracle/sv_oracle.c) on the host cores, rank 0, N = 1.
  --impl reference  times that CPU port alone (the reference is pure Python and does not travel
             to the GPU box; BASELINE.md section 2 has its own NumPy timings).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {  # name -> (num_qubits, layers)
    "sv12": (12, 2),
    "sv16": (16, 2),
    "sv20": (20, 2),
    "sv22": (22, 2),
    "sv24": (24, 4),
    "sv28": (28, 4),
    "sv30": (30, 4),
    "sv31": (31, 4),
}
METRIC = "objective+gradient evals/sec"
UNIT = "evals/s"
L2_BYTES = 126 * 1024 * 1024


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path, encoding="utf-8") as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peak():
    """
    FP64 pipe peak (DMMA and DFMA share it) measured on this pool's B200 with scripts/ubench_fp64.cu;
    the record is profiles/fp64_peak.json (MEASURED_PEAKS.json holds no FP64 figure).
    """
    path = os.path.join(ROOT, "profiles", "fp64_peak.json")
    try:
        with open(path, encoding="utf-8") as fh:
            rec = json.load(fh)
        return float(rec["fp64_tflops"]), f"measured (profiles/fp64_peak.json: {rec.get('source', '')})"
    except (OSError, KeyError, ValueError):
        return 148 * 1.965e9 * 128 / 1e12, "nominal (148 SMs x 1965 MHz x 128 flop/clk)"


def ncu_traffic(workload):
    """
    dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    `ncu --set full` capture that scripts/ncu_summary.py condensed into profiles/ncu_traffic.json
    (with the git hash of the build that was profiled); None if this workload was not captured.
    """
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path, encoding="utf-8") as fh:
            rec = json.load(fh).get(workload)
        return rec if rec else None
    except (OSError, ValueError):
        return None


def roofline_block(workload, n, P, stages_grad, stages_dag, grad_kernel_s, obj_kernel_s, step_s, grad_launches):
    """
    Roofline of the dominant kernel, the gradient sweep kernel (one launch runs every tile pass of the
    sweep).  The SURVEY 8(d) pair-run byte figure exceeds the HBM peak because a tile pass fuses several
    pair-runs per DRAM round trip, so the binding roof is the FP64 pipe:
      frac / achieved / peak : DMMA flops the sweep issues (6 m8n8k4 = 6 x 512 flop per 8 amplitude
          quadruples and stage) / CUDA-event time of the sweep, against the measured FP64 pipe peak;
      hbm_pairrun : the 8(d) contract figure, ALGORITHMIC bytes at pair-run granularity
          (4 * 16 * 2^n * P per gradient sweep) / sweep time, against the measured HBM copy peak;
      traffic: ncu dram bytes per launch of the profiled build (profiles/ncu_traffic.json) or null;
      eval_frac_*: the same two figures for the whole step (V^H sweep + gradient sweep) on ms_per_step.
    """
    hbm_peak, hbm_src = measured_peaks()
    f64_peak, f64_src = fp64_peak()
    V = 16.0 * 2**n
    dmma_grad = stages_grad * (2**n / 32.0) * 6 * 512.0
    dmma_apply = stages_dag * (2**n / 32.0) * 2 * 512.0
    alg_bytes = 4.0 * V * P
    achieved = dmma_grad / grad_kernel_s / 1e12
    tr = ncu_traffic(workload)
    return {
        "bound": "fp64", "kernel": "gradient sweep kernel (tile passes: cp.async tile staging, DMMA stage matrices)",
        "achieved": achieved, "peak": f64_peak, "unit": "TFLOP/s", "frac": achieved / f64_peak,
        "peak_source": f64_src, "launch_ms": grad_kernel_s * 1e3 / max(1, grad_launches),
        "launches_per_sweep": grad_launches, "dmma_flop_per_sweep": dmma_grad,
        "traffic": tr["dram_bytes_per_launch"] if tr else None, "traffic_capture": tr,
        "hbm_pairrun": {"algorithmic_bytes_per_sweep": alg_bytes, "achieved_gbs": alg_bytes / grad_kernel_s / 1e9,
                        "peak_gbs": hbm_peak, "frac": alg_bytes / grad_kernel_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                        "note": "SURVEY 8(d) pair-run granularity; > 1 because a tile pass fuses several pair-runs"},
        "eval_frac_fp64": (dmma_grad + dmma_apply) / step_s / 1e12 / f64_peak,
        "eval_frac_hbm_pairrun": (6.0 * V * P / step_s) / 1e9 / hbm_peak,
        "fp64_eval_kernels": {"dmma_tflops": (dmma_grad + dmma_apply) / (grad_kernel_s + obj_kernel_s) / 1e12,
                              "frac": (dmma_grad + dmma_apply) / (grad_kernel_s + obj_kernel_s) / 1e12 / f64_peak},
    }


def make_circuit(n, layers):
    from aqc_research_b200 import circuit_structures as cs
    from aqc_research_b200.parametric_circuit import TrotterAnsatz

    return TrotterAnsatz(n, cs.make_trotter_like_circuit(n, layers), True)


def workload_config(workload, n, layers):
    """The `config` object of the JSON line: identical in the GPU arm and the reference arm."""
    circ = make_circuit(n, layers)
    return {
        "workload": workload, "num_qubits": n, "layers": layers, "ansatz": "TrotterAnsatz 2nd order",
        "num_thetas": circ.num_thetas, "pair_runs": pair_runs(n, layers), "gate_units": gate_units(circ),
        "target": "synthetic complex128 state of 2^n amplitudes (the cost does not depend on its values)",
        "parallelism": "independent evaluations, one per GPU, no collective (multistart / horizons)",
    }


def pair_runs(n, layers):
    return (n - 1) * layers + n // 2  # P for a 2nd-order TrotterAnsatz


def gate_units(circ):
    return circ.num_qubits + circ.num_blocks + circ.half_layer_num_blocks


class ClockSampler:
    """
    Samples SM clocks / clock-event reasons of one GPU DURING the timed region: NVML in-process
    (5 ms period, so even a few-millisecond region gets samples), nvidia-smi as the fallback.
    """

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.rows = []  # (sm_mhz, sm_max_mhz, [reason names])
        self.proc = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.source = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if ids and self.index < len(ids) and ids[self.index].isdigit():
            return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            masks = [(nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                     (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                     (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                     (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                        bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                        self.rows.append((sm, smax, [n for m, n in masks if bits & m]))
                    except nv.NVMLError:
                        pass
                    self.stop_flag.wait(0.005)

            self.source = "nvml"
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:  # NVML missing: fall back to the CLI
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self._physical_index()}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.source = "nvidia-smi"
        self.thread = threading.Thread(target=self._read_cli, daemon=True)
        self.thread.start()

    def _read_cli(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) < 6:
                continue
            try:
                self.rows.append((float(r[0]), float(r[1]),
                                  [n for n, v in zip(self.NAMES, r[2:6]) if v.lower().startswith("active")]))
            except ValueError:
                continue

    def stop(self):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (NVML / nvidia-smi)"],
                    "samples": 0}
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        sm = [r[0] for r in self.rows]
        smax = [r[1] for r in self.rows]
        reasons = sorted({n for r in self.rows for n in r[2]})
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(smax)) if smax else None,
            "reasons": reasons,
            "samples": len(sm),
            "source": self.source,
        }


def cpu_port_eval_seconds(n, layers, seed):
    """One objective + single-term gradient with the C/OpenMP oracle port; returns seconds."""
    from aqc_research_b200 import utils
    from oracle import c_oracle as C

    C.use_all_cores()  # torchrun exports OMP_NUM_THREADS=1

    circ = make_circuit(n, layers)
    rng = np.random.RandomState(seed)
    th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    y = rng.rand(2**n) + 1j * rng.rand(2**n)
    y /= np.linalg.norm(y)
    t0 = time.perf_counter()
    z0 = C.apply_v(circ, th, y, dagger=True)
    hs = z0[[0] + [1 << q for q in range(n)]]
    w = np.zeros(2**n, dtype=np.complex128)
    w[0] = 1
    C.grad_sweep(circ, th, w, z0, inplace=True)
    dt = time.perf_counter() - t0
    del utils, hs
    return dt, gate_units(circ)


def cpu_baseline(n, layers, seed=7):
    """
    Bounded CPU sample: the full workload for n <= 21; above that the same circuit depth at
    n_s = 21 qubits, scaled by 2^(n - n_s) * G(n)/G(n_s) (cost is linear in amplitudes x gate units).
    """
    from oracle import c_oracle as C

    n_s = min(n, 21)
    secs, g_s = cpu_port_eval_seconds(n_s, layers, seed)
    g_full = gate_units(make_circuit(n, layers))
    scale = (2.0 ** (n - n_s)) * g_full / g_s
    est = secs * scale
    sample = (f"1 full eval at n={n_s}, L={layers} ({secs:.2f} s)" if n_s == n else
              f"1 eval at n={n_s}, L={layers} ({secs:.2f} s) scaled x{scale:.1f} "
              f"(2^{n - n_s} amplitudes x gate units {g_full}/{g_s})")
    return {"value": 1.0 / est, "unit": UNIT, "cores": C.num_threads(), "kind": "port",
            "sample": sample}


def run_reference_arm(args, n, layers):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(n, layers, seed=100 + i)
        if i >= args.warmup:
            vals.append(cb)
    value = float(np.mean([v["value"] for v in vals]))
    cb = dict(vals[-1])
    cb["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128 (f64)",
        "data": "synthetic",
        "config": workload_config(args.workload, n, layers),
        "details": {"note": "CPU port of the reference algorithm (oracle/sv_oracle.c, OpenMP); the "
                            "reference itself is pure Python/NumPy, see BASELINE.md section 2"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measure_gpu(n, layers, steps, warmup, device, flush_l2, sampler=None):
    """Returns dict with per-step device ms list, kernel split, e2e seconds list, launches."""
    from aqc_research_b200.model_sp_lhs.objective_lhs_sur_max import SpSurrogateObjectiveMax
    from aqc_research_b200.model_sp_lhs.objective_base import SLOT_TARGET, SLOT_VH_TARGET, SLOT_W, SLOT_Z

    circ = make_circuit(n, layers)
    rng = np.random.RandomState(1234 + n)
    th_star = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    params = dict(num_qubits=n, max_flips=1, maxiter=40, verbose=0, enable_optim_stats=False,
                  num_simulations=1, trunc_thr=1e-6, state_prep_func=None, device=device)
    objv = SpSurrogateObjectiveMax(user_parameters=params, circ=circ, front_layer=True)
    ws = objv.workspace
    # near target: V(theta*)|0>, generated on the device
    ws.set_basis(SLOT_TARGET, 0)
    ws.apply(th_star, SLOT_TARGET, SLOT_TARGET, dagger=False)
    objv._target = "device-near-target"  # resident; nothing to upload
    delta = 0.02
    while True:
        th = th_star + delta * np.pi * (2 * rng.rand(circ.num_thetas) - 1)
        objv.objective(th)
        if objv.max_no == 0 or delta < 1e-6:
            break
        delta *= 0.5
    fidelity = objv.fidelity
    idx = np.array([0] + [1 << q for q in range(n)], dtype=np.int64)

    flush = None
    if flush_l2:
        import torch
        flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=f"cuda:{device}")

    def do_flush():
        if flush is not None:
            import torch
            flush.fill_(1)
            torch.cuda.synchronize(device)

    step_ms, obj_ms, grad_ms, launches, grad_launches = [], [], [], 0, 1
    for it in range(warmup + steps):
        if it == warmup and sampler:
            sampler.start()
        th_i = th + 1e-3 * np.cos(np.arange(th.size) + it)
        do_flush()
        # one evaluation = V^H sweep + hs gather + gradient sweep from |0>, enqueued in one go
        # (aqc_sv_eval_begin); hs is read back as soon as the gather is done, the gradient at the end
        ws.timer_start()
        hs_i = ws.eval_begin(th_i, SLOT_TARGET, SLOT_VH_TARGET, idx, x_basis=0, w=SLOT_W, z=SLOT_Z)
        n_l = ws.last_num_launches
        g_i = ws.grad_end()
        ms = ws.timer_stop()
        o_ms, g_ms = ws.eval_times()
        o_l, g_l = n_l // 2, n_l - n_l // 2
        assert np.isfinite(hs_i[0, 0]) and np.isfinite(g_i[0, 0])
        if it >= warmup:
            step_ms.append(ms)
            obj_ms.append(o_ms)
            grad_ms.append(g_ms)
            launches += o_l + g_l
            grad_launches = g_l
    # end-to-end through the public objective class (host thetas in, f and gradient out)
    e2e_s = []
    for it in range(warmup + steps):
        th_i = th + 1e-3 * np.sin(np.arange(th.size) + it)
        do_flush()
        t0 = time.perf_counter()
        f = objv.objective(th_i)
        g = objv.gradient(th_i)
        dt = time.perf_counter() - t0
        if it >= warmup:
            e2e_s.append(dt)
    assert objv.max_no == 0 and np.isfinite(f) and np.all(np.isfinite(g))
    # the max_no != 0 case (two weighted gradient terms) is reported separately.  A
    # random target puts the leading flip state elsewhere; both terms are taken in one sweep.
    two_term = None
    ws.fill_random(SLOT_TARGET, 4321 + n)
    t2_s = []
    for it in range(warmup + min(steps, 10)):
        th_i = th + 1e-3 * np.sin(np.arange(th.size) + it)
        do_flush()
        t0 = time.perf_counter()
        f = objv.objective(th_i)
        g = objv.gradient(th_i)
        dt = time.perf_counter() - t0
        if it >= warmup and objv.max_no != 0:
            t2_s.append(dt)
    if t2_s:
        assert np.isfinite(f) and np.all(np.isfinite(g))
        two_term = {"e2e_value": 1.0 / float(np.mean(t2_s)), "unit": UNIT, "steps": len(t2_s),
                    "target": "random (rand_state distribution, generated on the device)",
                    "gradient_sweeps_per_eval": 1, "scope": "one GPU (rank 0), end to end through the objective class"}
    T = circ.num_thetas
    return {
        "two_term": two_term, "grad_launches": grad_launches,
        "circ": circ, "step_ms": step_ms, "obj_ms": obj_ms, "grad_ms": grad_ms, "launches": launches,
        "e2e_s": e2e_s, "fidelity": fidelity, "passes_grad": ws.num_passes(0), "passes_dag": ws.num_passes(2),
        "stages_grad": ws.num_stages(0), "stages_dag": ws.num_stages(2),
        "h2d": 2 * 8 * T, "d2h": 16 * (n + 1) + 16 * T,
    }


def measure_mps(n=50, chi=64, layers=20, steps=3, warmup=1, device=0, with_cpu=True, physical=False):
    """
    BASELINE.json configs[3]: MPS fidelity/gradient, 50 qubits, bond dimension 64, Trotter ansatz
    depth 20, trunc_thr = 1e-6.  One step = objective (V^H target + n+1 overlaps) + one gradient sweep.
    Two regimes:
      saturated (default) -- target = a random-angle Trotter circuit applied to the Neel state, ansatz
          angles ~ U(-pi, pi): every bond sits at the chi cap.  This is the COST upper bound of the
          configuration; numerically every split discards weight of order 0.1, so the values computed
          are not a meaningful approximation of anything (reported: ||z0||, discarded weight, cap hits);
      physical -- target = Trotter-evolved Neel state (evolution time 3), ansatz angles near the Trotter
          point: the production regime (discarded weight ~1e-6 per split, no cap hit).
    """
    from aqc_research_b200.mps_engine import MpsWorkspace

    circ = make_circuit(n, layers)
    rng = np.random.RandomState(50)
    if physical:
        from aqc_research_b200.model_sp_lhs.trotter import trotter as trotop

        th_t = trotop.init_ansatz_to_trotter(circ, np.zeros(circ.num_thetas), evol_time=3.0, delta=1.0)
        th = th_t + 0.01 * (2 * rng.rand(circ.num_thetas) - 1)
    else:
        th_t = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
        th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    ws = MpsWorkspace(circ, num_slots=4, chi_max=chi, trunc_thr=1e-6, device=device)
    neel = sum(1 << q for q in range(0, n, 2))
    ws.set_product(0, neel)
    ws.apply(th_t, 0, 0, dagger=False)
    bonds = [int(l.size) for l in ws.download(0)[1]]
    idx = np.array([neel] + [neel ^ (1 << q) for q in range(n)], dtype=np.int64)
    obj_ms, grad_ms, wall, launches = [], [], [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ws.objective(th, 0, 1, idx)
        o, lo = ws.last_kernel_ms, ws.last_num_launches
        trunc_obj = ws.truncation_stats() if it == warmup + steps - 1 else None
        ws.grad(th, x_basis=neel, z0=1, w=2, z=3)
        g, lg = ws.last_kernel_ms, ws.last_num_launches
        dt = time.perf_counter() - t0
        if it >= warmup:
            obj_ms.append(o), grad_ms.append(g), wall.append(dt)
            launches += lo + lg
    trunc_grad = ws.truncation_stats()
    z0_norm = float(abs(ws.dot(1, 1)) ** 0.5)
    ws.close()
    nb_tot = circ.num_blocks + circ.half_layer_num_blocks
    dots = 3 * n + 4 * nb_tot
    out = {
        "workload": f"mps n={n} chi={chi} layers={layers} (2nd-order Trotter, trunc_thr=1e-6), "
                    + ("physical regime: Trotter-evolved target, angles near the Trotter point" if physical
                       else "saturated regime: random angles, every bond at the cap"),
        "num_thetas": circ.num_thetas, "target_bond_dims_max": max(bonds), "steps": steps, "warmup": warmup,
        "value": 1e3 / float(np.mean(obj_ms) + np.mean(grad_ms)), "unit": UNIT,
        "e2e_value": 1.0 / float(np.mean(wall)),
        "kernel_ms": {"objective": float(np.mean(obj_ms)), "gradient": float(np.mean(grad_ms))},
        "gpu_launches": launches,
        # what the timed computation discards (saturated regime: every bond sits at the chi cap, so the
        # numbers are large -- the parity of truncated results is pinned at smaller sizes)
        "z0_norm": z0_norm, "truncation": {"objective": trunc_obj, "gradient": trunc_grad},
        "dominant_kernel": "mps_svd_kernel (block one-sided Jacobi, FP64 FMA pipe; ~85-89% of device time, profiles/r01_ncu_mps50.md)",
    }
    if with_cpu:
        # reference-equivalent cost: its gradient makes `dots` full-chain mps_dot calls (plus one
        # qiskit-aer run per gate, not reproducible here); time the oracle's restatement of mps_dot
        from oracle import mps_oracle as M

        rs = np.random.RandomState(1)
        a, b = M.random_mps(n, chi, rs), M.random_mps(n, chi, rs)
        t0 = time.perf_counter()
        M.mps_dot(a, b)
        t_dot = time.perf_counter() - t0
        out["cpu_baseline"] = {
            "value": 1.0 / (dots * t_dot), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"1 mps_dot at n={n}, chi={chi} ({t_dot:.3f} s) x {dots} inner products per gradient "
                      "(reference algorithm; excludes its qiskit-aer gate runs)",
        }
    return out


def measure_mat7(batch=64, layers=40, steps=5, warmup=2, device=0, with_cpu=True):
    """
    BASELINE.json configs[2]: unitary AQC (core_op_matrix path), 7 qubits, random SU target,
    cyclic_spin ansatz with 7*layers blocks; `batch` independent starts evaluated per step
    (multistart; across GPUs the starts are simply split).  value = starts * steps / time.
    """
    from aqc_research_b200 import circuit_structures as cs
    from aqc_research_b200.model_sketching.sk_core import BatchedSketchingObjective
    from aqc_research_b200.parametric_circuit import ParametricCircuit

    n = 7
    circ = ParametricCircuit(n, "cx", cs.create_ansatz_structure(n, "cyclic_spin", "full", n * layers))
    rng = np.random.RandomState(7)
    q, r = np.linalg.qr(rng.randn(2**n, 2**n) + 1j * rng.randn(2**n, 2**n))
    target = np.ascontiguousarray(q * (np.diag(r) / np.abs(np.diag(r))))
    objv = BatchedSketchingObjective(circ, target, batch=batch, device=device)
    ths = np.pi * np.clip(rng.randn(batch, circ.num_thetas), -1, 1)
    ws = objv.workspace
    ms, launches = [], 0
    for it in range(warmup + steps):
        ws.timer_start()
        f, g = objv.evaluate(ths + 1e-3 * it)
        t = ws.timer_stop()
        if it >= warmup:
            ms.append(t)
    flops = (24.0 + 80.0) * 2 ** (2 * n) * circ.num_blocks
    out = {
        "workload": f"unitary AQC n={n}, cyclic_spin {layers} layers ({circ.num_blocks} blocks), {batch} starts per step",
        "num_thetas": circ.num_thetas, "steps": steps, "warmup": warmup,
        "value": batch * 1e3 / float(np.mean(ms)), "unit": UNIT, "ms_per_step": float(np.mean(ms)),
        "fp64_tflops": flops * batch / (float(np.mean(ms)) * 1e-3) / 1e12,
    }
    if with_cpu:
        from oracle import c_oracle as C

        C.use_all_cores()
        t0 = time.perf_counter()
        z0 = C.apply_v(circ, ths[0], target.ravel(), dagger=True, log2_cols=n)
        C.grad_sweep(circ, ths[0], np.eye(2**n, dtype=np.complex128).ravel(), z0, log2_cols=n, inplace=False)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": C.num_threads(), "kind": "port",
                               "sample": f"1 objective+gradient of one start ({dt * 1e3:.1f} ms)"}
    return out


def measure_cd7(batch=64, layers=40, sweeps=3, device=0, with_cpu=True):
    """
    SURVEY 8(f) row 2: coordinate descent for unitary AQC (coord_descent_single_sweep), 7 qubits,
    cyclic_spin ansatz with 7*layers blocks, `batch` starts per call; value = sweeps of one start
    per second (every sweep updates all T angles once).
    """
    from aqc_research_b200 import circuit_structures as cs
    from aqc_research_b200.model_sketching.aqc_coord_descent import BatchedCoordinateDescent
    from aqc_research_b200.parametric_circuit import ParametricCircuit

    n = 7
    circ = ParametricCircuit(n, "cx", cs.create_ansatz_structure(n, "cyclic_spin", "full", n * layers))
    rng = np.random.RandomState(11)
    q, r = np.linalg.qr(rng.randn(2**n, 2**n) + 1j * rng.randn(2**n, 2**n))
    target = np.ascontiguousarray(q * (np.diag(r) / np.abs(np.diag(r))))
    ths = np.pi * np.clip(rng.randn(batch, circ.num_thetas), -1, 1)
    opt = BatchedCoordinateDescent(circ, target, batch=batch, device=device)
    opt.sweep(ths, num_sweeps=1)  # warm-up
    t0 = time.perf_counter()
    fobj, _ = opt.sweep(ths, num_sweeps=sweeps)
    dt = time.perf_counter() - t0
    dev_ms = opt.workspace.last_kernel_ms
    opt.close()
    out = {
        "workload": f"coordinate descent n={n}, cyclic_spin {layers} layers ({circ.num_blocks} blocks, "
                    f"{circ.num_thetas} angles), {batch} starts per call",
        "value": batch * sweeps / dt, "unit": "sweeps/s", "ms_per_sweep_batch": dt * 1e3 / sweeps,
        "device_ms_per_sweep_batch": dev_ms / sweeps, "fobj_first_last": [float(fobj[0, 0]), float(fobj[-1, 0])],
    }
    if with_cpu:
        from oracle import sv_oracle as O

        t0 = time.perf_counter()
        O.coord_descent_sweep(circ, ths[0], target)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "sweeps/s", "cores": 1, "kind": "port",
                               "sample": f"1 sweep of one start with the NumPy restatement ({dt:.2f} s)"}
    return out


def measure_sketch(n=12, m=64, layers=4, steps=3, device=0, with_cpu=True):
    """
    SURVEY 8(f) row 3: sketched unitary AQC with the 'eigen' generator (sk_core.py:410-462) at
    n qubits, m sketching vectors: per evaluation U^H Omega and U X (d x d . d x m complex GEMMs on
    the FP64 tensor pipe), thin QR, V^H apply and the gradient sweep on (d, m) matrices.
    """
    from aqc_research_b200 import circuit_structures as cs
    from aqc_research_b200.model_sketching import sk_core
    from aqc_research_b200.parametric_circuit import ParametricCircuit

    d = 1 << n
    circ = ParametricCircuit(n, "cx", cs.create_ansatz_structure(n, "spin", "full", (n - 1) * layers))
    rng = np.random.RandomState(3)
    q, r = np.linalg.qr(rng.randn(d, d) + 1j * rng.randn(d, d))
    target = np.ascontiguousarray(q * (np.diag(r) / np.abs(np.diag(r))))
    th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    np.random.seed(17)
    gen = sk_core.EigenSketchingVectors(m, target)
    objv = sk_core.SketchingObjectiveEx(circ, gen, device=device)
    ws = objv.workspace
    objv.objective_and_gradient(th)  # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        f, g = objv.objective_and_gradient(th)
    dt = (time.perf_counter() - t0) / steps
    ws.target_matmul(2, 3)
    gemm_ms = ws.last_kernel_ms
    ws.upload(2, rng.rand(d, m) + 1j * rng.rand(d, m))
    ws.orthonormalize(2, 3)
    qr_ms = ws.last_kernel_ms
    flops = 8.0 * d * d * m
    out = {
        "workload": f"sketched AQC n={n}, {m} 'eigen' sketching vectors, spin ansatz {circ.num_blocks} blocks",
        "value": 1.0 / dt, "unit": UNIT, "ms_per_eval": dt * 1e3,
        "zgemm_ms": gemm_ms, "zgemm_dmma_tflops": flops / (gemm_ms * 1e-3) / 1e12,
        "zgemm_frac_of_fp64_peak": flops / (gemm_ms * 1e-3) / 1e12 / fp64_peak()[0], "qr_ms": qr_ms,
        "fobj": float(f),
    }
    if with_cpu:
        x = rng.rand(d, m) + 1j * rng.rand(d, m)
        t0 = time.perf_counter()
        y = target @ x
        qq, _ = np.linalg.qr(x)
        dtc = time.perf_counter() - t0
        del y, qq
        out["cpu_baseline"] = {"value": dtc * 1e3, "unit": "ms for one U @ X + one thin QR (NumPy/LAPACK)",
                               "cores": os.cpu_count(), "kind": "port",
                               "sample": "generator linear algebra only (the reference adds a V^H sweep and the gradient)"}
    return out


def measure_lbfgs(n=12, layers=2, maxiter=40, evol_time=1.0, device=0, with_cpu=True):
    """
    BASELINE.json configs[1] end to end: the optimisation loop the reference
    runs (optimizer.py:585-590 -> scipy L-BFGS-B, fun = objv.objective, jac = objv.gradient) on
    the Neel state, target = Trotter evolution with 10x finer steps (generated on the device),
    theta_0 = init_ansatz_to_trotter.  Wall clock of the whole minimize() call, host side included.
    """
    from scipy.optimize import minimize
    from aqc_research_b200.model_sp_lhs.objective_lhs_sur_max import SpSurrogateObjectiveMax
    from aqc_research_b200.model_sp_lhs.trotter import trotter as trot

    circ = make_circuit(n, layers)
    th0 = trot.init_ansatz_to_trotter(circ, np.zeros(circ.num_thetas), evol_time=evol_time, delta=1.0)
    target = trot.trotter_state(n, evol_time=evol_time, num_steps=10 * layers, delta=1.0,
                                second_order=True, ini_state=trot.neel_init_state(n))
    params = dict(num_qubits=n, max_flips=1, maxiter=maxiter, verbose=0, enable_optim_stats=False,
                  num_simulations=1, trunc_thr=1e-6, state_prep_func=trot.neel_init_state, device=device)
    best = None
    for _ in range(2):  # the first run pays for allocation and table upload
        objv = SpSurrogateObjectiveMax(user_parameters=params, circ=circ, front_layer=True)
        objv.set_target(target)
        t0 = time.perf_counter()
        res = minimize(fun=objv.objective, x0=th0.copy(), jac=objv.gradient, method="L-BFGS-B",
                       options=dict(maxfun=5 * maxiter, maxiter=maxiter, ftol=10 * np.finfo(float).eps, eps=1e-8))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out = {
        "workload": f"L-BFGS-B run, n={n}, {layers} layers 2nd-order Trotter ansatz, maxiter={maxiter}, Neel state, "
                    f"Trotter target (t={evol_time})",
        "num_thetas": circ.num_thetas, "iterations": int(res.nit), "evaluations": int(res.nfev),
        "wall_s": best, "value": float(res.nfev) / best, "unit": UNIT,
        "final_fobj": float(res.fun), "final_fidelity": float(objv.fidelity),
    }
    if with_cpu:
        secs = min(cpu_port_eval_seconds(n, layers, 11 + k)[0] for k in range(3))  # first call starts the threads
        from oracle import c_oracle as C

        out["cpu_baseline"] = {"value": 1.0 / secs, "unit": UNIT, "cores": C.num_threads(), "kind": "port",
                               "sample": f"1 objective+gradient, best of 3 ({secs * 1e3:.2f} ms); the run needs {int(res.nfev)}"}
    return out


def measure_sharded(base_qubits, layers, steps, warmup, local_rank, world):
    """
    BASELINE.json configs[4], second half: ONE state vector over `world` GPUs (global-qubit
    sharding), weak scaling: n = base_qubits + log2(world) qubits, i.e. a constant 2^base_qubits
    amplitudes per GPU.  One step = objective + single-term gradient of the whole sharded state;
    every layout switch is fused into the last tile pass of its epoch (NVLink peer stores).  Wall
    clock per step between two barriers of the process group (= max over ranks).
    """
    import torch.distributed as dist
    from aqc_research_b200.sharded import DistComm, GpuShardBackend, ShardedStateVector

    g = world.bit_length() - 1
    n = base_qubits + g
    circ = make_circuit(n, layers)
    comm = DistComm()
    be = GpuShardBackend(circ, g, comm.rank, local_rank, num_slots=5)
    sv = ShardedStateVector(circ, comm, be)
    rng = np.random.RandomState(4321)
    th_star = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    th = th_star + 0.02 * np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    sv.set_basis("w", 0)
    sv.apply(th_star, "w", "target", dagger=False)  # near target V(theta*)|0>
    idx = np.array([0] + [1 << q for q in range(n)], dtype=np.int64)
    times = []
    for it in range(warmup + steps):
        sv.exchange_ms = sv.compute_ms = 0.0
        dist.barrier()
        t0 = time.perf_counter()
        hs = sv.objective(th, idx)
        grad = sv.grad(th, 0)
        dist.barrier()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append((dt, sv.compute_ms, sv.exchange_ms))
    assert np.all(np.isfinite(grad))
    t = np.array(times)
    # bytes every GPU sends per evaluation: each layout switch moves (1 - 1/world) of a 16 * 2^base_qubits
    # byte shard; the V^H sweep switches one vector per epoch boundary (and once more to return to the
    # rest layout), the gradient sweep two vectors per epoch boundary
    e_g, e_d = be.num_epochs(0), be.num_epochs(2)
    last_d = be.epoch_layout(2, e_d - 1)
    switches = (e_d - 1 + (1 if last_d != 0 else 0)) + 2 * (e_g - 1)
    sent = switches * 16.0 * 2**base_qubits * (1.0 - 1.0 / world)
    res = {"num_qubits": n, "layers": layers, "num_thetas": circ.num_thetas, "p2p_exchange": bool(sv.p2p),
           "fused_push": bool(sv.push), "epochs": {"gradient": e_g, "vh_apply": e_d},
           "s_per_step": float(t[:, 0].mean()), "compute_ms": float(t[:, 1].mean()),
           "barrier_wait_ms": float(t[:, 2].mean()), "fidelity": float(abs(hs[0]) ** 2),
           "vector_switches_per_eval": switches, "nvlink_bytes_sent_per_gpu_per_eval": sent,
           "nvlink_gbs_per_gpu_over_whole_step": sent / float(t[:, 0].mean()) / 1e9}
    sv.close()
    return res


def sharded_parity(n, layers, local_rank, world, seed=2024):
    """
    Self-test of the sharded CUDA path at real shard sizes: objective amplitudes
    and gradient of an n-qubit state over `world` GPUs against the SAME evaluation on one GPU
    (rank 0), north-star tolerance 1e-10 relative.  The target is generated on the device from the
    logical amplitude index, so both runs see the same vector.
    """
    from aqc_research_b200.engine import SvWorkspace
    from aqc_research_b200.sharded import DistComm, GpuShardBackend, ShardedStateVector

    g = world.bit_length() - 1
    circ = make_circuit(n, layers)
    comm = DistComm()
    be = GpuShardBackend(circ, g, comm.rank, local_rank, num_slots=5)
    sv = ShardedStateVector(circ, comm, be)
    rng = np.random.RandomState(seed)
    th = np.pi * (2 * rng.rand(circ.num_thetas) - 1)
    idx = np.array([0] + [1 << q for q in range(n)], dtype=np.int64)
    sv.set_target_random(seed)
    hs = sv.objective(th, idx)
    xb = int(idx[n // 2])
    grad = sv.grad(th, xb)
    fused = bool(sv.push)
    sv.close()
    out = None
    if comm.rank == 0:
        ws = SvWorkspace(circ, num_slots=4, device=local_rank)
        ws.fill_random(0, seed)
        hs1 = ws.objective(th, 0, 1, idx)[0]
        g1 = ws.grad(th, x_basis=xb, z0=1, w=2, z=3)[0]
        ws.close()
        e_hs = float(np.linalg.norm(hs - hs1) / np.linalg.norm(hs1))
        e_g = float(np.linalg.norm(grad - g1) / np.linalg.norm(g1))
        out = {"num_qubits": n, "layers": layers, "gpus": world, "amplitudes_per_gpu": 2 ** (n - g),
               "fused_push": fused, "rel_err_hs_vs_1gpu": e_hs, "rel_err_grad_vs_1gpu": e_g,
               "tolerance": 1e-10, "parity_ok": bool(e_hs < 1e-10 and e_g < 1e-10)}
    comm.barrier()
    return out


def run_sharded_bench(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    base, layers = args.shard_qubits, 4
    if world == 1:
        res = measure_gpu(base, layers, args.steps, args.warmup, 0, False, None)
        out = {"num_qubits": base, "s_per_step": float(np.mean(res["step_ms"])) * 1e-3, "barrier_wait_ms": 0.0,
               "compute_ms": float(np.mean(res["step_ms"])), "epochs": {"gradient": 1, "vh_apply": 1},
               "layers": layers, "num_thetas": res["circ"].num_thetas, "p2p_exchange": False}
    else:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        out = measure_sharded(base, layers, args.steps, args.warmup, local_rank, world)
    if rank == 0:
        n = out["num_qubits"]
        P = pair_runs(n, layers)
        peak, peak_src = measured_peaks()
        value = 1.0 / out["s_per_step"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": out["s_per_step"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c128 (f64)", "data": "synthetic",
            "config": {"workload": "svshard", "num_qubits": n, "layers": layers,
                       "ansatz": "TrotterAnsatz 2nd order", "num_thetas": out["num_thetas"],
                       "amplitudes_per_gpu": 2**args.shard_qubits, "parallelism": f"one state over {world} GPUs (global-qubit sharding)",
                       "epochs": out["epochs"], "p2p_exchange": out["p2p_exchange"],
                       "l2": "inputs larger than L2"},
            "roofline": {"bound": "hbm", "kernel": "dense_pass_kernel<2> (gradient tile pass; the last pass of every epoch stores its tiles into the peers' HBM over NVLink)",
                         "achieved": 96.0 * 2**n * P * value / 1e9 / world, "peak": peak, "unit": "GB/s",
                         "frac": 96.0 * 2**n * P * value / 1e9 / world / peak, "traffic": None,
                         "peak_source": peak_src,
                         "note": "whole evaluation (pair-run algorithmic bytes per GPU) incl. the layout switches"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 16 * out["num_thetas"],
                    "d2h_bytes_per_step": 16 * (n + 1) + 16 * out["num_thetas"]},
            "kernel_ms": {"compute": out["compute_ms"], "barrier_wait": out.get("barrier_wait_ms", 0.0)},
            "sharded": out,
            "gpu_launches": None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=os.environ.get("AQC_BENCH_WORKLOAD", "sv20"),
                    choices=sorted(WORKLOADS) + ["svshard"])
    ap.add_argument("--shard-qubits", type=int, default=28, help="svshard: qubits per GPU shard (log2 amplitudes)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra sv28 measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.workload == "svshard":
        if args.impl == "reference":
            args.workload = "sv28"
        else:
            args.warmup = min(args.warmup, 1)
            run_sharded_bench(args)
            return
    n, layers = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference_arm(args, n, layers)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist  # noqa: F811
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist.barrier()
    from aqc_research_b200 import _lib
    _lib.require_gpu()

    flush_l2 = 3 * 16 * 2**n < 4 * L2_BYTES
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if dist is not None:
        dist.barrier()
    res = measure_gpu(n, layers, args.steps, args.warmup, local_rank, flush_l2, sampler)
    clocks = sampler.stop() if sampler else None

    total_ms = float(np.sum(res["step_ms"]))
    e2e_total = float(np.sum(res["e2e_s"]))
    if dist is not None:
        import torch
        t = torch.tensor([total_ms, e2e_total], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_total = float(t[0]), float(t[1])
        dist.barrier()
    # N > 1: besides the replica headline, the ONE-state-over-N-GPUs path of BASELINE.json configs[4]
    # (global-qubit sharding, weak scaling at 2^shard_qubits amplitudes per GPU) and its parity
    # self-test against a single-GPU evaluation -- every rank takes part
    shard_extra = None
    if dist is not None and not args.no_extra:
        g = world.bit_length() - 1
        smp = ClockSampler(local_rank) if rank == 0 else None
        if smp:
            smp.start()
        shard_extra = {
            "svshard": measure_sharded(args.shard_qubits, 4, 2, 1, local_rank, world),
            "svshard_parity": sharded_parity(24 + g, 2, local_rank, world),
        }
        if smp:
            shard_extra["svshard"]["clocks"] = smp.stop()
        dist.barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    circ = res["circ"]
    P = pair_runs(n, layers)
    V = 16.0 * 2**n
    value = world * args.steps / (total_ms * 1e-3)
    e2e_value = world * args.steps / e2e_total
    grad_s = float(np.mean(res["grad_ms"])) * 1e-3
    obj_s = float(np.mean(res["obj_ms"])) * 1e-3
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c128 (f64)", "data": "synthetic",
        "config": workload_config(args.workload, n, layers),
        "details": {
            "target": "near: V(theta*)|0>, fidelity %.4f" % res["fidelity"],
            "l2": "flushed between timed steps (256 MiB write)" if flush_l2 else "inputs (3 x %.1f GiB) larger than L2" % (V / 2**30),
            "gpus": world,
            "tile_passes": {"gradient": res["passes_grad"], "vh_apply": res["passes_dag"]},
            "stages": {"gradient": res["stages_grad"], "vh_apply": res["stages_dag"]},
        },
        "roofline": roofline_block(args.workload, n, P, res["stages_grad"], res["stages_dag"], grad_s, obj_s,
                                   total_ms * 1e-3 / args.steps, res["passes_grad"]),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": res["h2d"],
                "d2h_bytes_per_step": res["d2h"]},
        "gpu_launches": res["launches"],
        "kernel_ms": {"vh_apply_sweep": obj_s * 1e3, "gradient_sweep": grad_s * 1e3},
        "two_term_case": res["two_term"],
        "clocks": clocks,
    }
    if shard_extra is not None:
        sh = shard_extra["svshard"]
        sh.update({"value": 1.0 / sh["s_per_step"], "unit": UNIT, "steps": 2, "warmup": 1,
                   "scaling": "weak (2^%d amplitudes per GPU)" % args.shard_qubits,
                   "parallelism": "one state over %d GPUs (global-qubit sharding, fused NVLink push)" % world})
        line["extra_workloads"] = shard_extra
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(n, layers)
    if world == 1 and not args.no_extra and args.workload == "sv20":
        # the HBM-meaningful size of BASELINE.json configs[4] (vectors >> L2), same step definition
        def clocked(fn, **kw):
            """Runs one extra workload with its own clock / throttle-reason record."""
            smp = ClockSampler(local_rank)
            smp.start()
            try:
                out = fn(**kw)
            finally:
                clk = smp.stop()
            out["clocks"] = clk
            return out

        n2, l2 = WORKLOADS["sv28"]
        smp = ClockSampler(local_rank)
        r2 = measure_gpu(n2, l2, 3, 3, local_rank, False, smp)
        clk2 = smp.stop()
        p2 = pair_runs(n2, l2)
        t2 = float(np.mean(r2["step_ms"])) * 1e-3
        g2 = float(np.mean(r2["grad_ms"])) * 1e-3
        o2 = float(np.mean(r2["obj_ms"])) * 1e-3
        line["extra_workloads"] = {"sv28": {
            "num_qubits": n2, "layers": l2, "num_thetas": r2["circ"].num_thetas, "steps": 3, "warmup": 3,
            "value": 1.0 / t2, "unit": UNIT, "ms_per_step": t2 * 1e3,
            "e2e_value": 1.0 / float(np.mean(r2["e2e_s"])),
            "two_term_case": r2["two_term"],
            "kernel_ms": {"vh_apply_sweep": o2 * 1e3, "gradient_sweep": g2 * 1e3},
            "tile_passes": {"gradient": r2["passes_grad"], "vh_apply": r2["passes_dag"]},
            "stages": {"gradient": r2["stages_grad"], "vh_apply": r2["stages_dag"]},
            "roofline": roofline_block("sv28", n2, p2, r2["stages_grad"], r2["stages_dag"], g2, o2, t2,
                                       r2["passes_grad"]),
            "clocks": clk2,
        }}
        cpu = not args.no_cpu_baseline
        for key, fn, kw in (("mps50", measure_mps, dict(device=local_rank, with_cpu=cpu)),
                            ("mps50_physical", measure_mps, dict(device=local_rank, with_cpu=False, physical=True)),
                            ("mat7", measure_mat7, dict(device=local_rank, with_cpu=cpu)),
                            ("cd7", measure_cd7, dict(device=local_rank, with_cpu=cpu)),
                            ("sketch12", measure_sketch, dict(device=local_rank, with_cpu=cpu)),
                            ("lbfgs12", measure_lbfgs, dict(device=local_rank, with_cpu=cpu))):
            try:
                line["extra_workloads"][key] = clocked(fn, **kw)
            except Exception as ex:  # an extra must never break the headline line (nor the other extras)
                line["extra_workloads"][key] = {"error": repr(ex)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
