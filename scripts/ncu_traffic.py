#!/usr/bin/env python
"""
Condenses an `ncu --set full` capture into profiles/ncu_traffic.json (read by bench.py's roofline block).
usage: ncu_traffic.py <report.ncu-rep> <workload> [kernel-substring]
Writes, for the workload: mean dram__bytes_read.sum + dram__bytes_write.sum per captured launch of the
kernel, the launch count, mean duration, DMMA-pipe and DRAM utilisation, and the git hash of the tree the
capture was taken from (the capture is made from a clean tree right after the commit).
"""
import csv, io, json, os, subprocess, sys

rep, workload = sys.argv[1], sys.argv[2]
needle = sys.argv[3] if len(sys.argv) > 3 else "dense_pass_kernel<2>"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(row, name, unit_scale=None):
    i = col[name]
    x = float(row[i].replace(",", ""))
    u = units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9,
             "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
    return x * scale


picked = [r for r in data if needle in r[col["Kernel Name"]]]
if not picked:
    sys.exit(f"no launch of {needle} in {rep}")
n = len(picked)
rec = {
    "kernel": needle,
    "launches_captured": n,
    "dram_bytes_per_launch": sum(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in picked) / n,
    "mean_launch_s_under_ncu": sum(val(r, "gpu__time_duration.sum") for r in picked) / n,
    "dmma_pipe_pct": sum(float(r[col["sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"]]) for r in picked) / n
    if "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active" in col else None,
    "dram_throughput_pct": sum(float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]) for r in picked) / n,
    "capture": os.path.basename(rep),
    "git": sys.argv[4] if len(sys.argv) > 4 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip(),
}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, "profiles", "ncu_traffic.json")
try:
    allrec = json.load(open(path))
except (OSError, ValueError):
    allrec = {}
allrec[workload] = rec
json.dump(allrec, open(path, "w"), indent=1)
print(json.dumps(rec, indent=1))
