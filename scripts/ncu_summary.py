#!/usr/bin/env python
"""Development aid: summarises an .ncu-rep (raw metrics + hottest SASS lines).  usage: ncu_summary.py rep [launch]"""
import csv, subprocess, sys, io

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w} [{units[i]}]", [d[i] for d in data])
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and 'per_issue_active' in h:
        vals = [round(float(d[i]), 2) for d in data]
        if max(vals) >= 0.3:
            print("stall", h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), vals)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'hdr': None, 'rows': []}
        blocks.append(cur)
        continue
    if cur is None:
        continue
    if cur['hdr'] is None:
        cur['hdr'] = r
        continue
    cur['rows'].append(r)
b = blocks[which]
h = b['hdr']
iS, iN, iE = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
tot = sum(int(r[iN]) for r in b['rows'] if len(r) > iN)
print('total samples', tot, 'n instr', len(b['rows']))
stalls = [x for x in h if x.startswith('stall_') and 'Not Issued' not in x]
for k, r in enumerate(b['rows']):
    n = int(r[iN])
    if n > tot * 0.01:
        top = sorted(((int(r[h.index(s)]), s) for s in stalls), reverse=True)[:2]
        print(k, r[iS].strip()[:60].ljust(60), n, f"{100*n/tot:.1f}%", r[iE], top)
