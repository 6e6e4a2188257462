#!/bin/bash
# development aid: three state-vector workloads, device-timed sweeps only
for w in sv20 sv24 sv28; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); print('$w', round(d['value'], 3), 'evals/s', d['kernel_ms'], 'e2e', round(d['e2e']['value'], 3))
    else:
        print(line, end='')
"
done
