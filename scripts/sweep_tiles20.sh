#!/bin/bash
# development aid: tile-shape sweep at n=20
run() {
  python bench.py --workload sv20 --steps 20 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); print('$1', round(d['value'], 1), 'evals/s', {k: round(v, 4) for k, v in d['kernel_ms'].items()}, d['config']['tile_passes'], d['config']['stages'], 'e2e', round(d['e2e']['value'],1))
    else:
        print(line, end='')
"
}
run "default"
AQC_DENSE_PAIRS=0 run "pairs=0"
AQC_DENSE_PAIRS=2 run "pairs=2"
AQC_TILE_BITS_GRAD=11 run "grad tb11 low=2"
AQC_TILE_BITS_GRAD=11 AQC_TILE_LOW_BITS=1 run "grad tb11 low=1"
AQC_TILE_LOW_BITS=1 run "grad tb10 low=1"
AQC_TILE_LOW_BITS=3 run "grad tb10 low=3"
AQC_TILE_BITS_APPLY=12 run "apply tb12 low=1"
AQC_TILE_BITS_APPLY=10 run "apply tb10 low=1"
