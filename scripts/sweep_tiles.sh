#!/bin/bash
# development aid: tile-shape sweep at n=28
run() {
  python bench.py --workload sv28 --steps 2 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); print('$1', round(d['value'], 3), 'evals/s', d['kernel_ms'], d['config']['tile_passes'])
    else:
        print(line, end='')
"
}
for low in 3 5 6; do AQC_TILE_LOW_BITS=$low run "low=$low"; done
AQC_TILE_BITS_APPLY=12 run "apply12 low=4"
AQC_TILE_BITS_APPLY=12 AQC_TILE_LOW_BITS=5 run "apply12 low=5"
AQC_TILE_BITS_APPLY=10 AQC_TILE_BITS_GRAD=10 run "tb10 low=4"
