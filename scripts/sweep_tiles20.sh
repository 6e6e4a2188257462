#!/bin/bash
# development aid: tile-shape sweep at n=20 / 24
run() {
  python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); print('$W $1', round(d['value'], 3), 'evals/s', d['kernel_ms'], d['config']['tile_passes'], 'e2e', round(d['e2e']['value'],1))
    else:
        print(line, end='')
"
}
for W in sv20 sv24; do
run "default"
AQC_TILE_BITS_APPLY=10 AQC_TILE_BITS_GRAD=10 run "tb10 low=4"
AQC_TILE_BITS_APPLY=10 AQC_TILE_BITS_GRAD=10 AQC_TILE_LOW_BITS=3 run "tb10 low=3"
AQC_TILE_BITS_APPLY=10 AQC_TILE_BITS_GRAD=10 AQC_TILE_LOW_BITS=2 run "tb10 low=2"
AQC_TILE_LOW_BITS=3 run "tb11 low=3"
AQC_TILE_LOW_BITS=2 run "tb11 low=2"
AQC_TILE_BITS_APPLY=12 run "apply12 low=4"
AQC_TILE_BITS_APPLY=9 AQC_TILE_BITS_GRAD=9 AQC_TILE_LOW_BITS=2 run "tb9 low=2"
done
