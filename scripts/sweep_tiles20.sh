#!/bin/bash
# development aid: tile-shape sweep at n=20 / 22
run() {
  python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); print('$W $1', round(d['value'], 1), 'evals/s', {k: round(v, 4) for k, v in d['kernel_ms'].items()}, d['config']['tile_passes'], 'e2e', round(d['e2e']['value'],1))
    else:
        print(line, end='')
"
}
for W in sv20; do
run "default"
AQC_TILE_LOW_BITS=3 run "tb11 low=3"
AQC_TILE_LOW_BITS=2 run "tb11 low=2"
AQC_TILE_LOW_BITS=1 run "tb11 low=1"
AQC_TILE_LOW_BITS=0 run "tb11 low=0"
AQC_TILE_BITS_APPLY=12 AQC_TILE_LOW_BITS=2 run "apply12 low=2"
AQC_TILE_BITS_APPLY=10 AQC_TILE_BITS_GRAD=10 AQC_TILE_LOW_BITS=2 run "tb10 low=2"
done
