#!/bin/bash
# development aid: tile-shape sweep at n=24/28
run() {
  python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); print('$W $1', round(d['value'], 3), 'evals/s', {k: round(v, 3) for k, v in d['kernel_ms'].items()}, d['config']['tile_passes'])
    else:
        print(line, end='')
"
}
for W in sv24 sv28; do
run "default"
AQC_TILE_BITS_GRAD=10 run "grad tb10 low=4"
AQC_TILE_BITS_GRAD=10 AQC_TILE_LOW_BITS=3 run "grad tb10 low=3"
AQC_TILE_LOW_BITS=3 run "tb11 low=3"
AQC_TILE_BITS_APPLY=12 run "apply tb12"
AQC_TILE_BITS_APPLY=12 AQC_TILE_LOW_BITS_APPLY=3 run "apply tb12 low=3"
done
