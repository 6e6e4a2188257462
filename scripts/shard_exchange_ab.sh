#!/bin/bash
# development aid: sharded bench with both exchange implementations.  usage: shard_exchange_ab.sh NGPUS SHARD_QUBITS
N=${1:-4}; Q=${2:-30}
for m in kernel push; do
  AQC_EXCHANGE=$m timeout 400 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N \
    --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --workload svshard --shard-qubits $Q --steps 2 --warmup 1 \
    2>/dev/null | grep '^{' > gpurun_out/shard_ab_$m.json
  python - "$m" <<'PY'
import json, sys
m = sys.argv[1]
for l in open(f"gpurun_out/shard_ab_{m}.json"):
    d = json.loads(l)
    print(m, d["value"], d["config"]["num_qubits"], d["kernel_ms"])
PY
done
