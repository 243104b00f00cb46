#!/bin/bash
# usage: gpu_sweep_ab.sh N  -- cfg-4 sweep on N GPUs with and without the bulk-store exchange
N=$1
mkdir -p gpurun_out
for mode in bulk nobulk; do
  if [ $mode = nobulk ]; then export MSS_PARTITION_NO_BULK=1; else unset MSS_PARTITION_NO_BULK; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench_sweep.py --gpus $N --steps 3 --warmup 1 --no-oracle > gpurun_out/sweep_n${N}_$mode.json 2> gpurun_out/sweep_n${N}_$mode.err
  python - <<PY
import json
l=[x for x in open('gpurun_out/sweep_n${N}_$mode.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1])
print('$mode', 'images/s', round(d['value']), 'ms', round(d['ms_per_step'],2), d['phases_ms'], d.get('exchange',{}).get('phase_ms_rank0'), d['matches_single_pool'])
PY
done
