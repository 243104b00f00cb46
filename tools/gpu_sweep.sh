#!/bin/bash
# usage: gpu_sweep.sh N TAG [extra bench_sweep args]
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench_sweep.py --gpus $N --steps 3 --warmup 1 --no-oracle "$@" > gpurun_out/sweep_n${N}_$TAG.json 2> gpurun_out/sweep_n${N}_$TAG.err
python - <<PY
import json
l=[x for x in open('gpurun_out/sweep_n${N}_$TAG.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1])
print('$TAG', 'images/s', round(d['value']), 'ms', round(d['ms_per_step'],2), d['phases_ms'], d.get('exchange',{}).get('phase_ms_rank0'), d['matches_single_pool'])
if 'continuous' in d: print(d['continuous'])
PY
tail -3 gpurun_out/sweep_n${N}_$TAG.err
