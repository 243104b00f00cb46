#!/bin/bash
# usage: gpu_n.sh N TAG   -- multi-GPU validation + bench on N GPUs of one box
N=$1; TAG=$2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_n${N}_$TAG.log 2>&1; echo "dist_check rc=$?"; grep -v "^W\|^\[W\|warn" gpurun_out/dist_check_n${N}_$TAG.log | tail -16
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tests/dist_c_abi_check.py > gpurun_out/dist_c_abi_n${N}_$TAG.log 2>&1; echo "c_abi rc=$?"; grep "C ABI\|rror" gpurun_out/dist_c_abi_n${N}_$TAG.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n${N}_$TAG.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_n${N}_$TAG.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'h2d_ceiling',d['e2e'].get('h2d_ceiling_GBs_per_gpu'),'copy_only',d['e2e'].get('copy_only_ceiling_value'))
print(json.dumps(d['extra']['eval_sweep'],indent=1)[:3500])
PY
