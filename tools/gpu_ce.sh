#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_n${N}_ce.log 2>&1; echo "dist_check rc=$?"; grep -c "True" gpurun_out/dist_check_n${N}_ce.log; grep "False\|rror" gpurun_out/dist_check_n${N}_ce.log | cut -c1-200 | tail
./tools/gpu_sweep.sh $N ce --exchange stream 2>&1 | grep -v "^\*\*\*\|OMP_NUM"
