#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_n${N}_s.log 2>&1; echo "dist_check rc=$?"; grep "stream\|rror\|False" gpurun_out/dist_check_n${N}_s.log | tail -12
./tools/gpu_sweep.sh $N auto
./tools/gpu_sweep.sh $N stream --exchange stream
