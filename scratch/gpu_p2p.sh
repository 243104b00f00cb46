#!/bin/bash
N=${1:-2}; IM=${2:-2000}
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_metrics.py -m gpu -x -q -k "partition or histogram" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/dist_gpu_check.py > $O/dist_check_n${N}.log 2>&1; grep "world\|Error\|error" $O/dist_check_n${N}.log | tail -20
bash scratch/gpu_sweep2.sh $N $IM
