#!/bin/bash
N=${1:-8}
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench_sweep.py --gpus $N --images 2000 > $O/sweep_n${N}_v5.json 2> $O/sweep_n${N}_v5.err; grep '^{' $O/sweep_n${N}_v5.json; tail -2 $O/sweep_n${N}_v5.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench_sweep.py --gpus $N --cfg 5 > $O/sweep5_n${N}_v5.json 2> $O/sweep5_n${N}_v5.err; grep '^{' $O/sweep5_n${N}_v5.json; tail -2 $O/sweep5_n${N}_v5.err
