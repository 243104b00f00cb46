#!/bin/bash
N=${1:-2}; IM=${2:-2000}
O=gpurun_out; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench_sweep.py --gpus $N --images $IM > $O/sweep_n${N}_$IM.json 2> $O/sweep_n${N}_$IM.err; grep '^{' $O/sweep_n${N}_$IM.json; tail -3 $O/sweep_n${N}_$IM.err
