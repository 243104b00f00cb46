#!/bin/bash
N=${1:-8}
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/topo_n$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/dist_gpu_check.py > $O/dist_check_n${N}.log 2>&1; grep "world" $O/dist_check_n${N}.log | tail -16
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench_sweep.py --gpus $N --images 2000 > $O/sweep_n${N}_2000.json 2> $O/sweep_n${N}_2000.err; grep '^{' $O/sweep_n${N}_2000.json; tail -3 $O/sweep_n${N}_2000.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 > $O/bench_n${N}.json 2> $O/bench_n${N}.err; grep '^{' $O/bench_n${N}.json; tail -3 $O/bench_n${N}.err
