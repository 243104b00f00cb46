#!/bin/bash
# cfg-4 sweep at whatever GPU count the box has ($1 = N, $2 = images)
N=${1:-1}; IM=${2:-2000}
O=gpurun_out; mkdir -p $O
if [ "$N" = "1" ]; then
  timeout 900 python bench_sweep.py --images 64 --steps 2 > $O/sweep_n1_64.json 2> $O/sweep_n1_64.err; cat $O/sweep_n1_64.json; tail -3 $O/sweep_n1_64.err
  timeout 900 python bench_sweep.py --images $IM > $O/sweep_n1_$IM.json 2> $O/sweep_n1_$IM.err; cat $O/sweep_n1_$IM.json; tail -3 $O/sweep_n1_$IM.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench_sweep.py --gpus $N --images $IM > $O/sweep_n${N}_$IM.json 2> $O/sweep_n${N}_$IM.err; cat $O/sweep_n${N}_$IM.json; tail -5 $O/sweep_n${N}_$IM.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 > $O/bench_n${N}.json 2> $O/bench_n${N}.err; cat $O/bench_n${N}.json; tail -3 $O/bench_n${N}.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/dist_gpu_check.py > $O/dist_check_n${N}.log 2>&1; tail -5 $O/dist_check_n${N}.log
fi
