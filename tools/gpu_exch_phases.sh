#!/bin/bash
N=$1; shift
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/exch_phases.py "$@" 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^W\|^$" | tee gpurun_out/exch_phases_n$N.txt
