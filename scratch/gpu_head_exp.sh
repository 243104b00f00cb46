#!/bin/bash
# sweep -D macro sets of the head kernel on the box: each argument is a quoted flag string
O=gpurun_out; mkdir -p $O
for F in "$@"; do
  MSS_NVCC_EXTRA="$F" python -m multishiftseg_b200.build --force > $O/build_head.log 2>&1 || { tail -5 $O/build_head.log; continue; }
  echo "== $F"; timeout 300 python -m pytest tests/test_gpu_head.py -m gpu -x -q 2>&1 | tail -1; timeout 300 python scratch/bench_head.py 2>&1 | head -1
done
python -m multishiftseg_b200.build --force > /dev/null 2>&1
