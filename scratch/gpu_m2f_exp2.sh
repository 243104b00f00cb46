#!/bin/bash
# sweep a -D macro of the shared-tap M2F kernel on the box: gpu_m2f_exp2.sh MACRO v1 v2 ...
O=gpurun_out; mkdir -p $O
M=$1; shift
for E in "$@"; do
  MSS_NVCC_EXTRA="-D$M=$E" python -m multishiftseg_b200.build --force > $O/build_$M$E.log 2>&1 || { tail -5 $O/build_$M$E.log; continue; }
  echo "== $M=$E"; timeout 300 python -m pytest tests/test_gpu_m2f.py -m gpu -x -q -k "tma_tcgen05 and not pixel" 2>&1 | tail -1; timeout 300 python scratch/bench_m2f.py 0 2>&1 | tail -1
done
python -m multishiftseg_b200.build --force > /dev/null 2>&1
