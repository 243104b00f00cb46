#!/bin/bash
# timing experiments on the shared-tap M2F kernel: rebuild on the box with -DTQ_EXPERIMENT=n (results are wrong on purpose)
O=gpurun_out; mkdir -p $O
for E in ${@:-0 1 2 3}; do
  MSS_NVCC_EXTRA="-DTQ_EXPERIMENT=$E" python -m multishiftseg_b200.build --force > $O/build_exp$E.log 2>&1 || { tail -5 $O/build_exp$E.log; continue; }
  echo "== TQ_EXPERIMENT=$E"; timeout 300 python scratch/bench_m2f.py 0 2>&1 | tail -2
done
python -m multishiftseg_b200.build --force > /dev/null 2>&1
