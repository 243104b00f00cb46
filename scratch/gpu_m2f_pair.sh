#!/bin/bash
echo "== default"; timeout 300 python scratch/bench_m2f.py 0 2>&1 | tail -1
touch multishiftseg_b200/csrc/m2f_semantic.cu
MSS_NVCC_EXTRA="-DTQ_RCP_PAIR=1" python -m multishiftseg_b200.build 2>&1 | tail -1
echo "== TQ_RCP_PAIR=1"; timeout 300 python scratch/bench_m2f.py 0 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_m2f.py -x -q 2>&1 | tail -3
