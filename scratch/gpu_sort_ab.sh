#!/bin/bash
echo "== SORT_CTAS_PER_SM=4 (built)"; timeout 300 python scratch/bench_sort.py 8388608 134217728 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_metrics.py -x -q 2>&1 | tail -2
touch multishiftseg_b200/csrc/radix_sort.cu
MSS_NVCC_EXTRA="-DSORT_CTAS_PER_SM=3" python -m multishiftseg_b200.build 2>&1 | tail -1
echo "== SORT_CTAS_PER_SM=3"; timeout 300 python scratch/bench_sort.py 8388608 134217728 2>&1 | tail -2
