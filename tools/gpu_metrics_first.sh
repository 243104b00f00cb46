#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_evaluator.py tests/test_gpu_scoring.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r02_pytest_a.log
cat gpurun_out/r02_pytest_a.log
timeout 300 python tools/bench_sort.py 2097152 8388608 33554432 134217728 2>&1 | tee gpurun_out/r02_sort_a.txt
timeout 120 python tools/bench_sort.py f16 33554432 2>&1 | tee -a gpurun_out/r02_sort_a.txt
