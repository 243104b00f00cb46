#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/r02_pytest_b.log
cat gpurun_out/r02_pytest_b.log
timeout 300 python tools/bench_sort.py 2097152 8388608 33554432 134217728 2>&1 | tee gpurun_out/r02_sort_b.txt
timeout 120 python tools/bench_sort.py f16 33554432 2>&1 | tee -a gpurun_out/r02_sort_b.txt
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -3
