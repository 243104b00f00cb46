#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_head.py -m gpu -x -q > $O/pytest_head.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_head.log
timeout 300 python scratch/bench_head.py 2>&1 | tail -3
