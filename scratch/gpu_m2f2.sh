#!/bin/bash
TAG=${1:-m}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_m2f.py -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_$TAG.log
timeout 300 python scratch/bench_m2f.py 0 8 > $O/m2f_$TAG.log 2>&1; cat $O/m2f_$TAG.log
