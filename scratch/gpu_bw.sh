#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_backward.py tests/test_gpu_scoring.py -x -q 2>&1 | grep -v "^$" | tail -${1:-4}
timeout 300 python scratch/bench_backward.py 2>&1 | tail -5
