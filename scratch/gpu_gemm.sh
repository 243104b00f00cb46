#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_maskgemm.py tests/test_gpu_head.py -x -q 2>&1 | grep -v "^$" | tail -${1:-3}
timeout 200 python scratch/bench_maskgemm.py 2>&1 | head -1; timeout 200 python scratch/bench_head.py 2>&1 | head -1
