#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_evaluator.py -x -q 2>&1 | tail -2
timeout 300 python scratch/bench_sort.py 2>&1 | tail -4
