#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_evaluator.py tests/test_gpu_scoring.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_sort.py 2097152 8388608 33554432 134217728 536870912 2>&1 | cut -c1-150
