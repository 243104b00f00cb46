#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_metrics.py -x -q -m gpu 2>&1 | tail -3
echo "== persistent"; timeout 300 python tools/bench_sort.py 2097152 8388608 33554432 134217728 536870912 2>&1 | cut -c1-110
echo "== one tile per CTA"; MSS_SORT_ONE_TILE_CTAS=1 timeout 300 python tools/bench_sort.py 2097152 8388608 33554432 134217728 536870912 2>&1 | cut -c1-110
