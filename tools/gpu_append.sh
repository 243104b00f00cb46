#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "append or evaluator or eval or metric or hardening" 2>&1 | tail -4
timeout 600 python tools/bench_sort.py 2097152 33554432 134217728 2>&1 | tee gpurun_out/append_$1.log
