#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_evaluator.py tests/test_gpu_hardening.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python tools/bench_sort.py 2097152 8388608 33554432 134217728 2>&1 | tee gpurun_out/r02_sort_c.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_eval_c.csv python tools/prof_run.py eval 64 > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_eval_c.csv | grep mss | tee gpurun_out/launches_eval_c_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_eval1_c.csv python tools/prof_run.py eval 1 > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_eval1_c.csv | grep mss | tee gpurun_out/launches_eval1_c_summary.txt
