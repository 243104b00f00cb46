#!/bin/bash
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_eval_d.csv python tools/prof_run.py eval 64 > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_eval_d.csv | grep mss
