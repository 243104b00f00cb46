#!/bin/bash
# quick GPU visit: parity tests, sort/eval sweep, launch list of the eval path
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_$TAG.log
timeout 300 python scratch/bench_sort.py > $O/sort_$TAG.log 2>&1; cat $O/sort_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv python scratch/prof_run2.py ${2:-eval} 16 > $O/ncu_list_$TAG.log 2>&1
python scratch/ncu_summary.py launches $O/launches_$TAG.csv
