#!/bin/bash
# quick GPU visit: parity tests (optionally a -k filter in $2), sort/eval sweep
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_$TAG.log
timeout 300 python scratch/bench_sort.py > $O/sort_$TAG.log 2>&1; cat $O/sort_$TAG.log
