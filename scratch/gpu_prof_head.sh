#!/bin/bash
TAG=${1:-ph}
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_gemm -s 2 -c 1 -f -o $O/prof_${TAG}_head python scratch/bench_head.py > $O/ncu_head_$TAG.log 2>&1
ncu -i $O/prof_${TAG}_head.ncu-rep --page raw --csv > $O/prof_${TAG}_head_raw.csv 2>/dev/null
ncu -i $O/prof_${TAG}_head.ncu-rep --page source --csv > $O/prof_${TAG}_head_source.csv 2>/dev/null
python scratch/ncu_summary.py raw $O/prof_${TAG}_head_raw.csv
rm -f $O/prof_${TAG}_head.ncu-rep
