#!/bin/bash
TAG=${1:-pm}
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_gemm_kernel -s 2 -c 1 -f -o $O/prof_${TAG}_maskgemm python scratch/bench_maskgemm.py > $O/ncu_maskgemm_$TAG.log 2>&1
ncu -i $O/prof_${TAG}_maskgemm.ncu-rep --page raw --csv > $O/prof_${TAG}_maskgemm_raw.csv 2>/dev/null
ncu -i $O/prof_${TAG}_maskgemm.ncu-rep --page source --csv > $O/prof_${TAG}_maskgemm_source.csv 2>/dev/null
ncu -i $O/prof_${TAG}_maskgemm.ncu-rep --page details > $O/prof_${TAG}_maskgemm_details.txt 2>/dev/null
python scratch/ncu_summary.py raw $O/prof_${TAG}_maskgemm_raw.csv
