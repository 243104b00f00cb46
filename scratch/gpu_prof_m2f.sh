#!/bin/bash
TAG=${1:-pm}
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:m2f_tc5q -s 2 -c 1 -f -o $O/prof_${TAG}_m2f python scratch/prof_run2.py m2f 16 > $O/ncu_m2f_$TAG.log 2>&1
ncu -i $O/prof_${TAG}_m2f.ncu-rep --page raw --csv > $O/prof_${TAG}_m2f_raw.csv 2>/dev/null
ncu -i $O/prof_${TAG}_m2f.ncu-rep --page source --csv > $O/prof_${TAG}_m2f_source.csv 2>/dev/null
python scratch/ncu_summary.py raw $O/prof_${TAG}_m2f_raw.csv
rm -f $O/prof_${TAG}_m2f.ncu-rep
