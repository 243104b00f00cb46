#!/bin/bash
TAG=${1:-v4}
O=gpurun_out; mkdir -p $O
cap() {  # name regex part skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -f -o $O/prof_${TAG}_$1 python scratch/prof_run2.py $3 16 > $O/ncu_$1_$TAG.log 2>&1
  ncu -i $O/prof_${TAG}_$1.ncu-rep --page raw --csv > $O/prof_${TAG}_$1_raw.csv 2>/dev/null
  ncu -i $O/prof_${TAG}_$1.ncu-rep --page source --csv > $O/prof_${TAG}_$1_source.csv 2>/dev/null
}
cap head pixel_gemm_kernel gemm 1
cap maskgemm pixel_gemm_kernel gemm 3
python scratch/ncu_summary.py raw $O/prof_${TAG}_head_raw.csv $O/prof_${TAG}_maskgemm_raw.csv | tee $O/ncu_full_${TAG}_gemm_summary.txt
rm -f $O/prof_${TAG}_*.ncu-rep
