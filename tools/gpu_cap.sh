#!/bin/bash
# usage: gpu_cap.sh TAG IMGS name:regex:skip ...   -- one ncu --set full capture per named kernel of eval_ood_measure
TAG=$1; IMGS=$2; shift 2
O=gpurun_out; mkdir -p $O
for spec in "$@"; do
  IFS=: read name regex skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o $O/prof_${TAG}_$name python tools/prof_run.py eval $IMGS > $O/ncu_${name}_$TAG.log 2>&1
  ncu -i $O/prof_${TAG}_$name.ncu-rep --page raw --csv > $O/prof_${TAG}_${name}_raw.csv 2>/dev/null
  ncu -i $O/prof_${TAG}_$name.ncu-rep --page source --csv > $O/prof_${TAG}_${name}_source.csv 2>/dev/null
  rm -f $O/prof_${TAG}_$name.ncu-rep
  python tools/ncu_summary.py raw $O/prof_${TAG}_${name}_raw.csv | head -14
  python tools/ncu_src.py $O/prof_${TAG}_${name}_source.csv 16
done
