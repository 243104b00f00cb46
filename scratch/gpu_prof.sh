#!/bin/bash
# ncu evidence only: launch list + one full capture per hot kernel (1 launch each, small reports)
TAG=${1:-v1}
O=gpurun_out
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv python scratch/prof_run2.py all 16 > $O/ncu_list_$TAG.log 2>&1
cap() {  # name regex part skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -f -o $O/prof_${TAG}_$1 python scratch/prof_run2.py $3 16 > $O/ncu_$1_$TAG.log 2>&1
  ncu -i $O/prof_${TAG}_$1.ncu-rep --page raw --csv > $O/prof_${TAG}_$1_raw.csv 2>/dev/null
  ncu -i $O/prof_${TAG}_$1.ncu-rep --page source --csv > $O/prof_${TAG}_$1_source.csv 2>/dev/null
}
cap score deeplab_score score 1
cap sweep onesweep_pass eval 5
cap hist radix_histogram eval 1
cap runs 'runs_kernel' eval 3
cap roc 'roc_compact' eval 3
cap leaf 'leaf_sum' eval 2
cap append 'eval_append' eval 1
cap m2f 'm2f_.*x4' m2f 2
du -sh $O; ls -la $O
# keep the merge-back under 64 MiB: drop the biggest reports if needed
while [ $(du -sm $O | cut -f1) -gt 60 ]; do f=$(ls -S $O/*.ncu-rep | head -1); echo "dropping $f"; rm -f $f; done
