#!/bin/bash
# ncu evidence for the metric stage: launch list of eval_ood_measure at IMGS images + full captures of its hot kernels
TAG=${1:-r02a}; IMGS=${2:-64}
O=gpurun_out
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_eval_$TAG.csv python tools/prof_run.py eval $IMGS > $O/ncu_list_eval_$TAG.log 2>&1
python tools/ncu_summary.py launches $O/launches_eval_$TAG.csv | tee $O/launches_eval_${TAG}_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_eval1_$TAG.csv python tools/prof_run.py eval 1 > $O/ncu_list_eval1_$TAG.log 2>&1
python tools/ncu_summary.py launches $O/launches_eval1_$TAG.csv | tee $O/launches_eval1_${TAG}_summary.txt
cap() {  # name regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/prof_${TAG}_$1 python tools/prof_run.py eval $IMGS > $O/ncu_$1_$TAG.log 2>&1
  ncu -i $O/prof_${TAG}_$1.ncu-rep --page raw --csv > $O/prof_${TAG}_$1_raw.csv 2>/dev/null
  ncu -i $O/prof_${TAG}_$1.ncu-rep --page source --csv > $O/prof_${TAG}_$1_source.csv 2>/dev/null
  rm -f $O/prof_${TAG}_$1.ncu-rep
}
cap sweep onesweep_pass 5
cap merge merge_counts 1
cap roc roc_compact 1
cap leaf leaf_sum 1
cap hist radix_histogram 1
python tools/ncu_summary.py raw $O/prof_${TAG}_*_raw.csv > $O/ncu_full_${TAG}_summary.txt
cat $O/ncu_full_${TAG}_summary.txt | head -150
