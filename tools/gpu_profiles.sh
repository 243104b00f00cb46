#!/bin/bash
# Round-2 evidence on one GPU: tests, launch lists (bench command + metric stage), one --set full capture per hot kernel.
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_pytest_gpu.log
# launch list of the bench command itself (numbers printed under ncu are not bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --sweep-images 64 --continuous-frames 8 > $O/${TAG}_ncu_bench.log 2>&1
python tools/ncu_summary.py launches $O/${TAG}_launches_bench.csv > $O/${TAG}_launches_bench_summary.txt; head -30 $O/${TAG}_launches_bench_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${TAG}_launches_eval64.csv python tools/prof_run.py eval 64 > /dev/null 2>&1
python tools/ncu_summary.py launches $O/${TAG}_launches_eval64.csv > $O/${TAG}_launches_eval64_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${TAG}_launches_eval1.csv python tools/prof_run.py eval 1 > /dev/null 2>&1
python tools/ncu_summary.py launches $O/${TAG}_launches_eval1.csv > $O/${TAG}_launches_eval1_summary.txt
cap() {  # name regex part skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -f -o $O/prof_${TAG}_$1 python tools/prof_run.py $3 64 > $O/ncu_$1_$TAG.log 2>&1
  ncu -i $O/prof_${TAG}_$1.ncu-rep --page raw --csv > $O/prof_${TAG}_$1_raw.csv 2>/dev/null
  ncu -i $O/prof_${TAG}_$1.ncu-rep --page source --csv > $O/prof_${TAG}_$1_source.csv 2>/dev/null
  rm -f $O/prof_${TAG}_$1.ncu-rep
}
cap score deeplab_score score 1
cap sweep onesweep_pass eval 5
cap hist radix_histogram eval 1
cap merge merge_counts eval 1
cap roc roc_compact eval 1
cap leaf leaf_sum eval 1
cap append eval_append eval 1
cap m2f 'm2f_tc5q' m2f 1
cap head 'pixel_gemm.*HeadEpi' gemm 1
cap maskgemm 'pixel_gemm.*MaskEpi' gemm 1
python tools/ncu_summary.py raw $O/prof_${TAG}_*_raw.csv > $O/${TAG}_ncu_full_summary.txt
for k in sweep merge roc leaf hist append; do echo "== $k"; python tools/ncu_src.py $O/prof_${TAG}_${k}_source.csv 6; done > $O/${TAG}_ncu_stalls.txt 2>&1
grep -A3 "^==" $O/${TAG}_ncu_full_summary.txt | grep "==\|time_duration\|dram__bytes" | head -60
