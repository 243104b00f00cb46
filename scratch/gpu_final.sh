#!/bin/bash
# Evidence visit (tag = $1): parity suite, smoke, both bench arms, sort sweep, cfg-4/5 sweeps, ncu launch lists, full captures
TAG=${1:-v4}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; tail -1 $O/smoke_$TAG.log
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; cat $O/bench_$TAG.json; tail -3 $O/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref_$TAG.json 2>&1; cat $O/bench_ref_$TAG.json
timeout 300 python scratch/bench_sort.py > $O/sort_$TAG.log 2>&1; cat $O/sort_$TAG.log
timeout 300 python scratch/bench_m2f.py 0 8 4 2 > $O/m2f_$TAG.log 2>&1; cat $O/m2f_$TAG.log
timeout 200 python scratch/bench_maskgemm.py > $O/maskgemm_$TAG.log 2>&1; cat $O/maskgemm_$TAG.log
timeout 200 python scratch/bench_head.py > $O/head_$TAG.log 2>&1; cat $O/head_$TAG.log
timeout 200 python scratch/bench_backward.py > $O/backward_$TAG.log 2>&1; cat $O/backward_$TAG.log
timeout 600 python bench_sweep.py > $O/sweep_$TAG.json 2> $O/sweep_$TAG.err; cat $O/sweep_$TAG.json
timeout 600 python bench_sweep.py --cfg 5 > $O/sweep5_$TAG.json 2> $O/sweep5_$TAG.err; cat $O/sweep5_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu_bench_$TAG.log 2>&1
python scratch/ncu_summary.py launches $O/launches_bench_$TAG.csv | head -30
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv python scratch/prof_run2.py all 16 > $O/ncu_list_$TAG.log 2>&1
python scratch/ncu_summary.py launches $O/launches_$TAG.csv | head -40
cap() {  # name regex part skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -f -o $O/prof_${TAG}_$1 python scratch/prof_run2.py $3 16 > $O/ncu_$1_$TAG.log 2>&1
  ncu -i $O/prof_${TAG}_$1.ncu-rep --page raw --csv > $O/prof_${TAG}_$1_raw.csv 2>/dev/null
  ncu -i $O/prof_${TAG}_$1.ncu-rep --page source --csv > $O/prof_${TAG}_$1_source.csv 2>/dev/null
}
cap score deeplab_score score 1
cap m2f 'm2f_tc5q' m2f 2
cap sweep onesweep_pass eval 5
cap hist radix_histogram eval 1
cap head pixel_gemm_kernel gemm 1
cap maskgemm pixel_gemm_kernel gemm 3
python scratch/make_traffic.py $O/prof_${TAG}_score_raw.csv $TAG
python scratch/ncu_summary.py raw $O/prof_${TAG}_*_raw.csv > $O/ncu_full_${TAG}_summary.txt 2>&1
rm -f $O/prof_${TAG}_*.ncu-rep
du -sh $O
