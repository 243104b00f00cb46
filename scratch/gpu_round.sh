#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list + full capture (tag = $1)
TAG=${1:-v1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks_$TAG.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2>&1
cat gpurun_out/bench_ref_$TAG.json
timeout 300 python scratch/bench_sort.py > gpurun_out/sort_$TAG.log 2>&1; cat gpurun_out/sort_$TAG.log
kill $SMI
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_bench_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -o gpurun_out/prof_$TAG -f python scratch/prof_run.py > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out
