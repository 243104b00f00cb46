#!/bin/bash
# CUB yardstick + the repo's sort at the same sizes, one box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r02_yard_smi.txt
./tools/cub_yardstick 2097152 8388608 33554432 134217728 536870912 > gpurun_out/r02_cub_yardstick.jsonl 2>&1
python scratch/bench_sort.py 2097152 8388608 33554432 134217728 > gpurun_out/r02_sort_before.txt 2>&1
cat gpurun_out/r02_cub_yardstick.jsonl gpurun_out/r02_sort_before.txt
