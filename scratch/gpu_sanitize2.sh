#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 \
  python -m pytest tests/test_gpu_metrics.py -x -q -k "kats or sort_pairs_exact or partition_many or partition_scatter or one_shot or label_dtypes" > $O/sanitize_sort_v5.log 2>&1
echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" $O/sanitize_sort_v5.log | head -12
