#!/bin/bash
# compute-sanitizer memcheck over the tests that exercise the kernels written or rewritten in round 2
O=gpurun_out; mkdir -p $O
export MSS_SANITIZE=1
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest -x -q -m gpu \
  tests/test_gpu_metrics.py tests/test_gpu_evaluator.py tests/test_gpu_backward.py tests/test_gpu_segmetric.py \
  -k "not 16_777_216 and not 5_000_011 and not histogram_counter_fold and not idempotent and not multi_gpu" > $O/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -12 $O/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 77 --print-limit 20 python -m pytest -x -q -m gpu \
  tests/test_gpu_metrics.py -k "kats or counts_and_tail or sort_two_segments or partition_scatter or two_stream" > $O/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -12 $O/r02_sanitizer_racecheck.log
