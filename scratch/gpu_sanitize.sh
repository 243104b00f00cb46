#!/bin/bash
# compute-sanitizer memcheck over the kernels added in the last visits (small shapes only)
O=gpurun_out; mkdir -p $O
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 \
  python -m pytest tests/test_gpu_backward.py tests/test_gpu_maskgemm.py tests/test_gpu_head.py -x -q \
  -k "not 128-256 and not cfg5 and not model_shape and not many_images and not 150-32" > $O/sanitize_v5.log 2>&1
echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" $O/sanitize_v5.log | head -12
