#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_evaluator.py tests/test_gpu_metrics.py -x -q 2>&1 | tail -2
timeout 300 python bench_sweep.py --steps 3 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['phases_ms'], d['result_hex'], d['matches_single_pool'])"
