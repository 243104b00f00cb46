#!/bin/bash
mkdir -p gpurun_out
python tools/exch_local.py 2 2>&1 | tail -1
python tools/exch_local.py 4 2>&1 | tail -1
python tools/exch_local.py 8 2>&1 | tail -1
python tools/exch_local.py 16 2>&1 | tail -1
timeout 600 python -m pytest tests -x -q -m gpu -k "exchange or partition or evaluator" 2>&1 | tail -3
