#!/bin/bash
# usage: gpu_t.sh "<pytest -k expression>" [extra command]
timeout 900 python -m pytest tests -x -q -m gpu -k "$1" 2>&1 | tail -15
shift
if [ -n "$1" ]; then bash -c "$*"; fi
