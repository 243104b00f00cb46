#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_pytest_c.log; cat gpurun_out/r02_pytest_c.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err; tail -5 gpurun_out/r02_bench_n1_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_a.json').read().strip().splitlines()[-1])
def show(x,ind=0):
    for k,v in x.items():
        if isinstance(v,dict): print(' '*ind+k+':'); show(v,ind+2)
        else: print(' '*ind+f"{k}: {str(v)[:160]}")
show(d)
PY
