#!/bin/bash
# final 1-GPU evidence of the round: full GPU suite, headline bench, reference arm, cfg-5 sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_pytest_final.log; cat gpurun_out/r02_pytest_final.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1_final.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "reference arm rc=$?"; cut -c1-600 gpurun_out/r02_bench_reference_arm.json
timeout 900 python bench_sweep.py --cfg 5 > gpurun_out/r02_sweep_cfg5_n1.json 2> gpurun_out/r02_sweep_cfg5_n1.err; echo "cfg5 rc=$?"; cut -c1-1500 gpurun_out/r02_sweep_cfg5_n1.json; tail -3 gpurun_out/r02_sweep_cfg5_n1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads([x for x in open('gpurun_out/r02_bench_n1_final.json').read().splitlines() if x.startswith('{')][-1])
print({k:d[k] for k in ('metric','value','unit','ms_per_step','gpu_launches','clocks')})
print('roofline',d['roofline']); print('cpu_baseline',d['cpu_baseline']); print('e2e',d['e2e'])
x=d['extra']; print(list(x))
print('sweep',{k:x['eval_sweep'][k] for k in ('images_s','ms_per_step','phases_ms','pool_matches_oracle','bit_exact_vs_pool') if k in x['eval_sweep']})
print('metrics_stage.sort',json.dumps(x['metrics_stage']['sort'])[:1500])
print('by_size',json.dumps(x['metrics_stage']['by_size'])[:1500])
PY
