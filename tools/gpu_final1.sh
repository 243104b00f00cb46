#!/bin/bash
# final 1-GPU evidence of the round: full GPU suite, headline bench, ncu of the two kernels that changed last, smoke
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > $O/r02_pytest_final.log; cat $O/r02_pytest_final.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1_final.json 2> $O/r02_bench_n1_final.err; echo "bench rc=$?"; tail -2 $O/r02_bench_n1_final.err
cap() {  # name regex script args...
  local name=$1 re=$2; shift 2
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$re -s 1 -c 1 -f -o $O/prof_r02_$name "$@" > $O/ncu_$name.log 2>&1
  ncu -i $O/prof_r02_$name.ncu-rep --page raw --csv > $O/prof_r02_${name}_raw.csv 2>/dev/null
  rm -f $O/prof_r02_$name.ncu-rep
}
cap appendwide eval_append_wide python tools/prof_run.py eval 64
cap exchlocal exchange_append_kernel python tools/exch_local.py 8
python tools/ncu_summary.py raw $O/prof_r02_appendwide_raw.csv $O/prof_r02_exchlocal_raw.csv > $O/r02_ncu_append_exchange_summary.txt; cat $O/r02_ncu_append_exchange_summary.txt | cut -c1-120
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<'PY'
import json
d=json.loads([x for x in open('gpurun_out/r02_bench_n1_final.json').read().splitlines() if x.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['roofline']); print(d['e2e'])
x=d['extra']; s=x['eval_sweep']; print({k:s[k] for k in ('images_s','ms_per_step','phases_ms','pool_matches_oracle','bit_exact_vs_pool')})
print(json.dumps(x['metrics']['by_size'])); print(json.dumps(x['metrics']['sort'])[:300])
PY
