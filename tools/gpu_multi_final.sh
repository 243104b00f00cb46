#!/bin/bash
# final multi-GPU evidence: validation of every exchange form, A/B of the two streamed forms, bench with the better one
N=$1; TAG=$2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_n${N}_$TAG.log 2>&1; echo "dist_check rc=$? true=$(grep -c True gpurun_out/dist_check_n${N}_$TAG.log) false=$(grep -c False gpurun_out/dist_check_n${N}_$TAG.log)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tests/dist_c_abi_check.py > gpurun_out/dist_c_abi_n${N}_$TAG.log 2>&1; echo "c_abi rc=$?"
./tools/gpu_sweep.sh $N ${TAG}_ce --exchange stream 2>&1 | grep -v "^\*\*\*\|OMP_NUM"
./tools/gpu_sweep.sh $N ${TAG}_sm --exchange stream_sm 2>&1 | grep -v "^\*\*\*\|OMP_NUM"
BEST=$(python - <<PY
import json
def v(t):
    try:
        l=[x for x in open('gpurun_out/sweep_n${N}_${TAG}_'+t+'.json').read().splitlines() if x.startswith('{')]
        return json.loads(l[-1])['value']
    except Exception:
        return 0
print('stream' if v('ce') >= v('sm') else 'stream_sm')
PY
)
echo "bench exchange = $BEST"
MSS_BENCH_EXCHANGE=$BEST timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n${N}_$TAG.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_n${N}_$TAG.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'h2d_ceiling',d['e2e'].get('h2d_ceiling_GBs_per_gpu'),'frac',d['e2e'].get('frac_of_copy_ceiling'))
s=d['extra']['eval_sweep']
print({k:s[k] for k in ('images_s','ms_per_step','phases_ms','pool_matches_oracle','bit_exact_vs_pool')}, s['exchange'].get('kind'), s['exchange'].get('phase_ms_rank0'))
print('continuous', {k:s['continuous'][k] for k in ('metric_ms','gkeys_s','phase_ms_rank0')})
PY
