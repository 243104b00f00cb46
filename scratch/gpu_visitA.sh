#!/bin/bash
# visit A: new mask GEMM (tests, timing), cfg-5 sweep, then the whole GPU suite
TAG=${1:-v4}
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_maskgemm.py -x -q > $O/pytest_maskgemm_$TAG.log 2>&1; echo "maskgemm pytest rc=$?"; tail -15 $O/pytest_maskgemm_$TAG.log
timeout 200 python scratch/bench_maskgemm.py > $O/maskgemm_$TAG.log 2>&1; cat $O/maskgemm_$TAG.log
timeout 300 python bench_sweep.py --cfg 5 --steps 2 --warmup 1 > $O/sweep5_$TAG.json 2> $O/sweep5_$TAG.err; echo "sweep5 rc=$?"; cat $O/sweep5_$TAG.json; tail -5 $O/sweep5_$TAG.err
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; tail -2 $O/smoke_$TAG.log
