#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_metrics.py -x -q -m gpu -k "partition or histogram" 2>&1 | tail -5
MSS_PARTITION_NO_BULK=1 timeout 600 python -m pytest tests/test_gpu_metrics.py -x -q -m gpu -k "partition" 2>&1 | tail -3
python - <<'PY'
import torch, numpy as np, sys, ctypes as C
sys.path.insert(0,'.')
from multishiftseg_b200.evaluator import CudaBackend
from multishiftseg_b200 import metric, _lib as L
be=CudaBackend('cuda')
g=torch.Generator(device='cuda').manual_seed(0)
n=1<<28
s=torch.randn(n,device='cuda',generator=g); lab=(torch.rand(n,device='cuda',generator=g)<0.05).to(torch.uint8)
buf=metric.PairBuffer(n,'cuda'); buf.append(s,lab); m,npos,_,_=buf.read_state()
neg,pos=buf.streams(m,npos)
q=torch.quantile(s[:1<<20],torch.linspace(0.125,0.875,7,device='cuda'))
# splitters in key space: descending score = ascending key; take keys of the quantile scores
kb=metric.PairBuffer(16,'cuda'); kb.append(q.flip(0).contiguous(), torch.zeros(7,dtype=torch.uint8,device='cuda'))
spl=sorted(int(x)&0xffffffff for x in kb.keys[:7].tolist())
import time
for name in ('bulk',):
    cn,cp=be.partition_count2(buf,m-npos,npos,spl,8)
    bk=[torch.empty(max(a+b,1)+8,dtype=torch.int32,device='cuda') for a,b in zip(cn,cp)]
    def run():
        be.partition_scatter2(buf,m-npos,npos,spl,8,[t.data_ptr() for t in bk],[0]*8,[a for a in cn])
    run(); torch.cuda.synchronize()
    ts=[]
    for _ in range(5):
        t0=time.perf_counter(); run(); torch.cuda.synchronize(); ts.append((time.perf_counter()-t0)*1e3)
    tc=[]
    for _ in range(5):
        t0=time.perf_counter(); be.partition_count2(buf,m-npos,npos,spl,8); tc.append((time.perf_counter()-t0)*1e3)
    print(name, 'keys',m,'counts',cn, 'scatter ms',sorted(ts)[2],'count ms',sorted(tc)[2], 'GB/s scatter', m*8/sorted(ts)[2]/1e6)
PY
