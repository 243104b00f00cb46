import sys, time, torch
sys.path.insert(0, ".")
from multishiftseg_b200 import metric, _lib as L
def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
g=torch.Generator(device="cuda").manual_seed(1)
for n in [int(a) for a in sys.argv[1:]] or [1<<21, 1<<23, 1<<25, 1<<27]:
    s=torch.randn(n,device="cuda",generator=g)
    lab=(torch.rand(n,device="cuda",generator=g)<0.05).to(torch.uint8)
    buf=metric.PairBuffer(n,"cuda"); buf.append(s,lab); m=buf.read_state()[0]
    k0,l0=buf.keys.clone(),buf.labs.clone()
    lib=L.load(); nb=lib.mss_sort_pairs_workspace_bytes(m); ws=torch.empty(nb,dtype=torch.uint8,device="cuda")
    st=torch.cuda.current_stream().cuda_stream
    def srt():
        buf.keys.copy_(k0); buf.labs.copy_(l0)
        lib.mss_sort_pairs(buf.keys.data_ptr(),buf.labs.data_ptr(),m,ws.data_ptr(),nb,st)
    def cp():
        buf.keys.copy_(k0); buf.labs.copy_(l0)
    t_s=ev(srt)-ev(cp)
    t0=time.perf_counter(); r=metric.eval_ood_measure(s,lab); torch.cuda.synchronize(); t_e=(time.perf_counter()-t0)*1e3
    t0=time.perf_counter(); r=metric.eval_ood_measure(s,lab); torch.cuda.synchronize(); t_e=(time.perf_counter()-t0)*1e3
    print(f"n={n:>10d} sort {t_s:8.3f} ms {m/t_s/1e6:7.2f} Gpairs/s ({m*44/t_s/1e6:7.0f} GB/s impl) | eval_ood_measure {t_e:8.3f} ms {n/t_e/1e3:8.1f} Mpix/s")
