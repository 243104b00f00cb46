"""Sort / metric stage timing on one GPU (development aid; bench.py reports the same numbers under extra.metrics)."""
import sys, time, torch
sys.path.insert(0, ".")
from multishiftseg_b200 import metric, _lib as L
def ev(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
g=torch.Generator(device="cuda").manual_seed(1)
mode = "cont"
args = [a for a in sys.argv[1:]]
if args and not args[0].isdigit():
    mode = args.pop(0)
for n in [int(a) for a in args] or [1<<21, 1<<23, 1<<25, 1<<27]:
    s=torch.randn(n,device="cuda",generator=g)
    if mode == "f16": s = s.half().float()
    lab=(torch.rand(n,device="cuda",generator=g)<0.05).to(torch.uint8)
    buf=metric.PairBuffer(n,"cuda"); buf.append(s,lab); m,npos,_,_=buf.read_state()
    def app():
        buf.reset(); buf.append(s,lab)
    t_a=ev(app)
    lab64=lab.long()
    def app64():
        buf.reset(); buf.append(s,lab64)
    t_a64=ev(app64)
    del lab64
    buf.reset(); buf.append(s,lab); m,npos,_,_=buf.read_state()
    k0=buf.keys.clone()
    lib=L.load(); nb=lib.mss_sort_keys_workspace_bytes(m); ws=torch.empty(nb,dtype=torch.uint8,device="cuda")
    st=torch.cuda.current_stream().cuda_stream
    import ctypes as C
    def srt():
        buf.keys.copy_(k0)
        lib.mss_eval_sort(C.byref(buf.c), m, ws.data_ptr(), nb, st)
    def cp():
        buf.keys.copy_(k0)
    t_s=ev(srt)-ev(cp)
    srt(); torch.cuda.synchronize()      # leave the streams sorted for the stages below
    neg,pos=buf.streams(m,npos)
    t_c=ev(lambda: metric.counts_from_sorted(neg,m-npos,pos,npos))
    tps,fps=metric.counts_from_sorted(neg,m-npos,pos,npos)
    t_t=ev(lambda: metric.metrics_tail(tps,fps))
    ts=[]
    for _ in range(5):
        t0=time.perf_counter(); r=metric.eval_ood_measure(s,lab); torch.cuda.synchronize(); ts.append((time.perf_counter()-t0)*1e3)
    t_e=sorted(ts)[2]
    print(f"{mode} n={n:>10d} T={tps.numel():>10d} append u8 {t_a:6.3f} ms ({(n*5+m*4)/t_a/1e6:5.0f} GB/s) i64 {t_a64:6.3f} ms | sort {t_s:8.3f} ms {m/t_s/1e6:7.2f} Gkeys/s ({m*36/t_s/1e6:7.0f} GB/s impl) | counts {t_c:7.3f} ms | tail {t_t:7.3f} ms | eval_ood_measure {t_e:8.3f} ms {n/t_e/1e3:8.1f} Mpix/s", flush=True)
