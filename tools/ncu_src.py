"""Top stall lines of one kernel from `ncu --page source --csv` output (SASS view)."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
N=int(sys.argv[2]) if len(sys.argv)>2 else 25
start=None
for i,r in enumerate(rows):
    if len(r)>3 and r[0]=='Address' and r[1]=='Source':
        hdr=r; start=i+1; break
ix={h:i for i,h in enumerate(hdr)}
key='# Samples'
data=[r for r in rows[start:] if len(r)==len(hdr) and r[0]!='Address']
def f(x):
    try: return float(x)
    except: return 0.0
tot=sum(f(r[ix[key]]) for r in data)
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={s:sum(f(r[ix[s]]) for r in data) for s in stalls}
print("total samples",tot, {k:round(v/tot*100,1) for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:8]})
top=sorted(data,key=lambda r:-f(r[ix[key]]))[:N]
for r in top:
    ss=sorted(((f(r[ix[s]]),s) for s in stalls), reverse=True)[:2]
    print(f"{f(r[ix[key]])/tot*100:5.1f}% {r[0][-5:]} {r[ix['Source']][:70]:70s} {ss[0][1]}:{ss[0][0]:.0f} {ss[1][1]}:{ss[1][0]:.0f}")
