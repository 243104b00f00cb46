"""Opcode mix + top-stall lines of one kernel from `ncu --page source --csv` (SASS view). usage: ncu_opmix.py file.csv [pairs]"""
import csv,collections,re,sys
rows=list(csv.reader(open(sys.argv[1])))
for i,r in enumerate(rows):
    if len(r)>3 and r[0]=='Address' and r[1]=='Source':
        hdr=r; start=i+1; break
ix={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[start:] if len(r)==len(hdr)]
tot=sum(float(r[ix['Instructions Executed']]) for r in data)
pairs=float(sys.argv[2]) if len(sys.argv)>2 else None
agg=collections.Counter()
for r in data:
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']])
    op=m.group(2).split('.')[0] if m else '?'
    agg[op]+=float(r[ix['Instructions Executed']])
print("total warp instr",tot)
for k,v in agg.most_common(28): print(f"{k:12s} {v/tot*100:5.1f}%  {v/1e6:8.2f}M" + (f"  {v/pairs:6.2f}/unit" if pairs else ""))
