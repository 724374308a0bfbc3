import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
iS=hdr.index('Source'); iN=hdr.index('# Samples'); iE=hdr.index('Instructions Executed')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
seg=0; segs=[]; cur={'n':0,'samples':0,'exec':0,'stalls':{}, 'first':None}
def flush():
    global cur
    segs.append(cur); cur={'n':0,'samples':0,'exec':0,'stalls':{}, 'first':None}
for r in data:
    src=r[iS]
    cur['n']+=1; cur['samples']+=int(r[iN] or 0); cur['exec']+=int(r[iE] or 0)
    for i in stall_cols:
        v=int(r[i] or 0)
        if v: cur['stalls'][hdr[i]]=cur['stalls'].get(hdr[i],0)+v
    if 'BAR.SYNC' in src or 'WARPSYNC' in src or 'EXIT' in src:
        cur['end']=src.strip(); flush()
tot=sum(s['samples'] for s in segs)
for k,s in enumerate(segs):
    top=sorted(s['stalls'].items(),key=lambda x:-x[1])[:5]
    print(f"seg{k}: ninstr={s['n']:5d} exec={s['exec']:>12d} samples={s['samples']:7d} ({100*s['samples']/tot:5.1f}%) end={s.get('end','')[:30]:30s} {top}")
