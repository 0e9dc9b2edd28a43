"""Top stalled SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > K.csv`:
    python tools/ncu_hotspots.py K.csv [N]      (samples, instruction, the three largest stall reasons)"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
recs=[]
sect=0
for n,r in enumerate(data):
    if len(r)<len(hdr): continue
    if r[ix['# Samples']]=='# Samples': sect+=1; continue
    if sect>0: break
    s=int(r[ix['# Samples']])
    st=[x for x in sorted(((int(r[ix[h]]),h[6:]) for h in stalls), reverse=True)[:3] if x[0]>0]
    recs.append((n,s,r[ix['Source']].strip(),st,int(r[ix['Instructions Executed']])))
tot=sum(x[1] for x in recs)
print(len(recs),"instructions",tot,"samples")
agg={}
for n,s,src,st,ie in recs:
    pass
top=int(sys.argv[2]) if len(sys.argv)>2 else 40
if len(sys.argv)>3:
    for n,s,src,st,ie in recs: print("%4d %5d %-70s %s"%(n,s,src[:70],st))
else:
    for n,s,src,st,ie in sorted(recs,key=lambda x:-x[1])[:top]:
        print("%4d %5d %-70s %s"%(n,s,src[:70],st))
