"""Summarise an `ncu --page source --csv` dump: opcode mix, hot SASS ranges (by executed count), stall samples."""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
thr=float(sys.argv[2]) if len(sys.argv)>2 else 5e5
hdr=[r for r in rows if 'Instructions Executed' in r][0]
iS=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iSamp=hdr.index('# Samples'); iT=hdr.index('Thread Instructions Executed')
data=[d for d in rows if len(d)>iT and d[iE].isdigit()]
def opof(s):
    t=s.strip().split()
    if not t: return '?'
    op=t[1] if t[0].startswith('@') and len(t)>1 else t[0]
    return op.split('.')[0]
ops=collections.Counter(); samp=collections.Counter(); tot=0; totS=0; thrI=0
for d in data:
    op=opof(d[iS]); e=int(d[iE]); ops[op]+=e; tot+=e; s=int(d[iSamp]); samp[op]+=s; totS+=s; thrI+=int(d[iT])
print('total warp instr',tot,'thread instr',thrI,'avg active',thrI/max(tot,1),'samples',totS,'n sass',len(data))
for op,c in ops.most_common(24): print(f"{op:10s} {c/tot*100:6.2f}%  samples {samp[op]/max(totS,1)*100:6.2f}%")
ex=[int(d[iE]) for d in data]
start=None
for i,e in enumerate(ex+[0]):
    if e>thr and start is None: start=i
    if e<=thr and start is not None:
        seg=ex[start:i]; print(f"  [{start:4d},{i:4d}) n={i-start:4d} avg exec {sum(seg)/len(seg)/1e6:7.2f}M total {sum(seg)/1e6:8.1f}M ({sum(seg)/tot*100:5.1f}%) samples {sum(int(d[iSamp]) for d in data[start:i])/max(totS,1)*100:5.1f}%")
        start=None
if len(sys.argv)>3:
    a,b=map(int,sys.argv[3].split(':'))
    for i in range(a,b): print(i, data[i][iS].strip(), data[i][iE], data[i][iSamp])
