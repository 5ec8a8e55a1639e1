import csv,sys,collections,subprocess
rep,kern=sys.argv[1],sys.argv[2]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--kernel-name','regex:'+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; agg=collections.defaultdict(lambda:[0,0,'']); 
hdr=None
for r in rows:
    if len(r)>=2 and r[0] in ('File Path','File Name'): cur=r[1].split('/')[-1]; continue
    if len(r)>2 and r[0]=='Line No': hdr=r; ix={h:i for i,h in enumerate(hdr)}; continue
    if hdr and len(r)>=len(hdr)-2 and r[0].isdigit():
        vi=r[hdr.index('Instructions Executed')]; vs=r[hdr.index('# Samples')]
        if not vi.isdigit(): continue
        a=agg[(cur,int(r[0]))]; a[0]+=int(vi); a[1]+=int(vs) if vs.isdigit() else 0; a[2]=r[1]
ti=sum(a[0] for a in agg.values()); ts=sum(a[1] for a in agg.values())
n=int(sys.argv[3]) if len(sys.argv)>3 else 40
print('total instr',ti,'samples',ts)
for k,a in sorted(agg.items(),key=lambda kv:-kv[1][1])[:n]:
    print(f'{k[0]:20s}:{k[1]:4d} instr {100*a[0]/ti:5.1f}% samples {100*a[1]/ts:5.1f}%  {a[2][:90]}')
