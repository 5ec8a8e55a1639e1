import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
ix={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)>=len(hdr)-2 and r[ix['# Samples']].isdigit()]
tot=sum(int(r[ix['# Samples']]) for r in data)
print('instructions',len(data),'total samples',tot)
keys=['stall_long_sb','stall_wait','stall_short_sb','stall_no_inst','stall_mio','stall_selected','stall_not_selected','stall_math','stall_branch_resolving','stall_dispatch','stall_lg','stall_barrier']
for k in keys:
    print(k, sum(int(r[ix[k]]) for r in data))
key=sys.argv[2] if len(sys.argv)>2 else 'stall_long_sb'
top=sorted(data,key=lambda r:-int(r[ix[key]]))[:int(sys.argv[3]) if len(sys.argv)>3 else 25]
for r in top:
    print(r[ix[key]], r[ix['# Samples']], r[ix['Address']][-5:], r[ix['Source']][:100])
