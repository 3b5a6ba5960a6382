import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
n=len(data)
W=int(sys.argv[2]) if len(sys.argv)>2 else 100
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot_ex=0;tot_s=0
for i in range(0,n,W):
    blk=data[i:i+W]
    ex=sum(int(r[ix['Instructions Executed']]) for r in blk)
    smp=sum(int(r[ix['# Samples']]) for r in blk)
    st={s:sum(int(r[ix[s]]) for r in blk) for s in stalls}
    top=sorted(st.items(), key=lambda x:-x[1])[:4]
    ops={}
    for r in blk:
        op=r[ix['Source']].split()[0] if not r[ix['Source']].strip().startswith('@') else r[ix['Source']].split()[1]
        op=op.split('.')[0]; ops[op]=ops.get(op,0)+1
    topo=sorted(ops.items(), key=lambda x:-x[1])[:3]
    tot_ex+=ex;tot_s+=smp
    print(f"{i:5d} ex={ex/1e6:8.1f}M smp={smp:6d} " + ' '.join(f"{k[6:]}:{v}" for k,v in top) + ' | ' + ' '.join(f"{k}:{v}" for k,v in topo))
print(tot_ex/1e6, tot_s)
