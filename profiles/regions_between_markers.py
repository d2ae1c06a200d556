import csv,re,sys
rows=list(csv.reader(open(sys.argv[1])))
WI=float(sys.argv[2]) if len(sys.argv)>2 else 23680
hdr=rows[1]; isrc=hdr.index('Source'); iex=hdr.index('Instructions Executed'); isam=hdr.index('# Samples')
stall=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not' not in h]
data=[r for r in rows[2:] if len(r)==len(hdr)]
tot=sum(int(r[iex]) for r in data); ts=sum(int(r[isam]) for r in data)
cur=dict(inst=0,sam=0,st={})
seen=set()
def flush(tag):
    top=sorted(cur['st'].items(), key=lambda kv:-kv[1])[:3]
    print(f"{tag:28s} inst {100*cur['inst']/tot:5.1f}% ({cur['inst']/WI:6.0f}/wi) samples {100*cur['sam']/ts:5.1f}%", [(h[6:],v) for h,v in top])
    cur.update(inst=0,sam=0,st={})
for idx,r in enumerate(data):
    s=r[isrc]
    cur['inst']+=int(r[iex]); cur['sam']+=int(r[isam])
    for i,h in stall:
        v=int(r[i] or 0)
        if v: cur['st'][h]=cur['st'].get(h,0)+v
    m=re.search(r'(BAR\.SYNC|ATOMG|STG|LDS\.U16|CALL|HSET2|F2FP|SHFL|LDS\.128|HMNMX2)',s)
    if m:
        k=m.group(1)
        if k=='BAR.SYNC': flush(f'{idx}:BAR'); seen=set()
        elif k not in seen and k in ('ATOMG','STG','LDS.U16','HSET2','SHFL','LDS.128','F2FP'):
            seen.add(k); flush(f'{idx}:first {k}')
flush('end')
print('total inst',tot,'per wi',tot/WI)
