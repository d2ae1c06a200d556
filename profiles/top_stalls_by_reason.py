import csv,re,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; isrc=hdr.index('Source'); iex=hdr.index('Instructions Executed'); isam=hdr.index('# Samples')
stall=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not' not in h]
data=[r for r in rows[2:] if len(r)==len(hdr)]
ts=sum(int(r[isam]) for r in data)
tot={}
for r in data:
    for i,h in stall:
        v=int(r[i] or 0)
        if v: tot[h]=tot.get(h,0)+v
print('samples',ts,{k[6:]:v for k,v in sorted(tot.items(), key=lambda kv:-kv[1])})
top=sorted(range(len(data)), key=lambda k:-int(data[k][isam]))[:int(sys.argv[2]) if len(sys.argv)>2 else 28]
for k in sorted(top):
    r=data[k]
    st=sorted(((int(r[i] or 0),h[6:]) for i,h in stall), reverse=True)[:2]
    print(k, r[isrc].strip()[:70], r[isam], st)
