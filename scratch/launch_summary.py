import csv,collections,re,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r: hdr=r
        continue
    d=dict(zip(hdr,r))
    try: v=float(d['Metric Value'].replace(',',''))
    except: continue
    n=re.sub(r'\(.*','',d['Kernel Name'])[:70]
    u=d['Metric Unit']
    if u=='ns': v/=1e3
    elif u=='ms': v*=1e3
    agg[n][0]+=1; agg[n][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(),key=lambda x:-x[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print(f"{v[1]/1e3:9.2f} ms {v[0]:5d} {100*v[1]/tot:5.1f}% {v[1]/v[0]:8.1f} us/launch  {k}")
print(f"total {tot/1e3:.2f} ms")
