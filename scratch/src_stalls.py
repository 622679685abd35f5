import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
print(rows[0][1][:80])
hdr=rows[1]
i_s=hdr.index('# Samples'); i_src=hdr.index('Source'); i_ex=hdr.index('Instructions Executed')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data=[]; tot=0; agg={}
for k,r in enumerate(rows[2:]):
    if len(r)<=max(stall_cols): continue
    try: n=int(r[i_s])
    except: continue
    tot+=n
    for i in stall_cols:
        if r[i] not in ('','0'): agg[hdr[i]]=agg.get(hdr[i],0)+int(r[i])
    data.append((n,k,r))
print("total",tot, sorted(agg.items(),key=lambda x:-x[1])[:8])
top=int(sys.argv[2]) if len(sys.argv)>2 else 30
for n,k,r in sorted(data,reverse=True)[:top]:
    st={hdr[i][6:]:int(r[i]) for i in stall_cols if r[i] not in('','0')}
    t=sorted(st.items(),key=lambda x:-x[1])[:2]
    print(f"{n:6d} {100*n/tot:5.1f}% #{k:4d} ex={r[i_ex]:>8s} {r[i_src].strip()[:64]:64s} {t}")
