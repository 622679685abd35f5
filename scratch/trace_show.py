import sys
ev=[tuple(map(int,l.split())) for l in open(sys.argv[1])]
lo,hi=int(sys.argv[2]),int(sys.argv[3])
names={1:'S.issue',2:'S.commit',3:'pfull.ok',4:'accempty.ok',5:'PV.commit',10:'wg.wait_s',11:'wg.s_ok',12:'wg.p_arrive',13:'wg.accfull_ok',14:'wg.epi_done'}
for base,label in ((0,'DKV'),(3,'DQ')):
    e=[(c,s-base,names[t],n) for s,t,n,c in ev if base<=s<base+3]
    e.sort()
    t0=e[0][0]
    print(label, "total", e[-1][0]-t0)
    for c,s,nm,n in e[lo:hi]:
        print(f"  {c-t0:8d} {'MMA' if s==0 else 'WG'+str(s-1):4s} n={n:3d} {nm}")
