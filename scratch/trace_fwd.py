import sys
ev=[tuple(map(int,l.split())) for l in open(sys.argv[1])]
lo,hi=int(sys.argv[2]),int(sys.argv[3])
names={1:'S.issue',2:'S.commit',3:'pfull.ok',4:'oempty.ok',5:'PV.commit',10:'sm.wait_s',11:'sm.s_ok',12:'sm.p1_done',13:'sm.ofull_ok',14:'sm.epi_done',15:'sm.sync2',16:'sm.p_arrive'}
e=[(c,s,names[t],n) for s,t,n,c in ev]
e.sort()
t0=e[0][0]
print("total", e[-1][0]-t0)
for c,s,nm,n in e[lo:hi]:
    print(f"  {c-t0:8d} {'MMA' if s==0 else 'SM'+str(s-1):4s} t={n:3d} {nm}")
