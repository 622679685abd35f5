import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "missm-benchmark_b200"))
import torch
from missm_b200 import ops
def t(fn,it=20):
    for _ in range(3): fn()
    e0,e1=torch.cuda.Event(True),torch.cuda.Event(True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it*1e3
M,D=14906,1024
torch.manual_seed(0)
x=torch.randn(M,D,device="cuda"); g=torch.randn(D,device="cuda"); b=torch.randn(D,device="cuda")
y,mean,rstd=ops.layernorm_fwd(x,g,b,1e-5)
dy=torch.randn(M,D,device="cuda").bfloat16(); dres=torch.randn(M,D,device="cuda")
big=torch.empty(64<<20,device="cuda")   # 256 MB: flush L2 between calls
def bwd(): 
    big.zero_()
    return ops.layernorm_bwd(dy,x,mean,rstd,g,dres=dres,want_bf16=True)
def flush(): big.zero_()
tb=t(bwd); tf=t(flush)
dx,dxb,dg,db,dcs=bwd()
xr=x.clone().requires_grad_(True)
torch.nn.functional.layer_norm(xr,(D,),g,b,1e-5).backward(dy.float())
ref=xr.grad+dres
rel=lambda a,b:((a.float()-b.float()).norm()/b.float().norm()).item()
byt=M*D*(2+4+4+4+2)
print(os.environ.get("MISSM_LN_BWD_1WARP"), "ln_bwd %.1f us (incl reduce) -> %.0f GB/s; err dx %.1e dg %.1e cs %.1e" % (tb-tf, byt/(tb-tf)/1e3, rel(dx,ref), rel(dg,(dy.float()*((x-mean[:,None])*rstd[:,None])).sum(0)), rel(dcs,ref.sum(0))))
