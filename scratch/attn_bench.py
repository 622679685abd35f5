import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "missm-benchmark_b200"))
import torch
from missm_b200 import ops
dev="cuda"
def rel(a,b): return ((a.float()-b.float()).norm()/b.float().norm()).item()
def t(fn,it=20):
    for _ in range(3): fn()
    e0,e1=torch.cuda.Event(True),torch.cuda.Event(True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it
def run(S,H,N,bench=False):
    D=H*64
    torch.manual_seed(N)
    qkv=(torch.randn(S*N,3*D,device=dev)*0.7).bfloat16()
    lay=ops.SeqLayout.spatial(S,N)
    out,lse=ops.attention_fwd(qkv,lay,H)
    torch.cuda.synchronize()
    f=qkv.float().view(S,N,3,H,64).permute(2,0,3,1,4).contiguous().requires_grad_(True)
    ref=torch.softmax(f[0]@f[1].transpose(-1,-2),-1)@f[2]
    ref_o=ref.permute(0,2,1,3).reshape(S*N,D)
    d_out=torch.randn(S*N,D,device=dev).bfloat16()
    ref_o.backward(d_out.float())
    dqkv,dcs=ops.attention_bwd(qkv,out,lse,d_out,lay,H,0.125)
    torch.cuda.synchronize()
    g=f.grad.permute(1,3,0,2,4).reshape(S*N,3*D).clone(); g[:,:D]*=0.125
    print(f"S={S} H={H} N={N}: fwd {rel(out,ref_o):.2e} dq {rel(dqkv[:,:D],g[:,:D]):.2e} dk {rel(dqkv[:,D:2*D],g[:,D:2*D]):.2e} dv {rel(dqkv[:,2*D:],g[:,2*D:]):.2e} cs {rel(dcs,g.sum(0)):.2e}", flush=True)
    if bench:
        fl=4*N*N*64*S*H
        ms=t(lambda: ops.attention_fwd(qkv,lay,H)); print(f"  fwd {ms*1e3:.1f} us {fl/ms/1e9:.0f} TF/s")
        ms=t(lambda: ops.attention_bwd(qkv,out,lse,d_out,lay,H,0.125)); print(f"  bwd {ms*1e3:.1f} us {2.5*fl/ms/1e9:.0f} TF/s", flush=True)
for cfg in [(1,1,257),(3,16,257),(2,2,130),(2,3,64),(1,2,272),(2,2,16),(4,2,200),(30,16,96)]:
    run(*cfg)
run(58,16,257,True)
run(64,16,257,True)
