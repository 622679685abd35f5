import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "missm-benchmark_b200"))
import torch, ctypes
from missm_b200 import ops
from missm_b200._lib import lib, check, stream_ptr
def t(fn,it=20):
    for _ in range(3): fn()
    e0,e1=torch.cuda.Event(True),torch.cuda.Event(True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it*1e3
S,H,N=58,16,257
D=H*64
qkv=(torch.randn(S*N,3*D,device="cuda")*0.7).bfloat16()
lay=ops.SeqLayout.spatial(S,N)
out,lse=ops.attention_fwd(qkv,lay,H)
d_out=torch.randn(S*N,D,device="cuda").bfloat16()
print(os.environ.get("MISSM_ATTN_TAIL_SKIP"), os.environ.get("MISSM_ATTN_NO_TAIL"), "fwd %.1f us  bwd %.1f us" % (t(lambda: ops.attention_fwd(qkv,lay,H)), t(lambda: ops.attention_bwd(qkv,out,lse,d_out,lay,H,0.125))))
