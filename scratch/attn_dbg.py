import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "missm-benchmark_b200"))
import torch
from missm_b200 import ops
S,H,N=int(sys.argv[1]),int(sys.argv[2]),int(sys.argv[3])
D=H*64
torch.manual_seed(N)
qkv=(torch.randn(S*N,3*D,device="cuda")*0.7).bfloat16()
lay=ops.SeqLayout.spatial(S,N)
print("fwd..",flush=True)
out,lse=ops.attention_fwd(qkv,lay,H); torch.cuda.synchronize(); print("fwd ok",flush=True)
d_out=torch.randn(S*N,D,device="cuda").bfloat16()
dqkv,dcs=ops.attention_bwd(qkv,out,lse,d_out,lay,H,0.125); torch.cuda.synchronize(); print("bwd ok",flush=True)
