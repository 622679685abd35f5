import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "missm-benchmark_b200"))
import torch
from missm_b200 import ops
dev="cuda"; M=58*257
torch.manual_seed(0)
A=torch.randn(M,1024,device=dev).bfloat16(); W1=torch.randn(4096,1024,device=dev).bfloat16()
b1=torch.randn(4096,device=dev); u=torch.empty(M,4096,device=dev,dtype=torch.bfloat16)
a=torch.empty(M,4096,device=dev,dtype=torch.bfloat16)
W2=torch.randn(1024,4096,device=dev).bfloat16(); b2=torch.randn(1024,device=dev); x=torch.randn(M,1024,device=dev)
Wq=torch.randn(3072,1024,device=dev).bfloat16(); qkv=torch.empty(M,3072,device=dev,dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(A,Wq,out=qkv)                                             # plain bf16 (QKV shape)
    ops.gemm(A,W1,bias=b1,epilogue=ops.EPI_GELU,aux_out=u,out=a)       # fc1 + QuickGELU
    ops.gemm(a,W2,bias=b2,epilogue=ops.EPI_RESID,aux_in=x,out_dtype=torch.float32)  # fc2 + residual
torch.cuda.synchronize(); print("ok")
