import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "missm-benchmark_b200"))
import torch
from missm_b200 import ops
dev="cuda"; Mt=58*257
def bench(M,N,K,split_k,iters=30):
    A=torch.randn(K,M,device=dev).bfloat16(); B=torch.randn(K,N,device=dev).bfloat16()
    out=torch.empty(M,N,device=dev,dtype=torch.float32)
    for _ in range(3): ops.gemm(A,B,a_mn=True,b_mn=True,out=out,split_k=split_k)
    e0,e1=torch.cuda.Event(True),torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): ops.gemm(A,B,a_mn=True,b_mn=True,out=out,split_k=split_k)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/iters
    print(f"wgrad M{M} N{N} K{K} split_k={split_k}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TF/s", flush=True)
for (M,N) in [(3072,1024),(1024,1024),(4096,1024),(1024,4096)]:
    for sk in (1,0,-1):
        bench(M,N,Mt,sk)
