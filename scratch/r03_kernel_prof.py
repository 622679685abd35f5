"""Round-2 kernels at the bench shape for `ncu --set full`: tcgen05 attention forward / backward (58 sequences x 16 heads,
N = 257) and the implicit-GEMM patch embedding (58 present images)."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "missm-benchmark_b200"))
import torch
from missm_b200 import ops

S, H, N = 58, 16, 257
D = H * 64
torch.manual_seed(0)
qkv = (torch.randn(S * N, 3 * D, device="cuda") * 0.7).bfloat16()
lay = ops.SeqLayout.spatial(S, N)
d_out = torch.randn(S * N, D, device="cuda").bfloat16()
px = torch.randn(64, 3, 224, 224, device="cuda")
idx = torch.arange(S, device="cuda", dtype=torch.int32)
w = ops.cast_bf16(torch.randn(D, 588, device="cuda") * 0.05, cols_dst=640)
pos = torch.randn(257, D, device="cuda")
tok = torch.empty(S * 257, D, device="cuda")
for _ in range(2):
    out, lse = ops.attention_fwd(qkv, lay, H)
    ops.attention_bwd(qkv, out, lse, d_out, lay, H, 0.125)
    assert ops.patch_embed_implicit(px, w, pos, tok, 14, 1, sample_index=idx, n_samples=S)
torch.cuda.synchronize()
print("ok")
