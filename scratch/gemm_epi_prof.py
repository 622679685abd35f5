"""The epilogue-heavy GEMM launches of a layer at the bench shape, for `ncu --set full --import-source on`:
out-proj + residual (f32 RMW), fc2 + residual, fc1 + QuickGELU, fc2-dgrad x QuickGELU'."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "missm-benchmark_b200"))
import torch  # noqa: E402
from missm_b200 import ops  # noqa: E402

dev, M, D, F = "cuda", 58 * 257, 1024, 4096
bf = torch.bfloat16
mk = lambda *s: (torch.randn(*s, device=dev) * 0.05).to(bf)      # noqa: E731
h, attn, dy, a, u = mk(M, D), mk(M, D), mk(M, D), mk(M, F), mk(M, F)
wo, w1, w2 = mk(D, D), mk(F, D), mk(D, F)
bo, b1, b2 = [torch.randn(n, device=dev) for n in (D, F, D)]
x = torch.randn(M, D, device=dev)
for _ in range(2):
    ops.gemm(attn, wo, bias=bo, epilogue=ops.EPI_RESID, aux_in=x, out_dtype=torch.float32)
    ops.gemm(a, w2, bias=b2, epilogue=ops.EPI_RESID, aux_in=x, out_dtype=torch.float32)
    ops.gemm(h, w1, bias=b1, epilogue=ops.EPI_GELU, aux_out=u)
    ops.gemm(dy, w2, b_mn=True, epilogue=ops.EPI_DGELU, aux_in=u)
torch.cuda.synchronize()
print("ok")
