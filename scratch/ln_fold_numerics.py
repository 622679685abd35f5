"""The bf16-cancellation check SURVEY.md section 7 asks for before LayerNorm is folded into a GEMM:
   standard:  y = bf16(LN(x)) . bf16(W)^T + b                      (what the kernels compute today)
   folded:    y = rstd * (bf16(x) . bf16(W gamma)^T - mean * c) + (W beta + b),   c = rowsum(bf16(W gamma))
both against the fp64 result, on residual-stream-like rows: unit-scale channels plus a few massive outlier channels
(CLIP ViT-L residual streams carry channels of magnitude 10-300) and a non-zero row mean.  CPU only."""
import torch

torch.manual_seed(0)
M, K, N = 2048, 1024, 1024
W = torch.randn(N, K, dtype=torch.float64) * 0.02
b = torch.randn(N, dtype=torch.float64) * 0.1
gamma = 1.0 + 0.2 * torch.randn(K, dtype=torch.float64)
beta = 0.1 * torch.randn(K, dtype=torch.float64)


def bf(t):
    return t.to(torch.float32).to(torch.bfloat16).to(torch.float64)


def run(outlier, mean_shift):
    x = torch.randn(M, K, dtype=torch.float64)
    idx = torch.randperm(K)[:4]
    x[:, idx] += outlier * (1.0 + 0.1 * torch.randn(M, 4, dtype=torch.float64))
    x += mean_shift
    x = x.to(torch.float32).to(torch.float64)                 # the fp32 residual stream
    mean = x.mean(1, keepdim=True)
    var = ((x - mean) ** 2).mean(1, keepdim=True)
    rstd = (var + 1e-5).rsqrt()
    h = (x - mean) * rstd * gamma + beta
    ref = h @ W.t() + b
    std = bf(h) @ bf(W).t() + b
    Wg = bf(W * gamma)
    c = Wg.sum(1)
    fold = rstd * (bf(x) @ Wg.t() - mean * c) + (W @ beta + b)
    # one-pass variance from partial sums in fp32, as an epilogue would accumulate it
    xf = x.to(torch.float32)
    m1 = xf.sum(1, keepdim=True) / K
    v1 = (xf * xf).sum(1, keepdim=True) / K - m1 * m1
    rstd1 = (v1.double() + 1e-5).rsqrt()
    rel = lambda a: ((a - ref).norm() / ref.norm()).item()
    return rel(std), rel(fold), ((rstd1 - rstd).abs() / rstd).max().item()


print("outlier  mean   rel.err standard   rel.err folded   one-pass rstd rel.err")
for outlier, shift in [(0, 0.0), (30, 0.0), (300, 0.0), (30, 0.5), (300, 2.0), (0, 5.0)]:
    s, f, r = run(outlier, shift)
    print(f"{outlier:7d} {shift:5.1f}   {s:14.2e}   {f:14.2e}   {r:12.2e}")
