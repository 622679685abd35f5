"""The 12 GEMM launches of one ViT-L layer (fwd + bwd) at the bench row count, timed one by one (CUDA events, L2
flushed between launches) with tile N forced to 256 / 128 and on auto, next to torch.matmul (cuBLAS) on the same
shape.  `python scratch/gemm_shapes.py [M]`"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "missm-benchmark_b200"))
import torch  # noqa: E402
from missm_b200 import ops  # noqa: E402

dev = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 58 * 257
D, F = 1024, 4096
bf = torch.bfloat16
torch.manual_seed(0)


def mk(*s):
    return (torch.randn(*s, device=dev) * 0.05).to(bf)


h, qkv, attn, dy, a, u = mk(M, D), mk(M, 3 * D), mk(M, D), mk(M, D), mk(M, F), mk(M, F)
wqkv, wo, w1, w2 = mk(3 * D, D), mk(D, D), mk(F, D), mk(D, F)
bq, bo, b1, b2 = [torch.randn(n, device=dev) for n in (3 * D, D, F, D)]
x = torch.randn(M, D, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

cases = [
    ("fwd qkv        N=3072 K=1024", lambda **k: ops.gemm(h, wqkv, bias=bq, scale_cols=D, col_scale=0.125, **k), 2.0 * M * 3 * D * D, lambda: h @ wqkv.t()),
    ("fwd out-proj   N=1024 K=1024 resid", lambda **k: ops.gemm(attn, wo, bias=bo, epilogue=ops.EPI_RESID, aux_in=x, out_dtype=torch.float32, **k), 2.0 * M * D * D, lambda: attn @ wo.t()),
    ("fwd fc1+gelu   N=4096 K=1024", lambda **k: ops.gemm(h, w1, bias=b1, epilogue=ops.EPI_GELU, aux_out=u, **k), 2.0 * M * F * D, lambda: h @ w1.t()),
    ("fwd fc2        N=1024 K=4096 resid", lambda **k: ops.gemm(a, w2, bias=b2, epilogue=ops.EPI_RESID, aux_in=x, out_dtype=torch.float32, **k), 2.0 * M * D * F, lambda: a @ w2.t()),
    ("bwd d_attn     N=1024 K=1024 dgrad", lambda **k: ops.gemm(dy, wo, b_mn=True, **k), 2.0 * M * D * D, lambda: dy @ wo),
    ("bwd d_wo       wgrad 1024x1024 K=M", lambda **k: ops.gemm(dy, attn, a_mn=True, b_mn=True, out_dtype=torch.float32, **k), 2.0 * M * D * D, lambda: dy.t() @ attn),
    ("bwd d_wqkv     wgrad 3072x1024 K=M", lambda **k: ops.gemm(qkv, h, a_mn=True, b_mn=True, out_dtype=torch.float32, **k), 2.0 * M * 3 * D * D, lambda: qkv.t() @ h),
    ("bwd d_h(qkv)   N=1024 K=3072 dgrad", lambda **k: ops.gemm(qkv, wqkv, b_mn=True, **k), 2.0 * M * 3 * D * D, lambda: qkv @ wqkv),
    ("bwd d_w2       wgrad 1024x4096 K=M", lambda **k: ops.gemm(dy, a, a_mn=True, b_mn=True, out_dtype=torch.float32, **k), 2.0 * M * D * F, lambda: dy.t() @ a),
    ("bwd d_u dgelu  N=4096 K=1024", lambda **k: ops.gemm(dy, w2, b_mn=True, epilogue=ops.EPI_DGELU, aux_in=u, **k), 2.0 * M * D * F, lambda: dy @ w2),
    ("bwd d_w1       wgrad 4096x1024 K=M", lambda **k: ops.gemm(a, h, a_mn=True, b_mn=True, out_dtype=torch.float32, **k), 2.0 * M * D * F, lambda: a.t() @ h),
    ("bwd d_h(fc1)   N=1024 K=4096 dgrad", lambda **k: ops.gemm(a, w1, b_mn=True, **k), 2.0 * M * D * F, lambda: a @ w1),
]


def timeit(fn, reps=6):
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3


tot = {"auto": 0.0, "256": 0.0, "128": 0.0, "cublas": 0.0, "best": 0.0}
print(f"M = {M}; us per launch (best of 6, L2 flushed) and PFLOP/s")
for name, fn, flop, ref in cases:
    t = {"auto": timeit(lambda: fn()), "256": timeit(lambda: fn(force_bn=256)), "128": timeit(lambda: fn(force_bn=128)),
         "cublas": timeit(ref)}
    for k in t:
        tot[k] += t[k]
    tot["best"] += min(t["256"], t["128"])
    print(f"{name:38s} auto {t['auto']:7.1f} ({flop / t['auto'] / 1e9:5.2f})  bn256 {t['256']:7.1f}  bn128 {t['128']:7.1f}  cuBLAS {t['cublas']:7.1f} ({flop / t['cublas'] / 1e9:5.2f})")
print("layer total: " + "  ".join(f"{k} {v:8.1f} us" for k, v in tot.items()))
