"""GPU bring-up check of the tcgen05 GEMM against torch (fp32 reference on bf16-rounded inputs)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "missm-benchmark_b200"))
import torch
from missm_b200 import ops

torch.manual_seed(0)
dev = "cuda"
def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()

ok = True
def report(name, err, tol=6e-3):
    global ok
    good = err < tol
    ok &= good
    print(f"{'PASS' if good else 'FAIL'} {name}: rel={err:.3e}", flush=True)

for (M, N, K) in [(128, 256, 64), (128, 128, 128), (300, 520, 200), (1000, 1024, 1024), (257 * 7, 3072, 1024)]:
    for a_mn in (False, True):
        for b_mn in (False, True):
            for bn in (128, 256):
                Mp = M if not a_mn else (M + 7) // 8 * 8   # MN-major needs 16B-aligned pitch
                Np = N
                A = torch.randn(Mp, K, device=dev).bfloat16()
                B = torch.randn(Np, K, device=dev).bfloat16()
                ref = A.float() @ B.float().t()
                a_in = A.t().contiguous() if a_mn else A
                b_in = B.t().contiguous() if b_mn else B
                out = ops.gemm(a_in, b_in, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, force_bn=bn, split_k=1)
                torch.cuda.synchronize()
                report(f"gemm M{Mp} N{Np} K{K} a_mn={int(a_mn)} b_mn={int(b_mn)} bn={bn}", rel(out, ref), 1e-5)

# epilogues
M, N, K = 700, 1024, 512
A = torch.randn(M, K, device=dev).bfloat16(); B = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
bias = torch.randn(N, device=dev)
acc = A.float() @ B.float().t()
out = ops.gemm(A, B, bias=bias, scale_cols=512, col_scale=0.125)
ref = acc + bias; ref[:, :512] *= 0.125
report("epi linear+bias+scale bf16", rel(out, ref))
u = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
out = ops.gemm(A, B, bias=bias, epilogue=ops.EPI_GELU, aux_out=u)
pre = acc + bias
report("epi gelu u", rel(u, pre)); report("epi gelu act", rel(out, pre * torch.sigmoid(1.702 * pre)))
res = torch.randn(M, N, device=dev)
out = ops.gemm(A, B, bias=bias, epilogue=ops.EPI_RESID, aux_in=res, out_dtype=torch.float32)
report("epi resid", rel(out, res + pre), 1e-5)
res2 = res.clone()
ops.gemm(A, B, bias=bias, epilogue=ops.EPI_RESID, aux_in=res2, out=res2)
report("epi resid inplace", rel(res2, res + pre), 1e-5)
uu = torch.randn(M, N, device=dev).bfloat16()
out = ops.gemm(A, B, epilogue=ops.EPI_DGELU, aux_in=uu)
s = torch.sigmoid(1.702 * uu.float()); ref = acc * (s * (1 + 1.702 * uu.float() * (1 - s)))
report("epi dgelu", rel(out, ref))
P = 100; Bsz = 7
pos = torch.randn(P + 1, N, device=dev)
tok = torch.zeros(Bsz * (P + 1), N, device=dev)
ops.gemm(A, B, epilogue=ops.EPI_PATCH, aux_in=pos, out=tok, patch_P=P)
ref = torch.zeros(Bsz, P + 1, N, device=dev); ref[:, 1:] = acc.view(Bsz, P, N) + pos[1:]
report("epi patch", rel(tok, ref.view(-1, N)), 1e-5)
# split-K wgrad shape
Mtok = 257 * 9
dY = torch.randn(Mtok, 1024, device=dev).bfloat16(); X = torch.randn(Mtok, 1024, device=dev).bfloat16()
out = ops.gemm(dY, X, a_mn=True, b_mn=True, out_dtype=torch.float32)
report("wgrad split-k auto", rel(out, dY.float().t() @ X.float()), 1e-5)

# timing
def bench(M, N, K, a_mn=False, b_mn=False, bn=0, iters=20, **kw):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=kw.pop("dt", torch.bfloat16))
    for _ in range(3): ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out=out, force_bn=bn, **kw)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): ops.gemm(A, B, a_mn=a_mn, b_mn=b_mn, out=out, force_bn=bn, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ref_ms = None
    if not a_mn and not b_mn:
        for _ in range(3): torch.matmul(A, B.t())
        e0.record()
        for _ in range(iters): torch.matmul(A, B.t())
        e1.record(); torch.cuda.synchronize(); ref_ms = e0.elapsed_time(e1) / iters
    tf = 2 * M * N * K / ms / 1e9
    print(f"time M{M} N{N} K{K} a_mn={int(a_mn)} b_mn={int(b_mn)} bn={bn}: {ms:.3f} ms {tf:.0f} TF/s" + (f" (cublas {2*M*N*K/ref_ms/1e9:.0f})" if ref_ms else ""), flush=True)

Mt = 58 * 257
def bench_epi(name, **kw):
    M, N, K = Mt, 4096, 1024
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(A, B, out=out, **kw)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20): ops.gemm(A, B, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"time epi {name}: {ms:.3f} ms {2*M*N*K/ms/1e9:.0f} TF/s", flush=True)
ub = torch.randn(Mt, 4096, device=dev).bfloat16()
bench_epi("plain")
bench_epi("gelu", bias=torch.randn(4096, device=dev), epilogue=ops.EPI_GELU, aux_out=torch.empty_like(ub))
bench_epi("dgelu", epilogue=ops.EPI_DGELU, aux_in=ub)
for bn in (256,):
    bench(Mt, 3072, 1024, bn=bn); bench(Mt, 1024, 1024, bn=bn); bench(Mt, 4096, 1024, bn=bn); bench(Mt, 1024, 4096, bn=bn)
bench(8192, 8192, 8192, bn=256)
bench(Mt, 1024, 4096, b_mn=True)
bench(4096, 1024, Mt, a_mn=True, b_mn=True, dt=torch.float32)
bench(1024, 1024, Mt, a_mn=True, b_mn=True, dt=torch.float32)
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
