"""Patch embedding of one tower of the bench workload (58 present images, 224 x 224, D = 1024): implicit GEMM
(csrc/patch_embed_tc.cu) vs the explicit path (patchify + GEMM with the EPI_PATCH epilogue).  CUDA events, L2 flushed."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "missm-benchmark_b200"))
import torch
from missm_b200 import ops

dev = "cuda"
n, D, ps, Kpad, P = 58, 1024, 14, 640, 256
px = torch.randn(64, 3, 224, 224, device=dev)
idx = torch.randperm(64, device=dev)[:n].sort().values.int()
w = ops.cast_bf16(torch.randn(D, 588, device=dev) * 0.05, cols_dst=Kpad)
pos = torch.randn(P + 1, D, device=dev)
tok = torch.empty(n * (P + 1), D, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, it=10):
    fn()
    ts = []
    for _ in range(it):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)


def implicit():
    assert ops.patch_embed_implicit(px, w, pos, tok, ps, 1, sample_index=idx, n_samples=n)


def explicit():
    patches = ops.patchify(px, ps, Kpad, 1, sample_index=idx, n_samples=n)
    ops.gemm(patches, w, out=tok, epilogue=ops.EPI_PATCH, aux_in=pos, patch_P=P)


flop = 2 * n * P * D * 588
ti, te = timed(implicit), timed(explicit)
print(f"implicit {ti:.1f} us ({flop / ti / 1e6:.0f} TFLOP/s)   explicit (patchify + GEMM) {te:.1f} us")
