"""Thin Python wrappers over the C ABI: shape/dtype validation happens here, in Python, and the
CUDA entry points receive raw device pointers + the current stream."""
import ctypes

import torch

from . import _lib
from ._lib import GemmArgs, check, lib, ptr, stream_ptr

EPI_LINEAR, EPI_GELU, EPI_RESID, EPI_DGELU, EPI_PATCH = range(5)


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, "expected a row-major 2-D tensor"
    return t.stride(0)


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=torch.bfloat16, bias=None,
         epilogue=EPI_LINEAR, aux_in=None, aux_out=None, scale_cols=0, col_scale=1.0,
         patch_P=0, out_rows=None, split_k=0, force_bn=0):
    """C[m,n] = epilogue(sum_k A[m,k] B[n,k]).  `a`: [M,K] (or [K,M] if a_mn), `b`: [N,K]
    (or [K,N] if b_mn); both bf16 CUDA tensors, row-major."""
    assert a.is_cuda and b.is_cuda and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    assert K == Kb, f"contraction mismatch {K} vs {Kb}"
    if out is None:
        out = torch.empty((out_rows if out_rows is not None else M, N), device=a.device,
                          dtype=out_dtype)
    assert out.dtype in (torch.bfloat16, torch.float32)
    g = GemmArgs()
    g.A, g.B, g.C = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.bias = bias.data_ptr() if bias is not None else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    g.aux_in = aux_in.data_ptr() if aux_in is not None else None
    g.aux_out = aux_out.data_ptr() if aux_out is not None else None
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldc = _ld(a), _ld(b), _ld(out)
    g.ld_aux_in = _ld(aux_in) if aux_in is not None else 0
    g.ld_aux_out = _ld(aux_out) if aux_out is not None else 0
    g.a_mn, g.b_mn = int(a_mn), int(b_mn)
    g.epilogue = epilogue
    g.out_f32 = int(out.dtype == torch.float32)
    g.scale_cols, g.col_scale = scale_cols, col_scale
    g.patch_P, g.split_k, g.force_bn = patch_P, split_k, force_bn
    check(lib().missm_gemm_bf16(ctypes.byref(g), stream_ptr()), "gemm_bf16")
    return out
