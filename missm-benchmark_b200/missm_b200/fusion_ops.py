"""Autograd wrapper of the fused masked-fusion CUDA kernels (csrc/fusion.cu) used by the default
`sum` head (reference: modal_sum.forward, src/model/baseline.py:52-61)."""
import ctypes

import torch

from ._lib import check, lib, stream_ptr

MAX_TOWERS = 8
F32 = torch.float32


class FusionSumArgs(ctypes.Structure):
    """Mirror of `missm_fusion_sum_args` (include/missm_b200.h)."""
    _fields_ = [
        ("emb", ctypes.c_void_p * MAX_TOWERS), ("weight", ctypes.c_void_p * MAX_TOWERS),
        ("bias", ctypes.c_void_p * MAX_TOWERS), ("d_emb", ctypes.c_void_p * MAX_TOWERS),
        ("d_weight", ctypes.c_void_p * MAX_TOWERS), ("d_bias", ctypes.c_void_p * MAX_TOWERS),
        ("codes", ctypes.c_int32 * MAX_TOWERS),
        ("missing_index", ctypes.c_void_p), ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p),
        ("pre", ctypes.c_void_p), ("out", ctypes.c_void_p), ("mean", ctypes.c_void_p), ("rstd", ctypes.c_void_p),
        ("n_modal", ctypes.c_int32), ("B", ctypes.c_int32), ("P", ctypes.c_int32), ("Fd", ctypes.c_int32),
        ("eps", ctypes.c_float),
    ]


def _fill(embs, ws, bs, codes, mi, gamma, beta, pre, out, mean, rstd, eps):
    a = FusionSumArgs()
    n = len(embs)
    for i in range(n):
        a.emb[i], a.weight[i], a.bias[i], a.codes[i] = embs[i].data_ptr(), ws[i].data_ptr(), bs[i].data_ptr(), codes[i]
    a.missing_index, a.gamma, a.beta = mi.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    a.pre, a.out, a.mean, a.rstd = pre.data_ptr(), out.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    a.n_modal, a.B, a.P, a.Fd, a.eps = n, embs[0].shape[0], embs[0].shape[1], ws[0].shape[0], eps
    return a


class _MaskedSumNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, n, codes, mi, eps, gamma, beta, *tensors):
        embs = [t.detach().contiguous() for t in tensors[:n]]
        ws = [t.detach().contiguous() for t in tensors[n:2 * n]]
        bs = [t.detach().contiguous() for t in tensors[2 * n:3 * n]]
        B, Fd, dev = embs[0].shape[0], ws[0].shape[0], embs[0].device
        pre = torch.empty((B, Fd), device=dev, dtype=F32)
        out = torch.empty((B, Fd), device=dev, dtype=F32)
        mean = torch.empty((B,), device=dev, dtype=F32)
        rstd = torch.empty((B,), device=dev, dtype=F32)
        a = _fill(embs, ws, bs, codes, mi, gamma.detach(), beta.detach(), pre, out, mean, rstd, eps)
        check(lib().missm_fusion_sum_fwd(ctypes.byref(a), stream_ptr()), "fusion_sum_fwd")
        ctx.n, ctx.codes, ctx.eps = n, codes, eps
        ctx.save_for_backward(mi, gamma, beta, pre, mean, rstd, *embs, *ws, *bs)
        return out

    @staticmethod
    def backward(ctx, d_out):
        n = ctx.n
        mi, gamma, beta, pre, mean, rstd = ctx.saved_tensors[:6]
        rest = ctx.saved_tensors[6:]
        embs, ws, bs = list(rest[:n]), list(rest[n:2 * n]), list(rest[2 * n:3 * n])
        B, Fd, dev = pre.shape[0], pre.shape[1], pre.device
        d_embs = [torch.empty_like(e) for e in embs]
        d_ws = [torch.empty_like(w) for w in ws]
        d_bs = [torch.empty_like(b) for b in bs]
        a = _fill(embs, ws, bs, ctx.codes, mi, gamma, beta, pre, pre, mean, rstd, ctx.eps)
        for i in range(n):
            a.d_emb[i], a.d_weight[i], a.d_bias[i] = d_embs[i].data_ptr(), d_ws[i].data_ptr(), d_bs[i].data_ptr()
        work = torch.empty((3, B, Fd), device=dev, dtype=F32)
        d_gamma = torch.empty((Fd,), device=dev, dtype=F32)
        d_beta = torch.empty((Fd,), device=dev, dtype=F32)
        d_out = d_out.contiguous()
        check(lib().missm_fusion_sum_bwd(ctypes.byref(a), ctypes.c_void_p(d_out.data_ptr()),
                                         ctypes.c_void_p(work.data_ptr()), ctypes.c_void_p(d_gamma.data_ptr()),
                                         ctypes.c_void_p(d_beta.data_ptr()), stream_ptr()), "fusion_sum_bwd")
        return (None, None, None, None, d_gamma, d_beta, *d_embs, *d_ws, *d_bs)


def masked_sum_norm(embs, weights, biases, codes, missing_index, gamma, beta, eps):
    """LayerNorm(sum_m present_m * Linear_m(emb_m)) -- fp32, one fused CUDA kernel each way."""
    n = len(embs)
    if n > MAX_TOWERS:
        raise ValueError(f"at most {MAX_TOWERS} modalities")
    for t in embs:
        if not t.is_cuda:
            raise RuntimeError("missm_b200: fusion inputs must be CUDA tensors (no CPU fallback)")
    mi = missing_index.reshape(-1).to(device=embs[0].device, dtype=torch.int64).contiguous()
    embs = [e.float() for e in embs]
    return _MaskedSumNorm.apply(n, [int(c) for c in codes], mi, float(eps), gamma, beta, *embs, *weights, *biases)
