"""fp32 VERIFICATION mode of the hot path (MISSM_PRECISION=fp32 or autograd.set_precision('fp32')).

Same blocks, same kernels' contracts, fp32-grade arithmetic end to end, so that embeddings / loss /
gradients can be checked against the reference's fp32 PyTorch path at <= 1e-5 (BASELINE.json
north_star) instead of the bf16 mode's <= 1e-2.  Not a performance path:

* every Linear / wgrad / dgrad is ONE launch of the bf16 tcgen05 GEMM over 3-way bf16 splits of both fp32
  operands laid out along the contraction dimension (ops.gemm_f32, csrc/fp32_mode.cu) -- tcgen05 has no
  fp32 MMA;
* LayerNorm, pooling, embeddings and the fusion kernels are the production kernels (they compute in fp32
  already) asked for fp32 outputs;
* attention and QuickGELU run on the CUDA cores in fp32 (missm_attention_f32_*, missm_gelu_f32_*).

Reference arithmetic: CLIPEncoderLayer.forward languagebind/image/modeling_image.py:86-158 (video
:192-264), CLIPAttention / CLIPMLP of transformers 4.3x, languagebind/__init__.py:79-83.
"""
import torch

from . import ops
from .ops import F32, EPI_PATCH, EPI_RESID


def _contig(g):
    return g if g.is_contiguous() else g.contiguous()


def _cached(cache, name, params, build):
    ver = tuple((p._version, p.data_ptr()) for p in params)
    hit = cache.get(name)
    if hit is not None and hit[0] == ver:
        return hit[1]
    val = build()
    cache[name] = (ver, val)
    return val


def packed_qkv_f32(cache, qw, kw, vw, qb, kb, vb):
    def build():
        D = qw.shape[0]
        w = torch.empty((3 * D, qw.shape[1]), device=qw.device, dtype=F32)
        b = torch.empty((3 * D,), device=qw.device, dtype=F32)
        for i, (wi, bi) in enumerate(((qw, qb), (kw, kb), (vw, vb))):
            ops.copy_f32(wi.detach().contiguous(), w[i * D:(i + 1) * D])
            ops.copy_f32(bi.detach().contiguous(), b[i * D:(i + 1) * D])
        return w, b
    return _cached(cache, "qkv32", (qw, kw, vw, qb, kb, vb), build)


# x + OutProj(Attention(QKV(LN(x [+ temporal embedding]))))     (autograd.AttnBlockFn in fp32)
class AttnBlockF32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, meta, cache, ln_w, ln_b, qw, qb, kw, kb, vw, vb, ow, ob, temb):
        D = x.shape[1]
        hd = D // meta.H
        wqkv, bqkv = packed_qkv_f32(cache, qw, kw, vw, qb, kb, vb)
        wo = ow.detach()
        if temb is not None:
            x_res = torch.empty_like(x)
            h, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, meta.eps, out_dtype=F32,
                                              add_rows=temb.detach().reshape(-1, D), add_period=meta.add_period,
                                              add_div=meta.add_div, x_out=x_res)
        else:
            x_res = x
            h, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, meta.eps, out_dtype=F32)
        qkv = ops.gemm_f32(h, wqkv, bias=bqkv, scale_cols=D, col_scale=hd ** -0.5)
        attn, lse = ops.attention_f32_fwd(qkv, meta.layout, meta.H, causal=meta.causal, key_mask=meta.key_mask,
                                          mask_rows=meta.mask_rows, mask_div=meta.mask_div)
        out = ops.gemm_f32(attn, wo, bias=ob.detach(), epilogue=EPI_RESID, aux_in=x_res)
        ctx.meta = meta
        ctx.has_temb = temb is not None
        ctx.save_for_backward(x_res, mean, rstd, h, qkv, attn, lse, wqkv, wo, ln_w)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x_res, mean, rstd, h, qkv, attn, lse, wqkv, wo, ln_w = ctx.saved_tensors
        meta = ctx.meta
        D = x_res.shape[1]
        hd = D // meta.H
        d_out = _contig(d_out)
        d_ob = ops.colsum_f32(d_out)
        d_ow = ops.gemm_f32(d_out, attn, a_mn=True, b_mn=True)                          # dY^T @ attn
        d_attn = ops.gemm_f32(d_out, wo, b_mn=True)                                     # dY @ Wo
        dqkv = ops.attention_f32_bwd(qkv, attn, lse, d_attn, meta.layout, meta.H, hd ** -0.5, causal=meta.causal,
                                     key_mask=meta.key_mask, mask_rows=meta.mask_rows, mask_div=meta.mask_div)
        d_bqkv = ops.colsum_f32(dqkv)
        d_wqkv = ops.gemm_f32(dqkv, h, a_mn=True, b_mn=True)                            # [3D, D]
        d_h = ops.gemm_f32(dqkv, wqkv, b_mn=True)
        dx, _, d_lnw, d_lnb, _ = ops.layernorm_bwd(d_h, x_res, mean, rstd, ln_w, dres=d_out)
        d_temb = None
        if ctx.has_temb:
            d_temb = ops.colsum_grouped(dx, meta.add_period, meta.add_div).view(1, meta.add_period, D)
        return (dx, None, None, d_lnw, d_lnb, d_wqkv[:D], d_bqkv[:D], d_wqkv[D:2 * D], d_bqkv[D:2 * D],
                d_wqkv[2 * D:], d_bqkv[2 * D:], d_ow, d_ob, d_temb)


# x + fc2(quick_gelu(fc1(LN(x))))     (autograd.MlpBlockFn in fp32)
class MlpBlockF32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps, cache, ln_w, ln_b, w1, b1, w2, b2):
        w1d, w2d = w1.detach(), w2.detach()
        h, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, eps, out_dtype=F32)
        u = ops.gemm_f32(h, w1d, bias=b1.detach())
        a = ops.gelu_f32_fwd(u)
        out = ops.gemm_f32(a, w2d, bias=b2.detach(), epilogue=EPI_RESID, aux_in=x)
        ctx.save_for_backward(x, mean, rstd, h, u, a, w1d, w2d, ln_w)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, mean, rstd, h, u, a, w1d, w2d, ln_w = ctx.saved_tensors
        d_out = _contig(d_out)
        d_b2 = ops.colsum_f32(d_out)
        d_w2 = ops.gemm_f32(d_out, a, a_mn=True, b_mn=True)
        d_a = ops.gemm_f32(d_out, w2d, b_mn=True)
        d_u = ops.gelu_f32_bwd(d_a, u)
        d_w1 = ops.gemm_f32(d_u, h, a_mn=True, b_mn=True)
        d_b1 = ops.colsum_f32(d_u)
        d_h = ops.gemm_f32(d_u, w1d, b_mn=True)
        dx, _, d_lnw, d_lnb, _ = ops.layernorm_bwd(d_h, x, mean, rstd, ln_w, dres=d_out)
        return dx, None, None, d_lnw, d_lnb, d_w1, d_b1, d_w2, d_b2


# CLIPVisionEmbeddings + pre_layrnorm on the present samples     (autograd.VisionEmbedFn in fp32)
class VisionEmbedF32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pixels, present_idx, n_present, geom, cache, cls, patch_w, pos, ln_w, ln_b):
        ps, T, gh, gw, eps = geom
        D = patch_w.shape[0]
        K = patch_w.shape[1] * ps * ps
        Kpad = (K + 7) // 8 * 8
        P = gh * gw

        def build():
            w = torch.zeros((D, Kpad), device=patch_w.device, dtype=F32)
            w[:, :K] = patch_w.detach().reshape(D, K)
            return w
        wp = _cached(cache, "patch32", (patch_w,), build)
        patches = ops.patchify_f32(pixels, ps, Kpad, T, sample_index=present_idx, n_samples=n_present)
        n_img = n_present * T
        tok = torch.empty((n_img * (P + 1), D), device=pixels.device, dtype=F32)
        ops.gemm_f32(patches, wp, out=tok, epilogue=EPI_PATCH, aux_in=pos.detach(), patch_P=P)
        ops.cls_rows(cls.detach(), pos.detach(), tok, n_img, P + 1)
        x0, mean, rstd = ops.layernorm_fwd(tok, ln_w, ln_b, eps, out_dtype=F32)
        ctx.dims = (n_img, P, D, K, tuple(patch_w.shape))
        ctx.save_for_backward(tok, mean, rstd, patches, ln_w)
        return x0

    @staticmethod
    def backward(ctx, d_x0):
        tok, mean, rstd, patches, ln_w = ctx.saved_tensors
        n_img, P, D, K, wshape = ctx.dims
        d_tok, _, d_lnw, d_lnb, _ = ops.layernorm_bwd(_contig(d_x0), tok, mean, rstd, ln_w)
        d_pos, _ = ops.embed_bwd(d_tok, n_img, P + 1)
        # fp32 gradient rows of the conv output = the non-CLS rows of d_tok (index arithmetic only)
        r = torch.arange(n_img * P, device=d_tok.device, dtype=torch.int32)
        d_patch = ops.gather_rows(d_tok, r + r // P + 1, n_img * P)
        d_w = ops.gemm_f32(d_patch, patches, a_mn=True, b_mn=True)                      # [D, Kpad]
        d_w = d_w[:, :K].reshape(wshape)
        d_cls = d_pos[0].clone()
        return None, None, None, None, None, d_cls, d_w, d_pos, d_lnw, d_lnb


# pooled rows -> LayerNorm -> (frame mean) -> projection -> L2 normalise -> * exp(logit_scale)   (PoolProjFn)
class PoolProjF32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rows, n_present, T, eps, scale, cache, ln_w, ln_b, proj_w):
        wp = proj_w.detach()
        pooled, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, eps, out_dtype=F32, row_index=rows,
                                               n_rows=n_present * T)
        if T != 1:
            pooled = ops.frame_mean(pooled, n_present, T, out_dtype=F32)
        z = ops.gemm_f32(pooled, wp)
        y, inv = ops.l2norm_scale_fwd(z, scale)
        ctx.info = (n_present, T, scale)
        ctx.save_for_backward(x, rows, mean, rstd, pooled, z, inv, wp, ln_w)
        return y

    @staticmethod
    def backward(ctx, d_y):
        x, rows, mean, rstd, pooled, z, inv, wp, ln_w = ctx.saved_tensors
        n_present, T, scale = ctx.info
        d_z = ops.l2norm_scale_bwd(_contig(d_y), z, inv, scale, out_dtype=F32)
        d_proj = ops.gemm_f32(d_z, pooled, a_mn=True, b_mn=True, split_k=1)
        d_pooled = ops.gemm_f32(d_z, wp, b_mn=True)
        if T != 1:
            d_pooled = ops.frame_mean_bwd(d_pooled, n_present, T)
        dx = torch.zeros_like(x)
        _, _, d_lnw, d_lnb, _ = ops.layernorm_bwd(d_pooled, x, mean, rstd, ln_w, row_index=rows, dx=dx)
        return dx, None, None, None, None, None, None, d_lnw, d_lnb, d_proj
