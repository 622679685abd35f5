"""torch.autograd.Functions that drive the CUDA kernels (one Function per residual block so
that DDP's bucket hooks see parameter gradients layer by layer, last layer first).

Activations: fp32 residual stream [M, D]; bf16 GEMM operands; fp32 accumulation everywhere.
Every Function returns an explicit gradient (never None) for every parameter it receives --
DDP at train_ddp.py:189 runs with find_unused_parameters=False.
"""
import os

import torch
from torch.utils.weak import WeakIdKeyDictionary

from . import ops
from .ops import BF16, F32, EPI_DGELU, EPI_GELU, EPI_PATCH, EPI_RESID

# ------------------------------------------------------------------------------------------
# side channel: bf16 copy of a residual-stream gradient, produced by the LayerNorm-backward of
# block k and consumed by block k-1 (autograd itself only carries the fp32 tensor)
# ------------------------------------------------------------------------------------------
_GRAD_BF16 = {}


_AB_NO_LN_COLSUM = os.environ.get("MISSM_AB_NO_LN_COLSUM") is not None     # A/B measurement switches
_AB_DGELU_COLSUM = os.environ.get("MISSM_AB_DGELU_COLSUM") is not None


def _publish_bf16(grad_f32, grad_bf16, colsum=None):
    if _AB_NO_LN_COLSUM:
        colsum = None
    # holding grad_f32 keeps its storage alive, so a pointer match means "the same tensor"
    _GRAD_BF16[grad_f32.data_ptr()] = (grad_f32, grad_bf16, colsum)


def _bf16_of(grad_f32):
    """-> (bf16 copy, column sums [D]) of a residual-stream gradient; both come for free from the
    LayerNorm-backward kernel that produced it, else they are computed here."""
    hit = _GRAD_BF16.pop(grad_f32.data_ptr(), None)
    if hit is not None and hit[0].shape == grad_f32.shape and hit[0]._version == grad_f32._version:
        return hit[1], (hit[2] if hit[2] is not None else ops.colsum(hit[1]))
    b = ops.cast_bf16(grad_f32)
    return b, ops.colsum(b)


# ------------------------------------------------------------------------------------------
# precision mode: 'bf16' (the product path) or 'fp32' (verification mode, autograd_f32.py)
# ------------------------------------------------------------------------------------------
_PRECISION = [os.environ.get("MISSM_PRECISION", "bf16").lower()]


def set_precision(mode):
    """'bf16': tcgen05 bf16 operands / fp32 accumulate (default).  'fp32': the verification mode -- same path,
    fp32-grade arithmetic (3-way split GEMMs, fp32 attention), for <= 1e-5 parity checks.  Returns the old mode."""
    if mode not in ("bf16", "fp32"):
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {mode!r}")
    old, _PRECISION[0] = _PRECISION[0], mode
    return old


def get_precision():
    if _PRECISION[0] not in ("bf16", "fp32"):
        raise ValueError(f"MISSM_PRECISION must be 'bf16' or 'fp32', got {_PRECISION[0]!r}")
    return _PRECISION[0]


def reset_side_channel():
    _GRAD_BF16.clear()


def _contig(g):
    return g if g.is_contiguous() else g.contiguous()


# ------------------------------------------------------------------------------------------
# bf16 operand copies of fp32 master weights, cached on the owning module by parameter version
# ------------------------------------------------------------------------------------------
def _versions(params):
    return tuple((p._version, p.data_ptr()) for p in params)


def cached_weight(cache, name, params, build):
    ver = _versions(params)
    hit = cache.get(name)
    if hit is not None and hit[0] == ver:
        return hit[1]
    val = build()
    cache[name] = (ver, val, tuple(params))
    return val


def _w2d(p):
    return p.detach().reshape(p.shape[0], -1)


# parameter -> (cache dict, entry name, slot) of the bf16 GEMM-operand copy made from it, for optimizers that can
# write that copy themselves while they update the parameter (optim.FusedAdam: `bf16_out` of missm_adam_multi)
_SINKS = WeakIdKeyDictionary()         # keyed by identity: Tensor.__eq__ is elementwise


def bf16_weight(cache, name, p, cols_dst=None):
    val = cached_weight(cache, name, (p,), lambda: ops.cast_bf16(_w2d(p), cols_dst=cols_dst))
    if cols_dst is None:
        _SINKS[p] = (cache, name, None)
    return val


def packed_qkv(cache, qw, kw, vw, qb, kb, vb):
    D = qw.shape[0]

    def build_w():
        w = torch.empty((3 * D, qw.shape[1]), device=qw.device, dtype=BF16)
        for i, wi in enumerate((qw, kw, vw)):
            ops.cast_bf16(wi.detach(), out=w[i * D:(i + 1) * D])
        return w

    def build_b():
        b = torch.empty((3 * D,), device=qw.device, dtype=F32)
        for i, bi in enumerate((qb, kb, vb)):
            ops.copy_f32(bi.detach(), b[i * D:(i + 1) * D])
        return b

    w = cached_weight(cache, "qkv_w", (qw, kw, vw), build_w)
    b = cached_weight(cache, "qkv_b", (qb, kb, vb), build_b)
    for i, wi in enumerate((qw, kw, vw)):
        _SINKS[wi] = (cache, "qkv_w", i)
    return w, b


def plan_operand_refresh(touched):
    """For an optimizer about to update the parameters `touched` through raw pointers: which of them have a bf16
    operand copy it may rewrite in the same pass.  -> ({id(param): bf16 destination tensor}, entries to `restamp`
    after the update).  An entry qualifies only if it is FRESH now and every parameter it was built from is either
    untouched or has this entry as its sink -- then rewriting the touched slices keeps it exact; anything else is
    left alone and rebuilds itself at the next forward through the ordinary version check."""
    touched_ids = {id(p) for p in touched}
    by_entry = {}
    for p in touched:
        sink = _SINKS.get(p)
        if sink is not None:
            by_entry.setdefault((id(sink[0]), sink[1]), (sink[0], sink[1], []))[2].append((p, sink[2]))
    dst, entries = {}, []
    for cache, name, plist in by_entry.values():
        hit = cache.get(name)
        if hit is None or hit[0] != _versions(hit[2]):
            continue                                           # absent or already stale
        mine = {id(p) for p, _ in plist}
        if any(id(q) in touched_ids and id(q) not in mine for q in hit[2]):
            continue
        for p, slot in plist:
            val = hit[1]
            if slot is not None:
                rows = p.shape[0]
                val = val[slot * rows:(slot + 1) * rows]
            if val.numel() != p.numel() or not val.is_contiguous():
                break
            dst[id(p)] = val
        else:
            entries.append((cache, name))
            continue
        for p, _ in plist:                                     # a slice did not line up: drop the whole entry
            dst.pop(id(p), None)
    return dst, entries


def restamp(entries):
    """Mark the entries of plan_operand_refresh current again (their parameters' versions have advanced and the
    optimizer rewrote their bf16 slices from the updated values)."""
    for cache, name in entries:
        hit = cache.get(name)
        if hit is not None:
            cache[name] = (_versions(hit[2]), hit[1], hit[2])


class AttnMeta:
    """Static description of one attention block."""

    def __init__(self, H, eps, layout, causal=False, key_mask=None, mask_rows=None, mask_div=1,
                 add_period=0, add_div=0):
        self.H, self.eps, self.layout = H, eps, layout
        self.causal, self.key_mask, self.mask_rows, self.mask_div = causal, key_mask, mask_rows, mask_div
        self.add_period, self.add_div = add_period, add_div


# ------------------------------------------------------------------------------------------
# x + OutProj(Attention(QKV(LN(x [+ temporal embedding]))))
#   reference: CLIPEncoderLayer.forward, languagebind/image/modeling_image.py:105-127 (temporal)
#   and :137-146 (spatial); CLIPAttention = transformers 4.3x
# ------------------------------------------------------------------------------------------
class AttnBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, meta, cache, ln_w, ln_b, qw, qb, kw, kb, vw, vb, ow, ob, temb):
        D = x.shape[1]
        hd = D // meta.H
        wqkv, bqkv = packed_qkv(cache, qw, kw, vw, qb, kb, vb)
        wo = bf16_weight(cache, "o", ow)
        if temb is not None:
            # hidden_states + temporal_embedding[:, :t]  (modeling_image.py:110-114); out of place
            x_res = torch.empty_like(x)
            h, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, meta.eps, add_rows=temb.detach().reshape(-1, D),
                                              add_period=meta.add_period, add_div=meta.add_div, x_out=x_res)
        else:
            x_res = x
            h, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, meta.eps)
        qkv = ops.gemm(h, wqkv, bias=bqkv, scale_cols=D, col_scale=hd ** -0.5)
        attn, lse = ops.attention_fwd(qkv, meta.layout, meta.H, causal=meta.causal, key_mask=meta.key_mask,
                                      mask_rows=meta.mask_rows, mask_div=meta.mask_div)
        out = ops.gemm(attn, wo, bias=ob.detach(), epilogue=EPI_RESID, aux_in=x_res, out_dtype=F32)
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
        d_out_b, d_ob = _bf16_of(d_out)
        # a frozen encoder (peft freezes everything but the adapters, modeling_image.py:793) asks for no weight
        # gradients: dgrad only
        wgrads = any(ctx.needs_input_grad[5:13])
        d_ow = ops.gemm(d_out_b, attn, a_mn=True, b_mn=True, out_dtype=F32) if wgrads else None   # dY^T @ attn
        d_attn = ops.gemm(d_out_b, wo, b_mn=True)                                      # dY @ Wo
        dqkv, d_bqkv = ops.attention_bwd(qkv, attn, lse, d_attn, meta.layout, meta.H, hd ** -0.5, causal=meta.causal,
                                         key_mask=meta.key_mask, mask_rows=meta.mask_rows, mask_div=meta.mask_div,
                                         want_colsum=wgrads)
        d_wqkv = ops.gemm(dqkv, h, a_mn=True, b_mn=True, out_dtype=F32) if wgrads else None      # [3D, D]
        d_h = ops.gemm(dqkv, wqkv, b_mn=True)
        dx, dx_b, d_lnw, d_lnb, dx_cs = ops.layernorm_bwd(d_h, x_res, mean, rstd, ln_w, dres=d_out, want_bf16=True)
        _publish_bf16(dx, dx_b, dx_cs)
        d_temb = None
        if ctx.has_temb:
            d_temb = ops.colsum_grouped(dx, meta.add_period, meta.add_div).view(1, meta.add_period, D)
        if not wgrads:
            return (dx, None, None, d_lnw, d_lnb) + (None,) * 8 + (d_temb,)
        return (dx, None, None, d_lnw, d_lnb, d_wqkv[:D], d_bqkv[:D], d_wqkv[D:2 * D], d_bqkv[D:2 * D],
                d_wqkv[2 * D:], d_bqkv[2 * D:], d_ow, d_ob, d_temb)


# ------------------------------------------------------------------------------------------
# The same block with peft LoRA adapters on q / k / v / out_proj (reference: convert_to_lora,
# modeling_image.py:775-793; peft's Linear: y = W x + b + (alpha / r) B(A(x))).
#
# Adapters ride on the tcgen05 GEMMs through the CONTRACTION dimension instead of being merged into W (a bf16
# copy of W + sBA would round away a delta that is ~2^-8 of W early in training) or run as separate rank-r GEMM
# chains: with T = X A_cat^T (one skinny GEMM, r_pad = 8 columns per group) stored NEXT to X in one row-major
# buffer [X | T], the layer is ONE GEMM over K' = K + r_pad against [W | sB]; in the backward [dY | dY sB] against
# the row-stacked [W ; A_cat] gives dX in ONE GEMM, and dA_cat = (dY sB)^T X, d(sB) = dY^T T are two skinny
# wgrads.  Every operand is a column view of a wider-pitched buffer; nothing is copied or transposed.
# The encoder's own weights are frozen by peft, so their wgrads / bias sums are skipped (needs_input_grad).
# ------------------------------------------------------------------------------------------
def _pad8(n):
    return (n + 7) // 8 * 8


def lora_packs(cache, base, adapters, scaling):
    """bf16 operand buffers of one LoRA attention block.  Parameter-space work (a few [D, r] slices per layer, done
    with torch indexing / casts, re-done only when a parameter changes): the frozen base part is written once,
    the adapter columns / rows are refreshed in place after an optimizer step."""
    qw, qb, kw, kb, vw, vb, ow, ob = base
    qA, qB, kA, kB, vA, vB, oA, oB = adapters
    D, r = qw.shape[0], qA.shape[0]
    R3, R1 = _pad8(3 * r), _pad8(r)

    def build_base():
        dev = qw.device
        wf_qkv = torch.zeros((3 * D, D + R3), device=dev, dtype=BF16)       # [W | sB]   fwd operand / dT operand
        wb_qkv = torch.zeros((3 * D + R3, D), device=dev, dtype=BF16)       # [W ; A]    dgrad operand (MN-major)
        wf_o = torch.zeros((D, D + R1), device=dev, dtype=BF16)
        wb_o = torch.zeros((D + R1, D), device=dev, dtype=BF16)
        for i, w in enumerate((qw, kw, vw)):
            wb_qkv[i * D:(i + 1) * D] = w.detach()
        wf_qkv[:, :D] = wb_qkv[:3 * D]
        wb_o[:D] = ow.detach()
        wf_o[:, :D] = wb_o[:D]
        bqkv = torch.cat([qb.detach(), kb.detach(), vb.detach()]).float().contiguous()
        return wf_qkv, wb_qkv, wf_o, wb_o, bqkv

    packs = cached_weight(cache, "lora_base", base, build_base)
    wf_qkv, wb_qkv, wf_o, wb_o, bqkv = packs

    def fill_adapters():
        with torch.no_grad():
            for i, (A, B) in enumerate(((qA, qB), (kA, kB), (vA, vB))):
                wf_qkv[i * D:(i + 1) * D, D + i * r:D + (i + 1) * r] = B.detach() * scaling
                wb_qkv[3 * D + i * r:3 * D + (i + 1) * r] = A.detach()
            wf_o[:, D:D + r] = oB.detach() * scaling
            wb_o[D:D + r] = oA.detach()
        return True

    ver = tuple((p._version, p.data_ptr()) for p in adapters) + (wf_qkv.data_ptr(),)
    if cache.get("lora_adapters") != ver:
        fill_adapters()
        cache["lora_adapters"] = ver
    return packs


class LoraAttnBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, meta, cache, ln_w, ln_b, qw, qb, kw, kb, vw, vb, ow, ob, temb,
                qA, qB, kA, kB, vA, vB, oA, oB, scaling):
        M, D = x.shape
        hd = D // meta.H
        r = qA.shape[0]
        R3, R1 = _pad8(3 * r), _pad8(r)
        wf_qkv, wb_qkv, wf_o, wb_o, bqkv = lora_packs(cache, (qw, qb, kw, kb, vw, vb, ow, ob),
                                                      (qA, qB, kA, kB, vA, vB, oA, oB), scaling)
        hcat = torch.empty((M, D + R3), device=x.device, dtype=BF16)        # [LN(x) | LN(x) A_cat^T]
        h = hcat[:, :D]
        if temb is not None:
            x_res = torch.empty_like(x)
            _, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, meta.eps, add_rows=temb.detach().reshape(-1, D),
                                              add_period=meta.add_period, add_div=meta.add_div, x_out=x_res, out=h)
        else:
            x_res = x
            _, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, meta.eps, out=h)
        ops.gemm(h, wb_qkv[3 * D:], out=hcat[:, D:])                         # T = h A_cat^T          [M, R3]
        qkvcat = torch.empty((M, 3 * D + R3), device=x.device, dtype=BF16)   # pitch shared with [dqkv | dT]
        qkv = qkvcat[:, :3 * D]
        ops.gemm(hcat, wf_qkv, bias=bqkv, scale_cols=D, col_scale=hd ** -0.5, out=qkv)
        attncat = torch.empty((M, D + R1), device=x.device, dtype=BF16)      # [attn | attn A_o^T]
        attn = attncat[:, :D]
        _, lse = ops.attention_fwd(qkv, meta.layout, meta.H, causal=meta.causal, key_mask=meta.key_mask,
                                   mask_rows=meta.mask_rows, mask_div=meta.mask_div, out=attn)
        ops.gemm(attn, wb_o[D:], out=attncat[:, D:])
        out = ops.gemm(attncat, wf_o, bias=ob.detach(), epilogue=EPI_RESID, aux_in=x_res, out_dtype=F32)
        ctx.meta, ctx.has_temb, ctx.r, ctx.scaling = meta, temb is not None, r, scaling
        ctx.save_for_backward(x_res, mean, rstd, hcat, qkvcat, attncat, lse, wf_qkv, wb_qkv, wf_o, wb_o, ln_w)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x_res, mean, rstd, hcat, qkvcat, attncat, lse, wf_qkv, wb_qkv, wf_o, wb_o, ln_w = ctx.saved_tensors
        meta, r, s = ctx.meta, ctx.r, ctx.scaling
        M, D = x_res.shape
        hd = D // meta.H
        R3, R1 = _pad8(3 * r), _pad8(r)
        need = ctx.needs_input_grad
        base_grads = any(need[5:13])                       # somebody un-froze the encoder's own weights
        h, qkv, attn = hcat[:, :D], qkvcat[:, :3 * D], attncat[:, :D]
        d_out = _contig(d_out)
        _GRAD_BF16.pop(d_out.data_ptr(), None)
        # ---- out_proj group: [dY | dY sB_o] ----
        dycat = ops.cast_bf16(d_out, cols_dst=D + R1)
        dy = dycat[:, :D]
        ops.gemm(dy, wf_o[:, D:], b_mn=True, out=dycat[:, D:])                               # dT_o = dY (sB_o)
        d_attncat = torch.empty((M, D + R1), device=d_out.device, dtype=BF16)                # pitch of attn
        d_attn = d_attncat[:, :D]
        ops.gemm(dycat, wb_o, b_mn=True, out=d_attn)                                         # dY W_o + dT_o A_o
        d_oA = ops.gemm(dycat[:, D:], attn, a_mn=True, b_mn=True, out_dtype=F32)[:r]         # dT_o^T attn
        d_oB = ops.gemm(dy, attncat[:, D:], a_mn=True, b_mn=True, out_dtype=F32)[:, :r] * s  # dY^T T_o
        d_ow = d_ob = None
        if base_grads:
            d_ow = ops.gemm(dy, attn, a_mn=True, b_mn=True, out_dtype=F32)
            d_ob = ops.colsum(dy)
        # ---- attention core ----
        dqkvcat = torch.empty((M, 3 * D + R3), device=d_out.device, dtype=BF16)
        dqkv = dqkvcat[:, :3 * D]
        _, d_bqkv = ops.attention_bwd(qkv, attn, lse, d_attn, meta.layout, meta.H, hd ** -0.5, causal=meta.causal,
                                      key_mask=meta.key_mask, mask_rows=meta.mask_rows, mask_div=meta.mask_div,
                                      dqkv_out=dqkv, want_colsum=base_grads)
        # ---- q / k / v group: [dqkv | dqkv sB_cat] ----
        ops.gemm(dqkv, wf_qkv[:, D:], b_mn=True, out=dqkvcat[:, 3 * D:])                     # dT      [M, R3]
        d_h = ops.gemm(dqkvcat, wb_qkv, b_mn=True)                                           # dqkv W + dT A_cat
        d_acat = ops.gemm(dqkvcat[:, 3 * D:], h, a_mn=True, b_mn=True, out_dtype=F32)        # [R3, D]
        d_sb = ops.gemm(dqkv, hcat[:, D:], a_mn=True, b_mn=True, out_dtype=F32)              # [3D, R3]
        d_A = [d_acat[i * r:(i + 1) * r] for i in range(3)]
        d_B = [d_sb[i * D:(i + 1) * D, i * r:(i + 1) * r] * s for i in range(3)]
        d_w = [None] * 3
        d_b = [None] * 3
        if base_grads:
            d_wqkv = ops.gemm(dqkv, h, a_mn=True, b_mn=True, out_dtype=F32)
            d_w = [d_wqkv[i * D:(i + 1) * D] for i in range(3)]
            d_b = [d_bqkv[i * D:(i + 1) * D] for i in range(3)]
        dx, dx_b, d_lnw, d_lnb, dx_cs = ops.layernorm_bwd(d_h, x_res, mean, rstd, ln_w, dres=d_out, want_bf16=True)
        _publish_bf16(dx, dx_b, dx_cs)
        d_temb = None
        if ctx.has_temb:
            d_temb = ops.colsum_grouped(dx, meta.add_period, meta.add_div).view(1, meta.add_period, D)
        return (dx, None, None, d_lnw, d_lnb, d_w[0], d_b[0], d_w[1], d_b[1], d_w[2], d_b[2], d_ow, d_ob, d_temb,
                d_A[0], d_B[0], d_A[1], d_B[1], d_A[2], d_B[2], d_oA, d_oB, None)


# ------------------------------------------------------------------------------------------
# x + fc2(quick_gelu(fc1(LN(x))))      reference: modeling_image.py:129-134, :148-151; CLIPMLP
# ------------------------------------------------------------------------------------------
class MlpBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps, cache, ln_w, ln_b, w1, b1, w2, b2):
        w1b = bf16_weight(cache, "fc1", w1)
        w2b = bf16_weight(cache, "fc2", w2)
        h, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, eps)
        u = torch.empty((x.shape[0], w1.shape[0]), device=x.device, dtype=BF16)
        a = ops.gemm(h, w1b, bias=b1.detach(), epilogue=EPI_GELU, aux_out=u)
        out = ops.gemm(a, w2b, bias=b2.detach(), epilogue=EPI_RESID, aux_in=x, out_dtype=F32)
        ctx.save_for_backward(x, mean, rstd, h, u, a, w1b, w2b, ln_w)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, mean, rstd, h, u, a, w1b, w2b, ln_w = ctx.saved_tensors
        d_out = _contig(d_out)
        d_out_b, d_b2 = _bf16_of(d_out)
        if not any(ctx.needs_input_grad[5:9]):             # frozen MLP (peft-wrapped encoder): dgrad only
            d_u = ops.gemm(d_out_b, w2b, b_mn=True, epilogue=EPI_DGELU, aux_in=u)
            d_h = ops.gemm(d_u, w1b, b_mn=True)
            dx, dx_b, d_lnw, d_lnb, dx_cs = ops.layernorm_bwd(d_h, x, mean, rstd, ln_w, dres=d_out, want_bf16=True)
            _publish_bf16(dx, dx_b, dx_cs)
            return dx, None, None, d_lnw, d_lnb, None, None, None, None
        d_w2 = ops.gemm(d_out_b, a, a_mn=True, b_mn=True, out_dtype=F32)
        if _AB_DGELU_COLSUM:
            d_b1 = torch.zeros((u.shape[1],), device=u.device, dtype=F32)
            d_u = ops.gemm(d_out_b, w2b, b_mn=True, epilogue=EPI_DGELU, aux_in=u, colsum_out=d_b1)
        else:
            d_u = ops.gemm(d_out_b, w2b, b_mn=True, epilogue=EPI_DGELU, aux_in=u)     # (dY @ W2) * gelu'(u)
        d_w1 = ops.gemm(d_u, h, a_mn=True, b_mn=True, out_dtype=F32)
        # (the GEMM can also emit this column sum -- colsum_out -- but the dGELU epilogue is already the
        #  bound of that kernel: measured 3 ms / step slower than this separate HBM-bound pass)
        if not _AB_DGELU_COLSUM:
            d_b1 = ops.colsum(d_u)
        d_h = ops.gemm(d_u, w1b, b_mn=True)
        dx, dx_b, d_lnw, d_lnb, dx_cs = ops.layernorm_bwd(d_h, x, mean, rstd, ln_w, dres=d_out, want_bf16=True)
        _publish_bf16(dx, dx_b, dx_cs)
        return dx, None, None, d_lnw, d_lnb, d_w1, d_b1, d_w2, d_b2


# ------------------------------------------------------------------------------------------
# CLIPVisionEmbeddings + pre_layrnorm on the PRESENT samples only
#   reference: video/modeling_video.py:42-51 (conv k = s = patch, CLS, + position), the
#   5-D -> (b t) reshape of modeling_image.py:636-639 and pre_layrnorm at :649
# ------------------------------------------------------------------------------------------
class VisionEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pixels, present_idx, n_present, geom, cache, cls, patch_w, pos, ln_w, ln_b):
        ps, T, gh, gw, eps = geom
        D = patch_w.shape[0]
        K = patch_w.shape[1] * ps * ps
        Kpad = (K + 7) // 8 * 8
        P = gh * gw
        wp = bf16_weight(cache, "patch", patch_w, cols_dst=Kpad)
        patches = ops.patchify(pixels, ps, Kpad, T, sample_index=present_idx, n_samples=n_present)
        n_img = n_present * T
        tok = torch.empty((n_img * (P + 1), D), device=pixels.device, dtype=F32)
        ops.gemm(patches, wp, out=tok, epilogue=EPI_PATCH, aux_in=pos.detach(), patch_P=P)
        ops.cls_rows(cls.detach(), pos.detach(), tok, n_img, P + 1)
        x0, mean, rstd = ops.layernorm_fwd(tok, ln_w, ln_b, eps, out_dtype=F32)
        ctx.dims = (n_img, P, D, K, tuple(patch_w.shape))
        ctx.save_for_backward(tok, mean, rstd, patches, ln_w)
        return x0

    @staticmethod
    def backward(ctx, d_x0):
        tok, mean, rstd, patches, ln_w = ctx.saved_tensors
        n_img, P, D, K, wshape = ctx.dims
        d_x0 = _contig(d_x0)
        _GRAD_BF16.pop(d_x0.data_ptr(), None)
        d_tok, _, d_lnw, d_lnb, _ = ops.layernorm_bwd(d_x0, tok, mean, rstd, ln_w)
        d_pos, d_patch = ops.embed_bwd(d_tok, n_img, P + 1)
        d_w = ops.gemm(d_patch, patches, a_mn=True, b_mn=True, out_dtype=F32)          # [D, Kpad]
        d_w = d_w[:, :K].reshape(wshape)
        d_cls = d_pos[0].clone()
        return None, None, None, None, None, d_cls, d_w, d_pos, d_lnw, d_lnb


# ------------------------------------------------------------------------------------------
# CLIPTextEmbeddings on the present samples   (modeling_image.py:463,494)
# ------------------------------------------------------------------------------------------
class TextEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, present_idx, n_present, tok_emb, pos_emb):
        L = ids.shape[1]
        x0 = ops.text_embed_fwd(ids, tok_emb.detach(), pos_emb.detach()[:L], sample_index=present_idx,
                                n_samples=n_present)
        ctx.save_for_backward(ids, present_idx if present_idx is not None else ids.new_empty(0))
        ctx.info = (n_present, tok_emb.shape[0], pos_emb.shape[0], present_idx is not None)
        return x0

    @staticmethod
    def backward(ctx, d_x0):
        ids, present_idx = ctx.saved_tensors
        n_present, vocab, n_pos, has_idx = ctx.info
        d_x0 = _contig(d_x0)
        _GRAD_BF16.pop(d_x0.data_ptr(), None)
        d_tok, d_pos_used = ops.text_embed_bwd(ids, d_x0, vocab, sample_index=present_idx if has_idx else None,
                                               n_samples=n_present)
        if d_pos_used.shape[0] != n_pos:
            d_pos = torch.zeros((n_pos, d_pos_used.shape[1]), device=d_x0.device, dtype=F32)
            d_pos[:d_pos_used.shape[0]] = d_pos_used
        else:
            d_pos = d_pos_used
        return None, None, None, d_tok, d_pos


# ------------------------------------------------------------------------------------------
# pooled rows -> LayerNorm -> (frame mean) -> projection -> L2 normalise -> * exp(logit_scale)
#   reference: modeling_image.py:658-662 (CLS, post_layernorm, mean over T), :514-522 (text:
#   final_layer_norm + EOT row), languagebind/__init__.py:79-83
# ------------------------------------------------------------------------------------------
class PoolProjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rows, n_present, T, eps, scale, cache, ln_w, ln_b, proj_w):
        wp = bf16_weight(cache, "proj", proj_w)
        n_rows = n_present * T
        if T == 1:
            pooled_b, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, eps, row_index=rows, n_rows=n_rows)
        else:
            pooled_f, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, eps, out_dtype=F32, row_index=rows,
                                                     n_rows=n_rows)
            pooled_b = ops.frame_mean(pooled_f, n_present, T)
        z = ops.gemm(pooled_b, wp, out_dtype=F32)
        y, inv = ops.l2norm_scale_fwd(z, scale)
        ctx.info = (n_present, T, scale)
        ctx.save_for_backward(x, rows, mean, rstd, pooled_b, z, inv, wp, ln_w)
        return y

    @staticmethod
    def backward(ctx, d_y):
        x, rows, mean, rstd, pooled_b, z, inv, wp, ln_w = ctx.saved_tensors
        n_present, T, scale = ctx.info
        d_z = ops.l2norm_scale_bwd(_contig(d_y), z, inv, scale)                        # bf16 [Bp, P]
        d_proj = ops.gemm(d_z, pooled_b, a_mn=True, b_mn=True, out_dtype=F32, split_k=1)
        if T == 1:
            d_pooled = ops.gemm(d_z, wp, b_mn=True)                                    # bf16
        else:
            d_pm = ops.gemm(d_z, wp, b_mn=True, out_dtype=F32)
            d_pooled = ops.frame_mean_bwd(d_pm, n_present, T)                          # f32 [Bp*T, D]
        dx = torch.zeros_like(x)
        dx_b = torch.zeros(x.shape, device=x.device, dtype=BF16)
        _, _, d_lnw, d_lnb, dx_cs = ops.layernorm_bwd(d_pooled, x, mean, rstd, ln_w, row_index=rows, dx=dx,
                                                       dx_bf16=dx_b)
        _publish_bf16(dx, dx_b, dx_cs)
        return dx, None, None, None, None, None, None, d_lnw, d_lnb, d_proj


# ------------------------------------------------------------------------------------------
# embeddings of the present samples back into batch order, zero rows for missing samples
# ------------------------------------------------------------------------------------------
class ScatterZeroFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb_present, slot_of, present_idx, n_present, B):
        ctx.n = n_present
        ctx.save_for_backward(present_idx)
        return ops.scatter_rows_zero(emb_present, slot_of, B)

    @staticmethod
    def backward(ctx, d_out):
        (present_idx,) = ctx.saved_tensors
        return ops.gather_rows(_contig(d_out), present_idx, ctx.n), None, None, None, None


# ------------------------------------------------------------------------------------------
# data parallelism: the first backward node of a tower shrinks the persistent grids, so that the
# all-reduce kernels DDP launches while the backward is still running find free SMs
# ------------------------------------------------------------------------------------------
class BackwardSmsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n_sms):
        ctx.n_sms = n_sms
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ops.set_persistent_sms(ctx.n_sms)
        return g, None


# ------------------------------------------------------------------------------------------
# block entry points used by the module trees: dispatch on the precision mode
# ------------------------------------------------------------------------------------------
def _pick(bf16_fn, f32_name):
    if get_precision() == "fp32":
        from . import autograd_f32
        return getattr(autograd_f32, f32_name)
    return bf16_fn


def attn_block(*args):
    return _pick(AttnBlockFn, "AttnBlockF32Fn").apply(*args)


def lora_attn_block(x, meta, cache, ln_w, ln_b, qw, qb, kw, kb, vw, vb, ow, ob, temb,
                    qA, qB, kA, kB, vA, vB, oA, oB, scaling):
    if get_precision() == "fp32":
        # verification mode: the adapter folded into an effective fp32 weight W + s B A (exact in fp32; torch
        # autograd carries dW_eff back to A and B -- [D, r] parameter-space products, not a performance path)
        from . import autograd_f32
        eff = [w + scaling * (B @ A) for w, A, B in ((qw, qA, qB), (kw, kA, kB), (vw, vA, vB), (ow, oA, oB))]
        return autograd_f32.AttnBlockF32Fn.apply(x, meta, {}, ln_w, ln_b, eff[0], qb, eff[1], kb, eff[2], vb,
                                                 eff[3], ob, temb)
    return LoraAttnBlockFn.apply(x, meta, cache, ln_w, ln_b, qw, qb, kw, kb, vw, vb, ow, ob, temb,
                                 qA, qB, kA, kB, vA, vB, oA, oB, scaling)


def mlp_block(*args):
    return _pick(MlpBlockFn, "MlpBlockF32Fn").apply(*args)


def vision_embed(*args):
    return _pick(VisionEmbedFn, "VisionEmbedF32Fn").apply(*args)


def pool_proj(*args):
    return _pick(PoolProjFn, "PoolProjF32Fn").apply(*args)
