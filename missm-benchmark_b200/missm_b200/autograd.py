"""torch.autograd.Functions that drive the CUDA kernels (one Function per residual block so
that DDP's bucket hooks see parameter gradients layer by layer, last layer first).

Activations: fp32 residual stream [M, D]; bf16 GEMM operands; fp32 accumulation everywhere.
Every Function returns an explicit gradient (never None) for every parameter it receives --
DDP at train_ddp.py:189 runs with find_unused_parameters=False.
"""
import os

import torch
from torch.utils.weak import WeakIdKeyDictionary

from . import blocks, ops
from .ops import BF16, F32, EPI_DGELU, EPI_GELU, EPI_PATCH, EPI_RESID

# ------------------------------------------------------------------------------------------
# side channel: the bf16 copy of a residual-stream gradient and its column sums (= the bias gradient of the Linear
# that produced the stream), which the LayerNorm-backward of block k emits for free and block k-1 consumes.
# autograd itself only carries the fp32 tensor, so the pair rides on that tensor OBJECT as an attribute (torch
# preserves a tensor's Python object, attributes included, while the engine holds it) -- no process-global state,
# nothing to reset, safe with several models / re-entrant backwards.  A gradient that was accumulated, copied or
# modified on the way arrives without (or with a stale) attribute and the consumer recomputes what it needs.
# ------------------------------------------------------------------------------------------
_AB_NO_LN_COLSUM = os.environ.get("MISSM_AB_NO_LN_COLSUM") is not None     # A/B measurement switch


def _give_side(grad_f32, grad_bf16, colsum=None):
    if _AB_NO_LN_COLSUM:
        colsum = None
    grad_f32._missm_side = (grad_bf16, colsum, grad_f32._version)


def _take_side(grad_f32):
    """-> (bf16 copy | None, column sums [D] | None) riding on this gradient tensor."""
    hit = grad_f32.__dict__.pop("_missm_side", None)
    if hit is not None and hit[2] == grad_f32._version and hit[0].shape == grad_f32.shape:
        return hit[0], hit[1]
    return None, None


def _bf16_of(grad_f32):
    """-> (bf16 copy, column sums [D]) of a residual-stream gradient, from the side channel or computed here."""
    b, cs = _take_side(grad_f32)
    if b is None:
        b = ops.cast_bf16(grad_f32)
    return b, (cs if cs is not None else ops.colsum(b))


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
# One encoder layer = a chain of residual blocks, each issued by ONE driver call (blocks.py / csrc/blocks.cu):
#   attention block  x + OutProj(Attention(QKV(LN(x [+ temporal embedding]))))
#                    reference: CLIPEncoderLayer.forward, languagebind/image/modeling_image.py:105-127 (temporal)
#                    and :137-146 (spatial); CLIPAttention = transformers 4.3x
#   MLP block        x + fc2(quick_gelu(fc1(LN(x))))      modeling_image.py:129-134, :148-151; CLIPMLP
# LoRA (peft Linear y = W x + b + (alpha / r) B(A(x)); reference convert_to_lora, modeling_image.py:775-793): the
# adapters ride on the tcgen05 GEMMs through the CONTRACTION dimension instead of being merged into W (a bf16 copy
# of W + sBA would round away a delta that is ~2^-8 of W early in training) or run as separate rank-r GEMM chains:
# with T = X A_cat^T (one skinny GEMM, r_pad = 8 columns per group) stored NEXT to X in one row-major buffer
# [X | T], the layer is ONE GEMM over K' = K + r_pad against [W | sB]; in the backward [dY | dY sB] against the
# row-stacked [W ; A_cat] gives dX in ONE GEMM, and dA_cat = (dY sB)^T X, d(sB) = dY^T T are two skinny wgrads.
# The encoder's own weights are frozen by peft, so their wgrads / bias sums are skipped (needs_input_grad).
# ------------------------------------------------------------------------------------------
def _pad8(n):
    """LoRA rank groups are padded to 64 columns = one 128-byte K block of the GEMM tiles (see include/missm_b200.h)."""
    return (n + 63) // 64 * 64


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



class BlockPlan:
    """Static description of one residual block of a layer: kind, where its parameters sit in the flat parameter
    tuple the Function receives, and the owning module's operand cache."""
    __slots__ = ("kind", "cache", "n", "has_temb", "lora", "scaling", "eps", "which")

    def __init__(self, kind, cache, has_temb=False, lora=False, scaling=1.0, eps=1e-5, which="spatial"):
        self.kind, self.cache, self.has_temb, self.lora, self.scaling, self.eps = kind, cache, has_temb, lora, scaling, eps
        self.which = which                     # attention: "spatial" or "temporal" (which AttnMeta applies)
        # attn: ln_w ln_b qw qb kw kb vw vb ow ob [temb] [qA qB kA kB vA vB oA oB];  mlp: ln_w ln_b w1 b1 w2 b2
        self.n = (10 + int(has_temb) + (8 if lora else 0)) if kind == "attn" else 6


def attn_weights(blk, ps):
    ln_w, ln_b, qw, qb, kw, kb, vw, vb, ow, ob = ps[:10]
    temb = ps[10] if blk.has_temb else None
    if blk.lora:
        ad = ps[10 + int(blk.has_temb):]
        wf_qkv, wb_qkv, wf_o, wb_o, bqkv = lora_packs(blk.cache, (qw, qb, kw, kb, vw, vb, ow, ob), ad, blk.scaling)
        return blocks.AttnWeights(ln_w, ln_b, wf_qkv, bqkv, wf_o, ob, temb, ad[0].shape[0], wb_qkv, wb_o)
    wqkv, bqkv = packed_qkv(blk.cache, qw, kw, vw, qb, kb, vb)
    return blocks.AttnWeights(ln_w, ln_b, wqkv, bqkv, bf16_weight(blk.cache, "o", ow), ob, temb)


def mlp_weights(blk, ps):
    ln_w, ln_b, w1, b1, w2, b2 = ps
    return blocks.MlpWeights(ln_w, ln_b, bf16_weight(blk.cache, "fc1", w1), b1, bf16_weight(blk.cache, "fc2", w2), b2)


class EncoderLayerFn(torch.autograd.Function):
    """x -> layer(x) for a chain of residual blocks (`plan`: tuple of BlockPlan; `metas`: {"spatial": AttnMeta,
    "temporal": AttnMeta | None}); `params` = the blocks' parameters, flattened in plan order.  Returns an explicit
    gradient for every parameter that asks for one (DDP at train_ddp.py:189 runs with find_unused_parameters=False)."""

    @staticmethod
    def forward(ctx, x, plan, metas, *params):
        states, cur, i = [], x, 0
        for blk in plan:
            ps = params[i:i + blk.n]
            i += blk.n
            if blk.kind == "attn":
                cur, st = blocks.attn_fwd(metas[blk.which], cur, attn_weights(blk, ps))
            else:
                cur, st = blocks.mlp_fwd(blk.eps, cur, mlp_weights(blk, ps))
            states.append(st)
        states[0].keep = states[0].keep[:1]          # the layer input is kept by save_for_backward below
        ctx.plan, ctx.states = plan, states
        ctx.save_for_backward(x)
        return cur

    @staticmethod
    def backward(ctx, d_out):
        plan, states = ctx.plan, ctx.states
        need = ctx.needs_input_grad
        d = _contig(d_out)
        d_b, d_cs = _take_side(d)
        out, i = [None] * (len(need) - 3), len(need) - 3
        for blk, st in zip(reversed(plan), reversed(states)):
            i -= blk.n
            if blk.kind == "attn":
                wgrad = any(need[3 + i + 2:3 + i + 10])
                dx, dx_b, G = blocks.attn_bwd(st, d, d_b, d_cs is not None, wgrad)
                D = dx.shape[1]
                g = [G["ln_w"], G["ln_b"]] + [None] * 8
                if wgrad:
                    wq, bq = G["w_qkv"], G["b_qkv"]
                    g[2:10] = [wq[:D], bq[:D], wq[D:2 * D], bq[D:2 * D], wq[2 * D:], bq[2 * D:], G["w_o"],
                               d_cs if d_cs is not None else G["b_o"]]
                if blk.has_temb:
                    g.append(G["temb"].view(1, -1, D))
                if blk.lora:
                    r, s = st.w.lora_r, blk.scaling
                    for k in range(3):
                        g += [G["a_cat"][k * r:(k + 1) * r], G["sb_cat"][k * D:(k + 1) * D, k * r:(k + 1) * r] * s]
                    g += [G["a_o"][:r], G["sb_o"][:, :r] * s]
            else:
                wgrad = any(need[3 + i + 2:3 + i + 6])
                dx, dx_b, G = blocks.mlp_bwd(st, d, d_b, d_cs is not None, wgrad)
                g = [G["ln_w"], G["ln_b"]] + [None] * 4
                if wgrad:
                    g[2:6] = [G["w1"], G["b1"], G["w2"], d_cs if d_cs is not None else G["b2"]]
            out[i:i + blk.n] = g
            d, d_b, d_cs = dx, dx_b, G["dx_colsum"]
        ctx.states = None
        _give_side(d, d_b, d_cs)
        return (d, None, None, *out)


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
        Kpad = (K + 63) // 64 * 64          # whole 64-element K blocks: no TMA box of the GEMM leaves the tensor
        P = gh * gw
        wp = bf16_weight(cache, "patch", patch_w, cols_dst=Kpad)
        n_img = n_present * T
        tok = torch.empty((n_img * (P + 1), D), device=pixels.device, dtype=F32)
        # implicit GEMM: the tcgen05 kernel gathers the patches from the fp32 image itself; the im2col matrix exists
        # only in the backward, and only if the conv weight takes a gradient
        patches = None
        if not ops.patch_embed_implicit(pixels, wp, pos.detach(), tok, ps, T, sample_index=present_idx,
                                        n_samples=n_present):
            patches = ops.patchify(pixels, ps, Kpad, T, sample_index=present_idx, n_samples=n_present)
            ops.gemm(patches, wp, out=tok, epilogue=EPI_PATCH, aux_in=pos.detach(), patch_P=P)
        ops.cls_rows(cls.detach(), pos.detach(), tok, n_img, P + 1)
        x0, mean, rstd = ops.layernorm_fwd(tok, ln_w, ln_b, eps, out_dtype=F32)
        ctx.dims = (n_img, P, D, K, tuple(patch_w.shape))
        ctx.regather = None if patches is not None else (ps, Kpad, T, n_present)
        ctx.save_for_backward(tok, mean, rstd, patches, ln_w, pixels if patches is None else None,
                              present_idx if patches is None else None)
        return x0

    @staticmethod
    def backward(ctx, d_x0):
        tok, mean, rstd, patches, ln_w, pixels, present_idx = ctx.saved_tensors
        n_img, P, D, K, wshape = ctx.dims
        d_x0 = _contig(d_x0)
        _take_side(d_x0)
        d_tok, _, d_lnw, d_lnb, _ = ops.layernorm_bwd(d_x0, tok, mean, rstd, ln_w)
        d_pos, d_patch = ops.embed_bwd(d_tok, n_img, P + 1)
        d_w = None
        if ctx.needs_input_grad[6]:
            if patches is None:      # forward ran the implicit GEMM: gather the wgrad operand now
                ps, Kpad, T, n_present = ctx.regather
                patches = ops.patchify(pixels, ps, Kpad, T, sample_index=present_idx, n_samples=n_present)
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
        _take_side(d_x0)
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
        _give_side(dx, dx_b, dx_cs)
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


def encoder_layer(x, plan, metas, params):
    """One CLIPEncoderLayer over the fp32 residual stream x [M, D].  bf16 mode: ONE autograd node whose forward /
    backward issue one driver call per residual block; fp32 verification mode: the per-block fp32 Functions."""
    if get_precision() != "fp32":
        return EncoderLayerFn.apply(x, plan, metas, *params)
    from . import autograd_f32
    i = 0
    for blk in plan:
        ps = params[i:i + blk.n]
        i += blk.n
        if blk.kind == "mlp":
            x = autograd_f32.MlpBlockF32Fn.apply(x, blk.eps, blk.cache, *ps)
            continue
        meta = metas[blk.which]
        base, temb = ps[:10], (ps[10] if blk.has_temb else None)
        if blk.lora:
            # verification mode: the adapter folded into an effective fp32 weight W + s B A (exact in fp32; torch
            # autograd carries dW_eff back to A and B -- [D, r] parameter-space products, not a performance path)
            ln_w, ln_b, qw, qb, kw, kb, vw, vb, ow, ob = base
            qA, qB, kA, kB, vA, vB, oA, oB = ps[10 + int(blk.has_temb):]
            eff = [w + blk.scaling * (B @ A) for w, A, B in ((qw, qA, qB), (kw, kA, kB), (vw, vA, vB), (ow, oA, oB))]
            x = autograd_f32.AttnBlockF32Fn.apply(x, meta, {}, ln_w, ln_b, eff[0], qb, eff[1], kb, eff[2], vb,
                                                  eff[3], ob, temb)
        else:
            x = autograd_f32.AttnBlockF32Fn.apply(x, meta, blk.cache, *base, temb)
    return x


def vision_embed(*args):
    return _pick(VisionEmbedFn, "VisionEmbedF32Fn").apply(*args)


def pool_proj(*args):
    return _pick(PoolProjFn, "PoolProjF32Fn").apply(*args)
