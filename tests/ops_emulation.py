"""TEST INFRASTRUCTURE ONLY: torch-CPU stand-ins for the C-ABI entry points (include/missm_b200.h), written
from the ABI contracts, so that the HOST logic above the ABI (autograd wiring of the fp32 verification mode,
operand layouts of the 3-way split GEMM, mask compaction plumbing) can be exercised by `-m "not gpu"` tests in a
container without a GPU.  Nothing under missm-benchmark_b200/ imports this file; the product has no CPU path
(tests/test_host_logic.py::test_product_has_no_cpu_fallback).  The CUDA kernels themselves are checked by the
`-m gpu` tests.
"""
import contextlib

import torch

BF16, F32 = torch.bfloat16, torch.float32
EPI_LINEAR, EPI_GELU, EPI_RESID, EPI_DGELU, EPI_PATCH = range(5)


def _gelu(x):
    return x * torch.sigmoid(1.702 * x)


def _gelu_grad(x):
    s = torch.sigmoid(1.702 * x)
    return s + 1.702 * x * s * (1 - s)


# ------------------------------------------------------------------ missm_gemm_bf16 (bf16 operands, fp32 accumulate)
def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=None, bias=None, epilogue=EPI_LINEAR, aux_in=None,
         aux_out=None, scale_cols=0, col_scale=1.0, patch_P=0, out_rows=None, split_k=0, force_bn=0,
         colsum_out=None, colsum_part=None):
    assert a.dtype == BF16 and b.dtype == BF16
    out_dtype = BF16 if out_dtype is None else out_dtype
    A = (a.t() if a_mn else a).double()
    Bm = (b.t() if b_mn else b).double()
    assert A.shape[1] == Bm.shape[1], (A.shape, Bm.shape)
    M, N = A.shape[0], Bm.shape[0]
    assert N % 8 == 0 and (a.stride(0) * 2) % 16 == 0 and (b.stride(0) * 2) % 16 == 0, "TMA pitch / N rules"
    v = (A @ Bm.t()).float()
    if bias is not None:
        v = v + bias
    if epilogue == EPI_LINEAR:
        if scale_cols:
            v[:, :scale_cols] *= col_scale
    elif epilogue == EPI_GELU:
        aux_out.copy_(v.to(aux_out.dtype))
        v = _gelu(v)
    elif epilogue == EPI_RESID:
        v = v + aux_in.float()
    elif epilogue == EPI_DGELU:
        v = v * _gelu_grad(aux_in.float())
    elif epilogue == EPI_PATCH:
        r = torch.arange(M)
        v = v + aux_in[1 + r % patch_P]
    if out is None:
        out = torch.empty((out_rows if out_rows is not None else M, N), dtype=out_dtype)
    if epilogue == EPI_PATCH:
        r = torch.arange(M)
        out[(r // patch_P) * (patch_P + 1) + 1 + r % patch_P] = v.to(out.dtype)
    else:
        out[:M] = v.to(out.dtype)
    return out


# ------------------------------------------------------------------ missm_expand6_bf16 (csrc/fp32_mode.cu)
def expand6(x, which, stack_rows, cols_pad=None):
    assert x.dtype == F32 and x.dim() == 2 and x.stride(1) == 1
    R, C = x.shape
    cp = C if cols_pad is None else cols_pad
    p1 = x.to(BF16)
    r1 = x - p1.float()
    p2 = r1.to(BF16)
    p3 = (r1 - p2.float()).to(BF16)
    pc = [p1, p2, p3]
    order = ([1, 2, 0, 1, 0, 0], [1, 0, 2, 0, 1, 0])[which]
    if stack_rows:
        assert cp == C
        return torch.cat([pc[i] for i in order], dim=0)
    out = torch.zeros((R, 6 * cp), dtype=BF16)
    for p, i in enumerate(order):
        out[:, p * cp:p * cp + C] = pc[i]
    return out


# ------------------------------------------------------------------ attention (fp32 verification kernels)
def _seq_rows(lay):
    s = torch.arange(lay.n_seq)
    t = torch.arange(lay.N)
    return ((s // lay.s_in) * lay.seq_outer + (s % lay.s_in) * lay.seq_inner)[:, None] + t[None, :] * lay.tok_stride


def _attn_core(qkv, lay, H, causal, key_mask, mask_rows, mask_div):
    D = qkv.shape[1] // 3
    rows = _seq_rows(lay)                                      # [S, N]
    x = qkv[rows.reshape(-1)].reshape(lay.n_seq, lay.N, 3, H, D // H)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))          # [S, H, N, hd]
    sc = q @ k.transpose(-1, -2)
    neg = torch.zeros((lay.n_seq, 1, lay.N, lay.N))
    if causal:
        neg = neg + torch.full((lay.N, lay.N), float('-inf')).triu(1)
    if key_mask is not None:
        r = torch.arange(lay.n_seq) // mask_div
        if mask_rows is not None:
            r = mask_rows[r].long()
        km = key_mask[r]                                        # [S, N]
        neg = neg + torch.where(km == 0, float('-inf'), 0.0)[:, None, None, :]
    sc = sc + neg
    lse = torch.logsumexp(sc, dim=-1)                           # [S, H, N]
    o = torch.softmax(sc, dim=-1) @ v                           # [S, H, N, hd]
    out = torch.zeros((qkv.shape[0], D))
    out[rows.reshape(-1)] = o.permute(0, 2, 1, 3).reshape(-1, D)
    return out, lse


def attention_f32_fwd(qkv, lay, H, *, causal=False, key_mask=None, mask_rows=None, mask_div=1):
    assert qkv.dtype == F32
    with torch.no_grad():
        return _attn_core(qkv, lay, H, causal, key_mask, mask_rows, mask_div)


def attention_f32_bwd(qkv, out, lse, d_out, lay, H, q_scale, *, causal=False, key_mask=None, mask_rows=None,
                      mask_div=1):
    D = qkv.shape[1] // 3
    with torch.enable_grad():
        x = qkv.detach().clone().requires_grad_(True)
        o, _ = _attn_core(x, lay, H, causal, key_mask, mask_rows, mask_div)
        (g,) = torch.autograd.grad(o, x, d_out)
    g = g.clone()
    g[:, :D] *= q_scale
    return g


# ------------------------------------------------------------------ layernorm
def layernorm_fwd(x, gamma, beta, eps, *, out_dtype=None, row_index=None, n_rows=None, add_rows=None, add_period=0,
                  add_div=0, x_out=None, want_stats=True, out=None):
    out_dtype = BF16 if out_dtype is None else out_dtype
    xs = x
    if row_index is not None:
        xs = x[row_index[:n_rows].long() if n_rows is not None else row_index.long()]
    elif n_rows is not None:
        xs = x[:n_rows]
    if add_rows is not None:
        r = torch.arange(xs.shape[0])
        xs = xs + add_rows[(r // add_div) % add_period]
        (x_out if x_out is not None else x).copy_(xs)
    mean = xs.mean(-1)
    var = ((xs - mean[:, None]) ** 2).mean(-1)
    rstd = (var + eps).rsqrt()
    y = ((xs - mean[:, None]) * rstd[:, None] * gamma.detach() + beta.detach()).to(out_dtype)
    if out is not None:
        assert out.shape == y.shape and out.dtype == out_dtype
        out.copy_(y)
        y = out
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, *, dres=None, row_index=None, dx=None, want_bf16=False, dx_bf16=None):
    dyf = dy.float()
    xs = x if row_index is None else x[row_index[:dy.shape[0]].long()]
    xhat = (xs - mean[:, None]) * rstd[:, None]
    g = dyf * gamma.detach()
    d = rstd[:, None] * (g - g.mean(-1, keepdim=True) - xhat * (g * xhat).mean(-1, keepdim=True))
    if dres is not None:
        d = d + dres
    b16 = dx_bf16
    if row_index is None:
        dx = d
        if want_bf16 or dx_bf16 is not None:
            b16 = d.to(BF16)
    else:
        rows = row_index[:dy.shape[0]].long()
        dx[rows] = d
        if dx_bf16 is not None:
            dx_bf16[rows] = d.to(BF16)
    return dx, b16, (dyf * xhat).sum(0), dyf.sum(0), d.sum(0)


# ------------------------------------------------------------------ small kernels
def gelu_f32_fwd(u):
    return _gelu(u)


def gelu_f32_bwd(d_a, u):
    return d_a * _gelu_grad(u)


def colsum_grouped(x, period, div):
    r = torch.arange(x.shape[0])
    g = (r // div) % period
    out = torch.zeros((period, x.shape[1]))
    out.index_add_(0, g, x)
    return out


def copy_f32(src, dst):
    dst.copy_(src.reshape(dst.shape))
    return dst


def _sample_rows(t, sample_index, n):
    return t[sample_index[:n].long()] if sample_index is not None else t[:n]


def patchify_f32(pixels, ps, Kpad, T=1, sample_index=None, n_samples=None):
    Bn = n_samples if n_samples is not None else pixels.shape[0]
    px = _sample_rows(pixels, sample_index, Bn)
    if T == 1:
        px = px.unsqueeze(2)                                    # b c 1 h w
    b, c, t, h, w = px.shape
    gh, gw = h // ps, w // ps
    p = px[:, :, :, :gh * ps, :gw * ps].reshape(b, c, t, gh, ps, gw, ps).permute(0, 2, 3, 5, 1, 4, 6)
    p = p.reshape(b * t * gh * gw, c * ps * ps)
    out = torch.zeros((p.shape[0], Kpad))
    out[:, :p.shape[1]] = p
    return out


def cls_rows(cls, pos, tok, Bn, ntok):
    tok[torch.arange(Bn) * ntok] = cls + pos[0]


def embed_bwd(dtok, Bn, ntok):
    d = dtok.reshape(Bn, ntok, -1)
    return d.sum(0), d[:, 1:].reshape(Bn * (ntok - 1), -1).to(BF16)


def frame_mean(x, Bn, T, out_dtype=None):
    out_dtype = BF16 if out_dtype is None else out_dtype
    return x.reshape(Bn, T, -1).mean(1).to(out_dtype)


def frame_mean_bwd(dout, Bn, T):
    return (dout[:, None, :] / T).expand(Bn, T, dout.shape[1]).reshape(Bn * T, -1).contiguous()


def l2norm_scale_fwd(x, scale):
    inv = 1.0 / x.norm(dim=-1)
    return x * inv[:, None] * scale, inv


def l2norm_scale_bwd(dy, x, inv, scale, out_dtype=None):
    out_dtype = BF16 if out_dtype is None else out_dtype
    yh = x * inv[:, None]
    return (scale * inv[:, None] * (dy - yh * (dy * yh).sum(-1, keepdim=True))).to(out_dtype)


def text_embed_fwd(ids, tok_emb, pos_emb, sample_index=None, n_samples=None):
    Bn = n_samples if n_samples is not None else ids.shape[0]
    i = _sample_rows(ids, sample_index, Bn)
    return (tok_emb[i] + pos_emb[None, :ids.shape[1]]).reshape(Bn * ids.shape[1], -1)


def text_embed_bwd(ids, dx, vocab, sample_index=None, n_samples=None):
    Bn = n_samples if n_samples is not None else ids.shape[0]
    L, D = ids.shape[1], dx.shape[1]
    i = _sample_rows(ids, sample_index, Bn)
    dtok = torch.zeros((vocab, D))
    dtok.index_add_(0, i.reshape(-1), dx)
    return dtok, dx.reshape(Bn, L, D).sum(0)


def argmax_rows(ids, sample_index=None, n_samples=None):
    Bn = n_samples if n_samples is not None else ids.shape[0]
    i = _sample_rows(ids, sample_index, Bn)
    return (torch.arange(Bn) * ids.shape[1] + i.to(torch.int32).argmax(-1)).to(torch.int32)


def compact_mask(missing_index, codes):
    B, T = missing_index.numel(), len(codes)
    idx = torch.zeros((T, B), dtype=torch.int32)
    slot = torch.full((T, B), -1, dtype=torch.int32)
    counts = torch.zeros((T,), dtype=torch.int32)
    for t, c in enumerate(codes):
        present = torch.nonzero(missing_index != c).reshape(-1).to(torch.int32)
        n = present.numel()
        idx[t, :n] = present
        slot[t, present.long()] = torch.arange(n, dtype=torch.int32)
        counts[t] = n
    return idx, slot, counts


def scatter_rows_zero(src, slot_of, B):
    dst = torch.zeros((B, src.shape[1]))
    ok = slot_of[:B] >= 0
    dst[ok] = src[slot_of[:B][ok].long()]
    return dst


def gather_rows(src, idx, n_rows):
    return src[idx[:n_rows].long()].clone()


def masked_sum_norm(embs, weights, biases, codes, missing_index, gamma, beta, eps):
    """modal_sum.forward, src/model/baseline.py:52-61 (what csrc/fusion.cu computes)."""
    tot = 0
    for e, w, b, c in zip(embs, weights, biases, codes):
        y = torch.nn.functional.linear(e.float(), w, b)
        tot = tot + torch.where((missing_index.reshape(-1) == c)[:, None], torch.zeros_like(y), y)
    return torch.nn.functional.layer_norm(tot, (tot.shape[1],), gamma, beta, eps)


# ------------------------------------------------------------------ bf16-mode entry points (product path)
def attention_fwd(qkv, lay, H, *, causal=False, key_mask=None, mask_rows=None, mask_div=1, out=None):
    assert qkv.dtype == BF16
    with torch.no_grad():
        o, lse = _attn_core(qkv.float(), lay, H, causal, key_mask, mask_rows, mask_div)
    if out is None:
        return o.to(BF16), lse
    out.copy_(o.to(BF16))
    return out, lse


def attention_bwd(qkv, out, lse, d_out, lay, H, q_scale, *, causal=False, key_mask=None, mask_rows=None, mask_div=1,
                  dqkv_out=None, want_colsum=True):
    assert d_out.dtype == BF16 and d_out.stride(0) == out.stride(0)
    g = attention_f32_bwd(qkv.float(), out.float(), lse, d_out.float(), lay, H, q_scale, causal=causal,
                          key_mask=key_mask, mask_rows=mask_rows, mask_div=mask_div).to(BF16)
    if dqkv_out is not None:
        assert dqkv_out.stride(0) == qkv.stride(0)
        dqkv_out.copy_(g)
        g = dqkv_out
    return g, (g.float().sum(0) if want_colsum else None)


def cast_bf16(src, out=None, cols_dst=None):
    rows, cols = src.shape
    cols_dst = cols if cols_dst is None else cols_dst
    if out is None:
        out = torch.zeros((rows, cols_dst), dtype=BF16)
    out[:, :cols] = src.to(BF16)
    out[:, cols:cols_dst] = 0
    return out


def colsum(x):
    return x.float().sum(0)


def patchify(pixels, ps, Kpad, T=1, sample_index=None, n_samples=None):
    return patchify_f32(pixels, ps, Kpad, T, sample_index, n_samples).to(BF16)


def patch_embed_implicit(pixels, w_bf16, pos, tok, ps, T=1, sample_index=None, n_samples=None):
    """missm_patch_embed_implicit: tok[img * (P + 1) + 1 + patch] = patches(img, patch) . w^T + pos[1 + patch]."""
    D, Kpad = w_bf16.shape
    if ps != 14 or Kpad > 640 or D % 128:
        return False
    A = patchify(pixels, ps, Kpad, T, sample_index, n_samples)
    P = pos.shape[0] - 1
    n_img = A.shape[0] // P
    v = (A.double() @ w_bf16.double().t()).float().view(n_img, P, D) + pos[1:].unsqueeze(0)
    tok.view(n_img, P + 1, D)[:, 1:, :] = v
    return True


# ------------------------------------------------------------------ missm_image_preprocess (csrc/preprocess.cu)
def _cubic_aa(x):
    a = -0.5
    x = x.abs()
    return torch.where(x < 1, ((a + 2) * x - (a + 3)) * x * x + 1,
                       torch.where(x < 2, (((x - 5) * x + 8) * x - 4) * a, torch.zeros_like(x)))


def _resample_matrix(in_size, out_size, antialias):
    """[out_size, in_size] fp32 matrix of the header's resampling rule along one axis."""
    m = torch.zeros((out_size, in_size), dtype=F32)
    scale = torch.tensor(in_size, dtype=F32) / torch.tensor(out_size, dtype=F32)
    o = torch.arange(out_size, dtype=F32)
    if not antialias:
        A = -0.75
        # one rounding, as a fused multiply-add gives (aten's CPU kernel and nvcc both contract this expression)
        src = (scale.double() * (o.double() + 0.5) - 0.5).to(F32)
        fl = src.floor()
        t = src - fl
        c1 = lambda x: ((A + 2) * x - (A + 3)) * x * x + 1
        c2 = lambda x: ((A * x - 5 * A) * x + 8 * A) * x - 4 * A
        for j, w in enumerate((c2(t + 1), c1(t), c1(1 - t), c2(2 - t))):
            idx = (fl.long() - 1 + j).clamp(0, in_size - 1)
            m.index_put_((torch.arange(out_size), idx), w, accumulate=True)
        return m
    support = 2.0 * scale if scale >= 1 else torch.tensor(2.0)
    invscale = 1.0 / scale if scale >= 1 else torch.tensor(1.0)
    center = scale * (o + 0.5)
    lo = (center - support + 0.5).to(torch.int64).clamp_min(0)               # truncation toward zero, then max(., 0)
    hi = (center + support + 0.5).to(torch.int64).clamp_max(in_size)
    for i in range(out_size):
        j = torch.arange(int(lo[i]), int(hi[i]))
        w = _cubic_aa((j.to(F32) - center[i] + 0.5) * invscale)
        m[i, j] = w / w.sum()
    return m


def image_preprocess(src, out, S, mean, std, *, antialias, pre_div=255.0, clip_lo=float("-inf"), clip_hi=float("inf"),
                     post_div=1.0):
    H, W = src.shape[0], src.shape[1]
    v = src.to(F32)
    v = v.unsqueeze(-1).expand(H, W, 3) if v.dim() == 2 else v
    v = (v / pre_div).clamp(clip_lo, clip_hi) / post_div
    RH, RW = (S, int(S * W / H)) if H <= W else (int(S * H / W), S)
    top, left = int(round((RH - S) / 2.0)), int(round((RW - S) / 2.0))
    my = _resample_matrix(H, RH, antialias)[top:top + S]                      # only the crop is ever computed
    mx = _resample_matrix(W, RW, antialias)[left:left + S]
    r = torch.einsum('yh,hwc,xw->cyx', my, v, mx)
    out.copy_((r - torch.tensor(mean)[:, None, None]) / torch.tensor(std)[:, None, None])
    return out



# ------------------------------------------------------------------ residual-block drivers (csrc/blocks.cu)
# Stand-ins for missm_attn_block_* / missm_mlp_block_*, composed from the op stand-ins above in the order the
# header documents.  `state` is opaque to the product's autograd code; here it simply holds the saved tensors.
class _EmuState:
    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.keep = (None,)


def _pad8(n):
    """LoRA rank groups are padded to 64 columns = one 128-byte K block of the GEMM tiles (see include/missm_b200.h)."""
    return (n + 63) // 64 * 64


def block_attn_fwd(meta, x, w):
    M, D = x.shape
    r = w.lora_r
    R3, R1 = (_pad8(3 * r), _pad8(r)) if r else (0, 0)
    hcat = torch.zeros((M, D + R3), dtype=BF16)
    if w.temb is not None:
        x_res = torch.empty_like(x)
        _, mean, rstd = layernorm_fwd(x, w.ln_w, w.ln_b, meta.eps, add_rows=w.temb.detach().reshape(-1, D),
                                      add_period=meta.add_period, add_div=meta.add_div, x_out=x_res, out=hcat[:, :D],
                                      out_dtype=BF16)
    else:
        x_res = x
        _, mean, rstd = layernorm_fwd(x, w.ln_w, w.ln_b, meta.eps, out=hcat[:, :D], out_dtype=BF16)
    if r:
        gemm(hcat[:, :D], w.wb_qkv[3 * D:], out=hcat[:, D:])
    qkvcat = torch.zeros((M, 3 * D + R3), dtype=BF16)
    gemm(hcat, w.w_qkv[:, :D + R3], bias=w.b_qkv, scale_cols=D, col_scale=0.125, out=qkvcat[:, :3 * D])
    attncat = torch.zeros((M, D + R1), dtype=BF16)
    _, lse = attention_fwd(qkvcat[:, :3 * D], meta.layout, meta.H, causal=meta.causal, key_mask=meta.key_mask,
                           mask_rows=meta.mask_rows, mask_div=meta.mask_div, out=attncat[:, :D])
    if r:
        gemm(attncat[:, :D], w.wb_o[D:], out=attncat[:, D:])
    out = gemm(attncat, w.w_o[:, :D + R1], bias=w.b_o.detach(), epilogue=EPI_RESID, aux_in=x_res, out_dtype=F32)
    return out, _EmuState(meta=meta, w=w, x_res=x_res, mean=mean, rstd=rstd, hcat=hcat, qkvcat=qkvcat, attncat=attncat,
                          lse=lse, R3=R3, R1=R1)


def block_attn_bwd(st, d_out, d_out_bf16, colsum_given, wgrad):
    meta, w, R3, R1 = st.meta, st.w, st.R3, st.R1
    M, D = d_out.shape
    r = w.lora_r
    h, qkv, attn = st.hcat[:, :D], st.qkvcat[:, :3 * D], st.attncat[:, :D]
    G = {}
    dycat = torch.zeros((M, D + R1), dtype=BF16)
    dycat[:, :D] = d_out_bf16 if d_out_bf16 is not None else d_out.to(BF16)
    dy = dycat[:, :D]
    if wgrad and not colsum_given:
        G["b_o"] = colsum(dy)
    d_attncat = torch.zeros((M, D + R1), dtype=BF16)
    if r:
        gemm(dy, w.w_o[:, D:D + R1], b_mn=True, out=dycat[:, D:])
        gemm(dycat, w.wb_o, b_mn=True, out=d_attncat[:, :D])
        G["a_o"] = gemm(dycat[:, D:], attn, a_mn=True, b_mn=True, out_dtype=F32)
        G["sb_o"] = gemm(dy, st.attncat[:, D:], a_mn=True, b_mn=True, out_dtype=F32)
    else:
        gemm(dy, w.w_o, b_mn=True, out=d_attncat[:, :D])
    if wgrad:
        G["w_o"] = gemm(dy, attn, a_mn=True, b_mn=True, out_dtype=F32)
    dqkvcat = torch.zeros((M, 3 * D + R3), dtype=BF16)
    dqkv = dqkvcat[:, :3 * D]
    _, cs = attention_bwd(qkv, attn, st.lse, d_attncat[:, :D], meta.layout, meta.H, 0.125, causal=meta.causal,
                          key_mask=meta.key_mask, mask_rows=meta.mask_rows, mask_div=meta.mask_div, dqkv_out=dqkv,
                          want_colsum=wgrad)
    if wgrad:
        G["b_qkv"] = cs
    if r:
        gemm(dqkv, w.w_qkv[:, D:D + R3], b_mn=True, out=dqkvcat[:, 3 * D:])
        d_h = gemm(dqkvcat, w.wb_qkv, b_mn=True)
        G["a_cat"] = gemm(dqkvcat[:, 3 * D:], h, a_mn=True, b_mn=True, out_dtype=F32)
        G["sb_cat"] = gemm(dqkv, st.hcat[:, D:], a_mn=True, b_mn=True, out_dtype=F32)
    else:
        d_h = gemm(dqkv, w.w_qkv, b_mn=True)
    if wgrad:
        G["w_qkv"] = gemm(dqkv, h, a_mn=True, b_mn=True, out_dtype=F32)
    dx, dx_b, G["ln_w"], G["ln_b"], G["dx_colsum"] = layernorm_bwd(d_h, st.x_res, st.mean, st.rstd, w.ln_w, dres=d_out,
                                                                   want_bf16=True)
    if w.temb is not None:
        G["temb"] = colsum_grouped(dx, meta.add_period, meta.add_div)
    return dx, dx_b, G


def block_mlp_fwd(eps, x, w):
    h, mean, rstd = layernorm_fwd(x, w.ln_w, w.ln_b, eps, out_dtype=BF16)
    u = torch.empty((x.shape[0], w.w1.shape[0]), dtype=BF16)
    a = gemm(h, w.w1, bias=w.b1.detach(), epilogue=EPI_GELU, aux_out=u)
    out = gemm(a, w.w2, bias=w.b2.detach(), epilogue=EPI_RESID, aux_in=x, out_dtype=F32)
    return out, _EmuState(w=w, x=x, mean=mean, rstd=rstd, h=h, u=u, a=a)


def block_mlp_bwd(st, d_out, d_out_bf16, colsum_given, wgrad):
    w = st.w
    dy = d_out_bf16 if d_out_bf16 is not None else d_out.to(BF16)
    G = {}
    if wgrad and not colsum_given:
        G["b2"] = colsum(dy)
    if wgrad:
        G["w2"] = gemm(dy, st.a, a_mn=True, b_mn=True, out_dtype=F32)
    d_u = gemm(dy, w.w2, b_mn=True, epilogue=EPI_DGELU, aux_in=st.u)
    if wgrad:
        G["w1"] = gemm(d_u, st.h, a_mn=True, b_mn=True, out_dtype=F32)
        G["b1"] = colsum(d_u)
    d_h = gemm(d_u, w.w1, b_mn=True)
    dx, dx_b, G["ln_w"], G["ln_b"], G["dx_colsum"] = layernorm_bwd(d_h, st.x, st.mean, st.rstd, w.ln_w, dres=d_out,
                                                                   want_bf16=True)
    return dx, dx_b, G


BLOCK_DRIVERS = {"attn_fwd": block_attn_fwd, "attn_bwd": block_attn_bwd, "mlp_fwd": block_mlp_fwd,
                 "mlp_bwd": block_mlp_bwd}

BF16_MODE_OPS = ["attention_fwd", "attention_bwd", "cast_bf16", "colsum", "patchify", "patch_embed_implicit"]

EMULATED_OPS = ["gemm", "expand6", "attention_f32_fwd", "attention_f32_bwd", "layernorm_fwd", "layernorm_bwd",
                "gelu_f32_fwd", "gelu_f32_bwd", "colsum_grouped", "copy_f32", "patchify_f32", "cls_rows", "embed_bwd",
                "frame_mean", "frame_mean_bwd", "l2norm_scale_fwd", "l2norm_scale_bwd", "text_embed_fwd",
                "text_embed_bwd", "argmax_rows", "compact_mask", "scatter_rows_zero", "gather_rows"]


@contextlib.contextmanager
def emulated_fp32_mode(precision="fp32", wide_bf16=False):
    """Patch missm_b200.ops (+ the fused `sum` head and the CUDA-only guards) with the stand-ins above and switch
    the host side to the fp32 verification mode.  ops.gemm_f32 / colsum_f32 stay the REAL host code.
    precision="bf16" exercises the PRODUCT blocks of autograd.py instead; with wide_bf16 every "bf16" buffer is
    really fp32 (the dtype constant is swapped), which isolates the host algebra -- operand views, pitches,
    gradient formulas -- from bf16 rounding."""
    from missm_b200 import autograd as ag, bank, blocks, fusion_ops, ops, towers
    here = globals()
    global BF16
    names = EMULATED_OPS + (BF16_MODE_OPS if precision == "bf16" else [])
    saved = {n: getattr(ops, n) for n in names}
    saved_dt = (BF16, ag.BF16, ops.BF16)
    saved_drivers = {n: getattr(blocks, n) for n in BLOCK_DRIVERS}
    if wide_bf16:
        BF16 = ag.BF16 = ops.BF16 = torch.float32
    saved_guard, saved_bank_guard, saved_fusion = towers._require_cuda, bank._require_cuda_index, fusion_ops.masked_sum_norm
    for n in names:
        setattr(ops, n, here[n])
    for n, f in BLOCK_DRIVERS.items():
        setattr(blocks, n, f)
    towers._require_cuda = lambda t, what: None
    bank._require_cuda_index = lambda mi, mdev: mi
    fusion_ops.masked_sum_norm = masked_sum_norm
    old = ag.set_precision(precision)
    try:
        yield
    finally:
        ag.set_precision(old)
        BF16, ag.BF16, ops.BF16 = saved_dt
        for n, f in saved.items():
            setattr(ops, n, f)
        for n, f in saved_drivers.items():
            setattr(blocks, n, f)
        towers._require_cuda, bank._require_cuda_index, fusion_ops.masked_sum_norm = saved_guard, saved_bank_guard, saved_fusion
