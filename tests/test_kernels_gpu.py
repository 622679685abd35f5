"""Unit parity of every CUDA kernel against a plain fp32 torch statement of the same op
(inputs rounded to bf16 where the kernel consumes bf16).  Run with `-m gpu` on a B200."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from missm_b200 import ops as o
    o.lib()
    return o


# ------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("a_mn", [False, True])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("shape", [(128, 256, 64), (300, 520, 200), (1799, 3072, 1024), (64, 768, 1024)])
def test_gemm_layouts(ops, a_mn, b_mn, shape):
    M, N, K = shape
    torch.manual_seed(M + N + K)
    Mp = (M + 7) // 8 * 8 if a_mn else M
    A = torch.randn(Mp, K, device=DEV).bfloat16()
    B = torch.randn(N, K, device=DEV).bfloat16()
    ref = A.float() @ B.float().t()
    for bn in (128, 256):
        out = ops.gemm(A.t().contiguous() if a_mn else A, B.t().contiguous() if b_mn else B,
                       a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, force_bn=bn, split_k=1)
        assert rel(out, ref) < 1e-5


@pytest.mark.parametrize("M", [700, 1500])       # 700: one-CTA kernel; 1500: CTA-pair (cta_group::2) kernel
def test_gemm_epilogues(ops, M):
    torch.manual_seed(1)
    N, K = 1024, 512
    A = torch.randn(M, K, device=DEV).bfloat16()
    B = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(N, device=DEV)
    acc = A.float() @ B.float().t()
    pre = acc + bias
    out = ops.gemm(A, B, bias=bias, scale_cols=512, col_scale=0.125)
    ref = pre.clone(); ref[:, :512] *= 0.125
    assert rel(out, ref) < 4e-3
    u = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    out = ops.gemm(A, B, bias=bias, epilogue=ops.EPI_GELU, aux_out=u)
    assert rel(u, pre) < 4e-3 and rel(out, pre * torch.sigmoid(1.702 * pre)) < 4e-3
    res = torch.randn(M, N, device=DEV)
    out = ops.gemm(A, B, bias=bias, epilogue=ops.EPI_RESID, aux_in=res, out_dtype=torch.float32)
    assert rel(out, res + pre) < 1e-5
    res2 = res.clone()
    ops.gemm(A, B, bias=bias, epilogue=ops.EPI_RESID, aux_in=res2, out=res2)
    assert rel(res2, res + pre) < 1e-5
    uu = torch.randn(M, N, device=DEV).bfloat16()
    cs = torch.zeros(N, device=DEV)
    out = ops.gemm(A, B, epilogue=ops.EPI_DGELU, aux_in=uu, colsum_out=cs)
    s = torch.sigmoid(1.702 * uu.float())
    dref = acc * (s * (1 + 1.702 * uu.float() * (1 - s)))
    assert rel(out, dref) < 4e-3
    assert rel(cs, dref.sum(0)) < 2e-5              # fused bias gradient (fp32 values, before the bf16 rounding)
    # the deterministic variant: per-32-row column sums of the values AS STORED, written by the owning warp (no atomics)
    part = torch.full(((M + 31) // 32, N), float("nan"), device=DEV)
    out2 = ops.gemm(A, B, epilogue=ops.EPI_DGELU, aux_in=uu, colsum_part=part)
    assert torch.equal(out2, out) and not torch.isnan(part).any()
    ref_part = torch.nn.functional.pad(out2.float(), (0, 0, 0, (-M) % 32)).view(-1, 32, N).sum(1)
    assert rel(part, ref_part) < 1e-6
    P = 100
    Bsz = M // P
    A, acc = A[:Bsz * P], acc[:Bsz * P]
    pos = torch.randn(P + 1, N, device=DEV)
    tok = torch.zeros(Bsz * (P + 1), N, device=DEV)
    ops.gemm(A, B, epilogue=ops.EPI_PATCH, aux_in=pos, out=tok, patch_P=P)
    ref = torch.zeros(Bsz, P + 1, N, device=DEV); ref[:, 1:] = acc.view(Bsz, P, N) + pos[1:]
    assert rel(tok, ref.view(-1, N)) < 1e-5


def test_gemm_splitk_wgrad(ops):
    torch.manual_seed(2)
    Mtok = 257 * 9
    dY = torch.randn(Mtok, 1024, device=DEV).bfloat16()
    X = torch.randn(Mtok, 1024, device=DEV).bfloat16()
    out = ops.gemm(dY, X, a_mn=True, b_mn=True, out_dtype=torch.float32)
    assert rel(out, dY.float().t() @ X.float()) < 1e-5
    # K tail (592-wide patch rows) and M == 0
    Xp = torch.randn(Mtok, 592, device=DEV).bfloat16()
    out = ops.gemm(dY, Xp, a_mn=True, b_mn=True, out_dtype=torch.float32)
    assert rel(out, dY.float().t() @ Xp.float()) < 1e-5
    e = ops.gemm(torch.empty(0, 64, device=DEV, dtype=torch.bfloat16), X[:8, :64].contiguous())
    assert e.shape == (0, 8)


# ------------------------------------------------------------------------------- layernorm
@pytest.mark.parametrize("D", [256, 768, 1024])
def test_layernorm_fwd_bwd(ops, D):
    torch.manual_seed(D)
    M = 1000
    x = (torch.randn(M, D, device=DEV) * 3 + 0.5).requires_grad_(True)
    g = torch.randn(D, device=DEV).requires_grad_(True)
    b = torch.randn(D, device=DEV).requires_grad_(True)
    ref = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5)
    y, mean, rstd = ops.layernorm_fwd(x.detach(), g.detach(), b.detach(), 1e-5, out_dtype=torch.float32)
    assert rel(y, ref) < 1e-6
    yb, _, _ = ops.layernorm_fwd(x.detach(), g.detach(), b.detach(), 1e-5)
    assert rel(yb, ref) < 4e-3
    dy = torch.randn(M, D, device=DEV)
    dres = torch.randn(M, D, device=DEV)
    ref.backward(dy)
    dx, dxb, dg, db, dcs = ops.layernorm_bwd(dy, x.detach(), mean, rstd, g.detach(), dres=dres, want_bf16=True)
    assert rel(dx, x.grad + dres) < 1e-5
    assert rel(dcs, (x.grad + dres).sum(0)) < 1e-5          # fused column sums (bias gradient of the producer)
    assert rel(dxb, x.grad + dres) < 4e-3
    assert rel(dg, g.grad) < 1e-5 and rel(db, b.grad) < 1e-5
    # bf16 dy
    dx2, _, dg2, _, _ = ops.layernorm_bwd(dy.bfloat16(), x.detach(), mean, rstd, g.detach())
    x.grad = None; g.grad = None
    torch.nn.functional.layer_norm(x, (D,), g, b, 1e-5).backward(dy.bfloat16().float())
    assert rel(dx2, x.grad) < 1e-5 and rel(dg2, g.grad) < 1e-5


def test_layernorm_gather_and_add(ops):
    torch.manual_seed(3)
    D, N, B, T = 256, 5, 6, 3
    x = torch.randn(B * T * N, D, device=DEV)
    g, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    rows = (torch.arange(B * T, device=DEV, dtype=torch.int32) * N)
    y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-5, out_dtype=torch.float32, row_index=rows)
    ref = torch.nn.functional.layer_norm(x[rows.long()], (D,), g, b, 1e-5)
    assert rel(y, ref) < 1e-6
    dy = torch.randn(B * T, D, device=DEV)
    dx = torch.zeros_like(x)
    dcs = ops.layernorm_bwd(dy, x, mean, rstd, g, row_index=rows, dx=dx)[4]
    xr = x.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr[rows.long()], (D,), g, b, 1e-5).backward(dy)
    assert rel(dx, xr.grad) < 1e-5
    assert rel(dcs, xr.grad.sum(0)) < 1e-5
    # temporal embedding add: x <- x + temb[(row // N) % T]
    temb = torch.randn(T, D, device=DEV)
    x2 = x.clone()
    y2, _, _ = ops.layernorm_fwd(x2, g, b, 1e-5, out_dtype=torch.float32, add_rows=temb, add_period=T, add_div=N)
    xe = (x.view(B, T, N, D) + temb[None, :, None, :]).reshape(-1, D)
    assert rel(x2, xe) < 1e-6
    assert rel(y2, torch.nn.functional.layer_norm(xe, (D,), g, b, 1e-5)) < 1e-6


# ------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, causal, kmask):
    # q,k,v: [S, H, N, hd] fp32; q already scaled
    s = q @ k.transpose(-1, -2)
    N = q.shape[2]
    if causal:
        s = s + torch.full((N, N), float("-inf"), device=q.device).triu(1)
    if kmask is not None:
        s = s.masked_fill(kmask[:, None, None, :] == 0, float("-inf"))
    return torch.softmax(s, -1) @ v


@pytest.mark.parametrize("cfg", [(3, 16, 257, False, False), (2, 12, 77, True, True), (2, 16, 593, False, False),
                                 (5, 4, 8, False, False), (1, 2, 64, True, False), (2, 2, 130, False, True),
                                 # tcgen05 path: several items per CTA with the odd row on the CUDA cores (257, 129),
                                 # and shapes without an odd row (256, 272, 200)
                                 (24, 16, 257, False, False), (40, 8, 129, False, False), (3, 4, 256, False, False),
                                 (2, 4, 272, False, False), (3, 4, 200, False, False),
                                 # one step per item and several items per CTA (N <= 96)
                                 (30, 16, 96, False, False), (40, 16, 64, False, False)])
def test_attention_fwd_bwd(ops, cfg):
    S, H, N, causal, use_mask = cfg
    torch.manual_seed(N)
    D = H * 64
    qkv = (torch.randn(S * N, 3 * D, device=DEV) * 0.7).bfloat16()
    kmask = None
    if use_mask:
        lens = torch.randint(1, N + 1, (S,), device=DEV)
        kmask = (torch.arange(N, device=DEV)[None, :] < lens[:, None]).long().contiguous()
    lay = ops.SeqLayout.spatial(S, N)
    out, lse = ops.attention_fwd(qkv, lay, H, causal=causal, key_mask=kmask)
    f = qkv.float().view(S, N, 3, H, 64).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    ref = _attn_ref(f[0], f[1], f[2], causal, kmask)          # [S,H,N,hd]
    ref_o = ref.permute(0, 2, 1, 3).reshape(S * N, D)
    assert rel(out, ref_o) < 6e-3
    last = torch.arange(S, device=DEV) * N + N - 1              # the odd row of every sequence, on its own
    assert rel(out[last], ref_o[last]) < 6e-3
    if not causal and kmask is None:
        lse_ref = torch.logsumexp(f[0].detach() @ f[1].detach().transpose(-1, -2), -1)      # [S,H,N]
        assert (lse - lse_ref).abs().max() < 2e-2
    d_out = (torch.randn(S * N, D, device=DEV)).bfloat16()
    ref_o.backward(d_out.float())
    q_scale = 0.125
    dqkv, dcs = ops.attention_bwd(qkv, out, lse, d_out, lay, H, q_scale, causal=causal, key_mask=kmask)
    gref = f.grad.permute(1, 3, 0, 2, 4).reshape(S * N, 3 * D).clone()
    gref[:, :D] *= q_scale
    assert rel(dqkv[:, :D], gref[:, :D]) < 1.5e-2
    assert rel(dqkv[:, D:2 * D], gref[:, D:2 * D]) < 1.5e-2
    assert rel(dqkv[:, 2 * D:], gref[:, 2 * D:]) < 1.5e-2
    if N % 128 == 1:                                    # the odd row computed on the CUDA cores
        assert rel(dqkv[last], gref[last]) < 1.5e-2
    assert rel(dcs, gref.sum(0)) < 1.5e-2               # q/k/v bias gradients


@pytest.mark.parametrize("T", [8, 5, 3, 1])
def test_attention_temporal_layout(ops, T):
    """Temporal attention reads [(b t) n d] in place: sequence (b, n) over t.  T <= 8 runs the one-warp-per-
    (sequence, head) kernel of csrc/attention_small.cu, forward and backward, vs fp32 torch autograd."""
    torch.manual_seed(5 + T)
    B, N, H = 2, 5, 2
    D = H * 64
    qkv = (torch.randn(B * T * N, 3 * D, device=DEV) * 0.7).bfloat16()
    lay = ops.SeqLayout.temporal(B, T, N)
    out, lse = ops.attention_fwd(qkv, lay, H)
    f = qkv.float().view(B, T, N, 3, H, 64).permute(3, 0, 2, 4, 1, 5).reshape(3, B * N, H, T, 64).contiguous()
    f.requires_grad_(True)
    ref = _attn_ref(f[0], f[1], f[2], False, None)           # [(b n), H, T, hd]
    ref_o = ref.view(B, N, H, T, 64).permute(0, 3, 1, 2, 4).reshape(B * T * N, D)
    assert rel(out, ref_o) < 6e-3
    lse_ref = torch.logsumexp(f[0].detach() @ f[1].detach().transpose(-1, -2), -1)      # [(b n), H, T]
    assert (lse - lse_ref).abs().max() < 2e-2
    d_out = torch.randn(B * T * N, D, device=DEV).bfloat16()
    ref_o.backward(d_out.float())
    dqkv, dcs = ops.attention_bwd(qkv, out, lse, d_out, lay, H, 0.125)
    g = f.grad.view(3, B, N, H, T, 64).permute(1, 4, 2, 0, 3, 5).reshape(B * T * N, 3 * D).clone()
    g[:, :D] *= 0.125
    for c in range(3):
        assert rel(dqkv[:, c * D:(c + 1) * D], g[:, c * D:(c + 1) * D]) < 1.5e-2, c
    assert rel(dcs, g.sum(0)) < 1.5e-2


# --------------------------------------------------------------------------------- helpers
def test_cast_colsum_patchify(ops):
    torch.manual_seed(6)
    w = torch.randn(300, 588, device=DEV)
    wb = ops.cast_bf16(w, cols_dst=592)
    assert torch.equal(wb[:, :588], w.bfloat16()) and (wb[:, 588:] == 0).all()
    x = torch.randn(3000, 1024, device=DEV).bfloat16()
    assert rel(ops.colsum(x), x.float().sum(0)) < 1e-5
    px = torch.randn(5, 3, 28, 70, device=DEV)
    idx = torch.tensor([4, 0, 2], device=DEV, dtype=torch.int32)
    pt = ops.patchify(px, 14, 592, sample_index=idx, n_samples=3)
    vid = torch.randn(4, 3, 2, 28, 28, device=DEV)
    pv = ops.patchify(vid, 14, 592, T=2, sample_index=idx[:2] % 4, n_samples=2)
    fr = vid[(idx[:2] % 4).long()].permute(0, 2, 1, 3, 4).reshape(4, 3, 28, 28)
    rv = torch.nn.functional.unfold(fr, kernel_size=14, stride=14).transpose(1, 2).reshape(-1, 588)
    assert torch.equal(pv[:, :588], rv.bfloat16())
    ref = torch.nn.functional.unfold(px[idx.long()], kernel_size=14, stride=14).transpose(1, 2).reshape(-1, 588)
    assert torch.equal(pt[:, :588], ref.bfloat16()) and (pt[:, 588:] == 0).all()


@pytest.mark.parametrize("case", ["image", "ragged", "video", "audio_grid"])
def test_patch_embed_implicit_gemm(ops, case):
    """The implicit-GEMM patch embedding (csrc/patch_embed_tc.cu: tcgen05 fed from the fp32 image, no im2col matrix)
    against (a) the explicit path it replaces, patchify + GEMM with the EPI_PATCH epilogue, and (b) the fp32
    convolution of the reference (Conv2d k = stride = 14, no bias, + position rows; modeling_video.py:29-51)."""
    torch.manual_seed(11)
    D, ps, Kpad = 256, 14, 640
    T = 1
    if case == "image":          # 3 gathered samples of 5, 224 x 224: 768 patches = 6 full tiles
        px = torch.randn(5, 3, 224, 224, device=DEV)
        idx, n = torch.tensor([4, 0, 2], device=DEV, dtype=torch.int32), 3
    elif case == "ragged":       # 3 x 8 x 5 = 120 patches: one partial tile, no gather
        px = torch.randn(3, 3, 112, 70, device=DEV)
        idx, n = None, 3
    elif case == "video":        # [B, C, T, H, W], 2 of 4 clips, 2 frames each
        px = torch.randn(4, 3, 2, 56, 84, device=DEV)
        idx, n, T = torch.tensor([3, 1], device=DEV, dtype=torch.int32), 2, 2
    else:                        # the audio tower's 8 x 74 patch grid (112 x 1036 spectrogram), 2 samples
        px = torch.randn(2, 3, 112, 1036, device=DEV)
        idx, n = None, 2
    H, W = px.shape[-2:]
    P = (H // ps) * (W // ps)
    w = torch.randn(D, 3, ps, ps, device=DEV) * 0.05
    wb = ops.cast_bf16(w.reshape(D, -1), cols_dst=Kpad)
    pos = torch.randn(P + 1, D, device=DEV)
    n_img = n * T
    tok = torch.full((n_img * (P + 1), D), 7.0, device=DEV)
    assert ops.patch_embed_implicit(px, wb, pos, tok, ps, T, sample_index=idx, n_samples=n)
    ref_tok = torch.full_like(tok, 7.0)
    patches = ops.patchify(px, ps, Kpad, T, sample_index=idx, n_samples=n)
    ops.gemm(patches, wb, out=ref_tok, epilogue=ops.EPI_PATCH, aux_in=pos, patch_P=P)
    torch.cuda.synchronize()
    # same bf16 operands, same K order, fp32 accumulation: the two tensor-core paths agree to rounding of the sum
    assert rel(tok, ref_tok) < 1e-6, rel(tok, ref_tok)
    assert torch.equal(tok.view(n_img, P + 1, D)[:, 0], torch.full((n_img, D), 7.0, device=DEV))     # CLS rows untouched
    sel = px if idx is None else px[idx.long()]
    if T > 1:
        sel = sel.permute(0, 2, 1, 3, 4).reshape(n_img, 3, H, W)
    conv = torch.nn.functional.conv2d(sel, w, stride=ps).flatten(2).transpose(1, 2) + pos[1:]
    assert rel(tok.view(n_img, P + 1, D)[:, 1:], conv) < 6e-3


def test_embed_pool_l2norm(ops):
    torch.manual_seed(7)
    B, ntok, D = 4, 9, 256
    dtok = torch.randn(B * ntok, D, device=DEV)
    dpos, dpatch = ops.embed_bwd(dtok, B, ntok)
    assert rel(dpos, dtok.view(B, ntok, D).sum(0)) < 1e-6
    assert torch.equal(dpatch, dtok.view(B, ntok, D)[:, 1:].reshape(-1, D).bfloat16())
    x = torch.randn(B * 3, D, device=DEV)
    assert rel(ops.frame_mean(x, B, 3, torch.float32), x.view(B, 3, D).mean(1)) < 1e-6
    d = torch.randn(B, D, device=DEV)
    assert rel(ops.frame_mean_bwd(d, B, 3), (d[:, None, :] / 3).expand(B, 3, D).reshape(-1, D)) < 1e-6
    xr = torch.randn(B, 768, device=DEV, requires_grad=True)
    scale = math.exp(2.6592)
    ref = xr / xr.norm(p=2, dim=-1, keepdim=True) * scale
    y, inv = ops.l2norm_scale_fwd(xr.detach(), scale)
    assert rel(y, ref) < 1e-6
    dy = torch.randn(B, 768, device=DEV)
    ref.backward(dy)
    assert rel(ops.l2norm_scale_bwd(dy, xr.detach(), inv, scale, torch.float32), xr.grad) < 1e-5


def test_text_embed_argmax(ops):
    torch.manual_seed(8)
    B, L, D, V = 5, 77, 256, 1000
    ids = torch.randint(1, V - 2, (B, L), device=DEV)
    ids[:, 0] = V - 2
    for b in range(B):
        ids[b, 5 + b:] = V - 1          # EOT + EOT padding -> first maximum wins
    tok, pos = torch.randn(V, D, device=DEV), torch.randn(L, D, device=DEV)
    sel = torch.tensor([3, 1, 4], device=DEV, dtype=torch.int32)
    out = ops.text_embed_fwd(ids, tok, pos, sample_index=sel, n_samples=3)
    ref = tok[ids[sel.long()]] + pos[None]
    assert torch.equal(out, ref.view(-1, D))
    rows = ops.argmax_rows(ids, sample_index=sel, n_samples=3)
    exp = torch.arange(3, device=DEV) * L + ids[sel.long()].to(torch.int32).argmax(-1)
    assert torch.equal(rows.long(), exp)
    dx = torch.randn(3 * L, D, device=DEV)
    dtok, dpos = ops.text_embed_bwd(ids, dx, V, sample_index=sel, n_samples=3)
    rt = torch.zeros(V, D, device=DEV).index_add_(0, ids[sel.long()].view(-1), dx)
    assert rel(dtok, rt) < 1e-5 and rel(dpos, dx.view(3, L, D).sum(0)) < 1e-6


# ------------------------------------------------------------------------------ compaction
@pytest.mark.parametrize("B", [1, 8, 64, 333])
def test_compaction_bit_exact(ops, B):
    """Indices must equal torch.nonzero(missing_index != code) exactly."""
    g = torch.Generator().manual_seed(B)
    mi = torch.randint(0, 7, (B,), generator=g).to(DEV)
    codes = [1, 2, 3, 4, 5, 6]
    idx, slot, counts = ops.compact_mask(mi, codes)
    c = counts.cpu()
    for t, code in enumerate(codes):
        exp = torch.nonzero(mi != code).flatten().to(torch.int32)
        assert int(c[t]) == exp.numel()
        assert torch.equal(idx[t, :exp.numel()], exp)
        s = torch.full((B,), -1, dtype=torch.int32, device=DEV)
        s[exp.long()] = torch.arange(exp.numel(), dtype=torch.int32, device=DEV)
        assert torch.equal(slot[t], s)
    # gather / zero-filled scatter round trip
    x = torch.randn(B, 768, device=DEV)
    n = int(c[3])
    gath = ops.gather_rows(x, idx[3], n)
    assert torch.equal(gath, x[idx[3, :n].long()])
    back = ops.scatter_rows_zero(gath, slot[3], B)
    ref = x.clone(); ref[mi == codes[3]] = 0
    assert torch.equal(back, ref)
