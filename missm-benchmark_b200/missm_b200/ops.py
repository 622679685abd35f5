"""Thin Python wrappers over the C ABI (include/missm_b200.h).

Shape/dtype validation happens here, in Python; the CUDA entry points receive raw device
pointers plus torch's current stream.  Nothing in this file computes on the CPU and nothing
falls back to torch ops: if the library is missing, `_lib.lib()` raises.
"""
import ctypes
import os

import torch

from ._lib import AttnArgs, GemmArgs, check, lib, stream_ptr

EPI_LINEAR, EPI_GELU, EPI_RESID, EPI_DGELU, EPI_PATCH = range(5)
BF16, F32 = torch.bfloat16, torch.float32

# (bench.py's bookkeeping -- kernels launched, CUDA events around every GEMM -- lives in the library itself:
#  missm_launch_count / missm_gemm_profile, so that launches issued by the block drivers are seen too)


def set_persistent_sms(n):
    """SMs the persistent kernels use from now on (n <= 0: default).  See include/missm_b200.h."""
    check(lib().missm_set_persistent_sms(int(n)), "set_persistent_sms")


def set_coresident(on):
    """Co-resident variants of the persistent kernels on / off (include/missm_b200.h: missm_set_coresident)."""
    check(lib().missm_set_coresident(int(bool(on))), "set_coresident")


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, "expected a row-major 2-D tensor"
    return t.stride(0)


# --------------------------------------------------------------------------------------- GEMM
def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=BF16, bias=None,
         epilogue=EPI_LINEAR, aux_in=None, aux_out=None, scale_cols=0, col_scale=1.0,
         patch_P=0, out_rows=None, split_k=0, force_bn=0, colsum_out=None, colsum_part=None):
    """C[m,n] = epilogue(sum_k A[m,k] B[n,k]).  `a`: [M,K] (or [K,M] if a_mn), `b`: [N,K]
    (or [K,N] if b_mn); both bf16 CUDA tensors, row-major.  `colsum_out` (f32 [N], pre-zeroed) receives
    the column sums of the fp32 values written to C (fused bias gradient)."""
    assert a.is_cuda and b.is_cuda and a.dtype == BF16 and b.dtype == BF16
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    assert K == Kb, f"contraction mismatch {K} vs {Kb}"
    if out is None:
        out = torch.empty((out_rows if out_rows is not None else M, N), device=a.device,
                          dtype=out_dtype)
    assert out.dtype in (BF16, F32)
    g = GemmArgs()
    g.A, g.B, g.C = a.data_ptr(), b.data_ptr(), out.data_ptr()
    if bias is not None:
        assert bias.dtype == F32 and bias.numel() == N and bias.is_contiguous()
        g.bias = bias.data_ptr()
    g.aux_in = aux_in.data_ptr() if aux_in is not None else None
    g.aux_out = aux_out.data_ptr() if aux_out is not None else None
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldc = _ld(a), _ld(b), _ld(out)
    g.ld_aux_in = _ld(aux_in) if aux_in is not None else 0
    g.ld_aux_out = _ld(aux_out) if aux_out is not None else 0
    g.a_mn, g.b_mn = int(a_mn), int(b_mn)
    g.epilogue = epilogue
    g.out_f32 = int(out.dtype == F32)
    g.scale_cols, g.col_scale = scale_cols, col_scale
    g.patch_P, g.split_k, g.force_bn = patch_P, split_k, force_bn
    if colsum_out is not None:
        assert colsum_out.dtype == F32 and colsum_out.numel() == N and colsum_out.is_contiguous()
        g.colsum_out = colsum_out.data_ptr()
    if colsum_part is not None:      # f32 [ceil(M / 32), N]: per-32-row column sums of the stored values (no atomics)
        assert colsum_part.dtype == F32 and colsum_part.is_contiguous() and tuple(colsum_part.shape) == ((M + 31) // 32, N)
        g.colsum_part = colsum_part.data_ptr()
    check(lib().missm_gemm_bf16(ctypes.byref(g), stream_ptr()), "gemm_bf16")
    return out


# ---------------------------------------------------------------------------------- attention
class SeqLayout:
    """Where sequence s / token t lives: row = (s // s_in)*seq_outer + (s % s_in)*seq_inner
    + t*tok_stride."""

    def __init__(self, n_seq, N, s_in=1, seq_outer=None, seq_inner=0, tok_stride=1):
        self.n_seq, self.N, self.s_in = n_seq, N, s_in
        self.seq_outer = N if seq_outer is None else seq_outer
        self.seq_inner, self.tok_stride = seq_inner, tok_stride

    @staticmethod
    def spatial(n_seq, N):
        return SeqLayout(n_seq, N)

    @staticmethod
    def temporal(B, T, N):
        # activations are [(b t) n d]; sequence (b, n) runs over t
        return SeqLayout(B * N, T, s_in=N, seq_outer=T * N, seq_inner=1, tok_stride=N)


def _attn_args(qkv, out, lse, lay, H, causal, key_mask, mask_rows, mask_div):
    D = qkv.shape[1] // 3
    a = AttnArgs()
    a.qkv, a.out, a.lse = qkv.data_ptr(), out.data_ptr(), lse.data_ptr()
    a.ld_qkv, a.ld_o = _ld(qkv), _ld(out)
    a.seq_outer, a.seq_inner, a.tok_stride = lay.seq_outer, lay.seq_inner, lay.tok_stride
    a.D, a.H, a.N, a.head_dim = D, H, lay.N, D // H
    a.n_seq, a.s_in = lay.n_seq, lay.s_in
    a.causal, a.mask_div = int(causal), mask_div
    if key_mask is not None:
        assert key_mask.dtype == torch.int64 and key_mask.is_contiguous() and key_mask.shape[-1] == lay.N
        a.key_mask = key_mask.data_ptr()
    if mask_rows is not None:
        assert mask_rows.dtype == torch.int32
        a.mask_rows = mask_rows.data_ptr()
    return a


def attention_fwd(qkv, lay, H, *, causal=False, key_mask=None, mask_rows=None, mask_div=1, out=None):
    """qkv bf16 [rows, 3D] (q pre-scaled) -> (out bf16 [rows, D], lse f32 [n_seq, H, N]).  `qkv` / `out` may be
    column views of wider row-major buffers (pitch a multiple of 8 elements)."""
    assert qkv.dtype == BF16 and qkv.is_cuda
    D = qkv.shape[1] // 3
    if out is None:
        out = torch.empty((qkv.shape[0], D), device=qkv.device, dtype=BF16)
    assert out.dtype == BF16 and out.shape == (qkv.shape[0], D)
    lse = torch.empty((lay.n_seq, H, lay.N), device=qkv.device, dtype=F32)
    a = _attn_args(qkv, out, lse, lay, H, causal, key_mask, mask_rows, mask_div)
    check(lib().missm_attention_fwd(ctypes.byref(a), stream_ptr()), "attention_fwd")
    return out, lse


def attention_bwd(qkv, out, lse, d_out, lay, H, q_scale, *, causal=False, key_mask=None,
                  mask_rows=None, mask_div=1, dqkv_out=None, want_colsum=True):
    """-> (dqkv bf16 [rows, 3D], column sums of dqkv f32 [3D] = the q/k/v bias gradients); the q block is
    the gradient w.r.t. the UN-scaled projection.  (Emitting the column sums from the kernels' epilogues
    was measured slower -- contended atomics -- so one extra HBM-bound pass over dqkv computes them.)"""
    assert d_out.dtype == BF16 and d_out.shape == out.shape and _ld(d_out) == _ld(out)
    if dqkv_out is None:
        dqkv = torch.empty(qkv.shape, device=qkv.device, dtype=BF16) if _ld(qkv) == qkv.shape[1] else None
        assert dqkv is not None, "a strided qkv view needs dqkv_out with the same pitch"
    else:
        dqkv = dqkv_out
        assert dqkv.dtype == BF16 and dqkv.shape == qkv.shape and _ld(dqkv) == _ld(qkv)
    delta = torch.empty_like(lse)
    a = _attn_args(qkv, out, lse, lay, H, causal, key_mask, mask_rows, mask_div)
    a.d_out, a.delta, a.dqkv, a.q_scale = d_out.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), q_scale
    csum = None
    a.colsum_done = 0
    check(lib().missm_attention_bwd(ctypes.byref(a), stream_ptr()), "attention_bwd")
    if want_colsum and not a.colsum_done:
        csum = colsum(dqkv)
    return dqkv, csum


# ---------------------------------------------------------------------------------- layernorm
def layernorm_fwd(x, gamma, beta, eps, *, out_dtype=BF16, row_index=None, n_rows=None,
                  add_rows=None, add_period=0, add_div=0, x_out=None, want_stats=True, out=None):
    """x f32 [R, D] -> (y [M, D], mean [M], rstd [M]); M = len(row_index) or R.  `out`: write y into this
    row-major (possibly wider-pitched) [M, D] view instead of a fresh tensor."""
    assert x.dtype == F32 and x.is_cuda
    D = x.shape[1]
    M = n_rows if n_rows is not None else (row_index.numel() if row_index is not None else x.shape[0])
    if out is not None:
        assert out.shape == (M, D) and out.dtype == out_dtype
        y = out
    else:
        y = torch.empty((M, D), device=x.device, dtype=out_dtype)
    mean = torch.empty((M,), device=x.device, dtype=F32) if want_stats else None
    rstd = torch.empty((M,), device=x.device, dtype=F32) if want_stats else None
    check(lib().missm_layernorm_fwd(_p(x), _ld(x), _p(row_index), _p(add_rows), add_period, add_div,
                                    _p(x_out if x_out is not None else x) if add_rows is not None else None,
                                    _p(gamma), _p(beta),
                                    _p(y), _ld(y), int(out_dtype == BF16), _p(mean), _p(rstd), M, D,
                                    eps, stream_ptr()), "layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, *, dres=None, row_index=None, dx=None,
                  want_bf16=False, dx_bf16=None):
    """-> (dx f32 like x, dx_bf16 or None, dgamma [D], dbeta [D], dx_colsum [D]).  With row_index, dx
    must be a pre-zeroed full-size tensor and only the indexed rows are written.  dx_colsum = column
    sums of the written rows of dx: the bias gradient of the Linear whose output this stream is."""
    D = x.shape[1]
    M = dy.shape[0]
    if dx is None:
        assert row_index is None
        dx = torch.empty_like(x)
    if want_bf16 and dx_bf16 is None:
        dx_bf16 = torch.empty(x.shape, device=x.device, dtype=BF16)
    nparts = lib().missm_ln_bwd_num_partials(M)
    partial = torch.empty((nparts, 3, D), device=x.device, dtype=F32)
    dgb = torch.empty((3, D), device=x.device, dtype=F32)
    dgamma, dbeta, dcol = dgb[0], dgb[1], dgb[2]
    check(lib().missm_layernorm_bwd(_p(dy), _ld(dy), int(dy.dtype == BF16), _p(x), _ld(x),
                                    _p(row_index), _p(mean), _p(rstd), _p(gamma), _p(dres), _p(dx),
                                    _p(dx_bf16), _p(partial), _p(dgamma), _p(dbeta), _p(dcol), M, D,
                                    stream_ptr()), "layernorm_bwd")
    return dx, dx_bf16, dgamma, dbeta, dcol


# ------------------------------------------------------------------------------------ helpers
def cast_bf16(src, out=None, cols_dst=None):
    """fp32 [rows, cols] -> bf16 [rows, cols_dst] (zero padded)."""
    assert src.dtype == F32 and src.dim() == 2 and src.stride(1) == 1
    rows, cols = src.shape
    cols_dst = cols if cols_dst is None else cols_dst
    if out is None:
        out = torch.empty((rows, cols_dst), device=src.device, dtype=BF16)
    check(lib().missm_cast_f32_bf16(_p(src), _ld(src), _p(out), _ld(out), rows, cols, cols_dst,
                                    stream_ptr()), "cast_f32_bf16")
    return out


def colsum(x):
    """bf16 [M, N] -> f32 [N] column sums (bias gradient)."""
    assert x.dtype == BF16
    M, N = x.shape
    R = lib().missm_colsum_num_partials(M)
    partial = torch.empty((R, N), device=x.device, dtype=F32)
    out = torch.empty((N,), device=x.device, dtype=F32)
    check(lib().missm_colsum_bf16(_p(x), _ld(x), M, N, _p(partial), _p(out), stream_ptr()), "colsum")
    return out


def patchify(pixels, ps, Kpad, T=1, sample_index=None, n_samples=None):
    """pixels f32 [*, C, H, W] (T = 1) or [*, C, T, H, W] -> bf16 [Bn * T * gh * gw, Kpad] patches of
    the present samples."""
    assert pixels.dtype == F32 and pixels.is_contiguous()
    assert pixels.dim() == (4 if T == 1 else 5)
    C, H, W = pixels.shape[1], pixels.shape[-2], pixels.shape[-1]
    Bn = n_samples if n_samples is not None else pixels.shape[0]
    out = torch.empty((Bn * T * (H // ps) * (W // ps), Kpad), device=pixels.device, dtype=BF16)
    check(lib().missm_patchify(_p(pixels), _p(sample_index), _p(out), Bn, C, T, H, W, ps, Kpad,
                               stream_ptr()), "patchify")
    return out


def colsum_grouped(x, period, div):
    """x f32 [M, D] -> f32 [period, D]: sums of the rows r with (r // div) % period == g."""
    assert x.dtype == F32 and x.is_contiguous()
    M, D = x.shape
    out = torch.empty((period, D), device=x.device, dtype=F32)
    check(lib().missm_colsum_grouped_f32(_p(x), M, D, period, div, _p(out), stream_ptr()), "colsum_grouped")
    return out


def copy_f32(src, dst):
    assert src.dtype == F32 and dst.dtype == F32 and src.is_contiguous() and dst.is_contiguous()
    assert src.numel() == dst.numel()
    check(lib().missm_copy_f32(_p(src), _p(dst), src.numel(), stream_ptr()), "copy_f32")
    return dst


def patch_embed_implicit(pixels, w_bf16, pos, tok, ps, T=1, sample_index=None, n_samples=None):
    """Patch-embedding conv as an implicit GEMM (csrc/patch_embed_tc.cu): fills the patch rows of `tok`
    f32 [n_img * (P + 1), D] from pixels f32 [*, C, (T,) H, W], w_bf16 [D, Kpad] and pos f32 [P + 1, D].
    Returns False if the library declines the shape (caller: patchify + gemm)."""
    assert pixels.dtype == F32 and pixels.is_contiguous() and tok.dtype == F32 and pos.dtype == F32
    assert pixels.dim() == (4 if T == 1 else 5)
    C, H, W = pixels.shape[1], pixels.shape[-2], pixels.shape[-1]
    Bn = n_samples if n_samples is not None else pixels.shape[0]
    D, Kpad = w_bf16.shape
    from . import _lib
    _lib.CALLS[0] += 1
    rc = lib().missm_patch_embed_implicit(_p(pixels), _p(sample_index), _p(w_bf16), _p(pos), _p(tok), Bn, C, T, H, W,
                                          ps, Kpad, D, stream_ptr())
    if rc == -1:
        return False
    if rc != 0:
        check(rc, "patch_embed_implicit")
    return True


def cls_rows(cls, pos, tok, Bn, ntok):
    check(lib().missm_cls_rows(_p(cls), _p(pos), _p(tok), Bn, ntok, tok.shape[1], stream_ptr()), "cls_rows")


def embed_bwd(dtok, Bn, ntok):
    D = dtok.shape[1]
    dpos = torch.empty((ntok, D), device=dtok.device, dtype=F32)
    dpatch = torch.empty((Bn * (ntok - 1), D), device=dtok.device, dtype=BF16)
    check(lib().missm_embed_bwd(_p(dtok), _p(dpos), _p(dpatch), Bn, ntok, D, stream_ptr()), "embed_bwd")
    return dpos, dpatch


def frame_mean(x, Bn, T, out_dtype=BF16):
    D = x.shape[1]
    out = torch.empty((Bn, D), device=x.device, dtype=out_dtype)
    check(lib().missm_frame_mean(_p(x), _p(out), int(out_dtype == BF16), Bn, T, D, stream_ptr()), "frame_mean")
    return out


def frame_mean_bwd(dout, Bn, T):
    D = dout.shape[1]
    din = torch.empty((Bn * T, D), device=dout.device, dtype=F32)
    check(lib().missm_frame_mean_bwd(_p(dout), _p(din), Bn, T, D, stream_ptr()), "frame_mean_bwd")
    return din


def l2norm_scale_fwd(x, scale):
    Bn, P = x.shape
    y = torch.empty_like(x)
    inv = torch.empty((Bn,), device=x.device, dtype=F32)
    check(lib().missm_l2norm_scale_fwd(_p(x), _p(y), _p(inv), scale, Bn, P, stream_ptr()), "l2norm_fwd")
    return y, inv


def l2norm_scale_bwd(dy, x, inv, scale, out_dtype=BF16):
    Bn, P = x.shape
    dx = torch.empty((Bn, P), device=x.device, dtype=out_dtype)
    check(lib().missm_l2norm_scale_bwd(_p(dy), _p(x), _p(inv), scale, _p(dx), int(out_dtype == BF16),
                                       Bn, P, stream_ptr()), "l2norm_bwd")
    return dx


def text_embed_fwd(ids, tok_emb, pos_emb, sample_index=None, n_samples=None):
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    L = ids.shape[1]
    Bn = n_samples if n_samples is not None else ids.shape[0]
    D = tok_emb.shape[1]
    out = torch.empty((Bn * L, D), device=ids.device, dtype=F32)
    check(lib().missm_text_embed_fwd(_p(ids), _p(sample_index), _p(tok_emb), _p(pos_emb), _p(out), Bn,
                                     L, D, stream_ptr()), "text_embed_fwd")
    return out


def text_embed_bwd(ids, dx, vocab, sample_index=None, n_samples=None):
    L = ids.shape[1]
    Bn = n_samples if n_samples is not None else ids.shape[0]
    D = dx.shape[1]
    dtok = torch.zeros((vocab, D), device=dx.device, dtype=F32)
    dpos = torch.empty((L, D), device=dx.device, dtype=F32)
    check(lib().missm_text_embed_bwd(_p(ids), _p(sample_index), _p(dx), _p(dtok), _p(dpos), Bn, L, D,
                                     stream_ptr()), "text_embed_bwd")
    return dtok, dpos


def argmax_rows(ids, sample_index=None, n_samples=None):
    L = ids.shape[1]
    Bn = n_samples if n_samples is not None else ids.shape[0]
    out = torch.empty((Bn,), device=ids.device, dtype=torch.int32)
    check(lib().missm_argmax_rows(_p(ids), _p(sample_index), _p(out), Bn, L, stream_ptr()), "argmax_rows")
    return out


# --------------------------------------------------------------------------------- compaction
def compact_mask(missing_index, codes):
    """missing_index int64 [B] (CUDA) -> (present_idx int32 [T, B], slot_of int32 [T, B],
    counts int32 [T]) for the T towers whose missing codes are `codes`."""
    assert missing_index.dtype == torch.int64 and missing_index.is_cuda and missing_index.is_contiguous()
    B, T = missing_index.numel(), len(codes)
    dev = missing_index.device
    idx = torch.empty((T, B), device=dev, dtype=torch.int32)
    slot = torch.empty((T, B), device=dev, dtype=torch.int32)
    counts = torch.empty((T,), device=dev, dtype=torch.int32)
    codes_host = (ctypes.c_int32 * T)(*[int(c) for c in codes])
    check(lib().missm_compact_mask(_p(missing_index), B, ctypes.cast(codes_host, ctypes.c_void_p), T,
                                   _p(idx), _p(slot), _p(counts), stream_ptr()), "compact_mask")
    return idx, slot, counts


def scatter_rows_zero(src, slot_of, B):
    P = src.shape[1]
    dst = torch.empty((B, P), device=slot_of.device, dtype=F32)
    check(lib().missm_scatter_rows_zero(_p(src), _p(slot_of), _p(dst), B, P, stream_ptr()), "scatter_rows_zero")
    return dst


def gather_rows(src, idx, n_rows):
    """dst[r] = src[idx[r]] for 2-D contiguous src (any dtype, row bytes multiple of 16)."""
    assert src.is_contiguous()
    row_bytes = src.stride(0) * src.element_size()
    dst = torch.empty((n_rows,) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
    check(lib().missm_gather_rows(_p(src), _p(idx), _p(dst), n_rows, row_bytes, stream_ptr()), "gather_rows")
    return dst


# ------------------------------------------------------------------------------ input pipeline
class PreprocArgs(ctypes.Structure):
    """Mirror of `missm_preproc_args` (include/missm_b200.h)."""
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p),
                ("src_f32", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("S", ctypes.c_int32),
                ("antialias", ctypes.c_int32),
                ("pre_div", ctypes.c_float), ("clip_lo", ctypes.c_float), ("clip_hi", ctypes.c_float),
                ("post_div", ctypes.c_float), ("mean", ctypes.c_float * 3), ("std_", ctypes.c_float * 3)]


def image_preprocess(src, out, S, mean, std, *, antialias, pre_div=255.0, clip_lo=float("-inf"),
                     clip_hi=float("inf"), post_div=1.0):
    """src: CUDA uint8 [H, W, 3] or float32 [H, W] (contiguous) -> out: CUDA float32 [3, S, S] (a slice of the batch
    tensor): divide / clip / divide, bicubic resize of the shorter side to S, center crop, normalise."""
    assert src.is_cuda and src.is_contiguous() and out.is_cuda and out.is_contiguous()
    assert out.dtype == F32 and tuple(out.shape) == (3, S, S)
    f32 = src.dtype == F32
    assert (f32 and src.dim() == 2) or (src.dtype == torch.uint8 and src.dim() == 3 and src.shape[2] == 3), \
        "expected uint8 [H, W, 3] or float32 [H, W]"
    a = PreprocArgs()
    a.src, a.dst, a.src_f32 = src.data_ptr(), out.data_ptr(), int(f32)
    a.H, a.W, a.S, a.antialias = src.shape[0], src.shape[1], S, int(bool(antialias))
    a.pre_div, a.clip_lo, a.clip_hi, a.post_div = pre_div, clip_lo, clip_hi, post_div
    a.mean, a.std_ = (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std)
    check(lib().missm_image_preprocess(ctypes.byref(a), stream_ptr()), "image_preprocess")
    return out


class VideoArgs(ctypes.Structure):
    """Mirror of `missm_video_args`."""
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("T", ctypes.c_int32), ("H", ctypes.c_int32),
                ("W", ctypes.c_int32), ("S", ctypes.c_int32), ("hflip", ctypes.c_int32),
                ("mean", ctypes.c_float * 3), ("std_", ctypes.c_float * 3)]


def video_preprocess(frames, out, S, mean, std, hflip=False):
    """frames: CUDA uint8 [T, H, W, 3] (decoded RGB frames of one clip) -> out: CUDA float32 [3, T, S, S]."""
    assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3
    frames = frames.contiguous()
    T, H, W = frames.shape[:3]
    assert out.is_cuda and out.is_contiguous() and out.dtype == F32 and tuple(out.shape) == (3, T, S, S)
    a = VideoArgs()
    a.src, a.dst, a.T, a.H, a.W, a.S, a.hflip = frames.data_ptr(), out.data_ptr(), T, H, W, S, int(bool(hflip))
    a.mean, a.std_ = (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std)
    check(lib().missm_video_preprocess(ctypes.byref(a), stream_ptr()), "video_preprocess")
    return out


class FbankArgs(ctypes.Structure):
    """Mirror of `missm_fbank_args`."""
    _fields_ = [("wave", ctypes.c_void_p), ("wave_all", ctypes.c_void_p), ("n_samples", ctypes.c_int64),
                ("n_total", ctypes.c_int64), ("mel_weights", ctypes.c_void_p), ("mel", ctypes.c_void_p),
                ("wave_sum", ctypes.c_void_p), ("out", ctypes.c_void_p), ("n_mel", ctypes.c_int32),
                ("target", ctypes.c_int32), ("offsets", ctypes.c_int32 * 3), ("mean", ctypes.c_float),
                ("std_", ctypes.c_float)]


def audio_fbank(wave_all, mel_weights, target, offsets, mean, std, out=None):
    """wave_all: CUDA float32 [channels, n] as torchaudio.load returns it (at the model's sample rate) -> CUDA float32
    [3, n_mel, target] (kaldi fbank of channel 0 after subtracting the mean of the whole tensor, chunk offsets
    `offsets` (frames), (x - mean) / (2 std)).  Returns (out, n_frames)."""
    assert wave_all.is_cuda and wave_all.dtype == F32 and wave_all.dim() == 2 and wave_all.is_contiguous()
    n_mel = mel_weights.shape[0]
    assert mel_weights.is_cuda and mel_weights.dtype == F32 and mel_weights.is_contiguous() and mel_weights.shape[1] == 257
    n = wave_all.shape[1]
    nf = int(lib().missm_fbank_num_frames(n))
    if nf <= 0:
        raise ValueError(f"audio of {n} samples is shorter than one 25 ms frame")
    dev = wave_all.device
    mel = torch.empty((nf, n_mel), device=dev, dtype=F32)
    wsum = torch.empty((1,), device=dev, dtype=torch.float64)
    if out is None:
        out = torch.empty((3, n_mel, target), device=dev, dtype=F32)
    a = FbankArgs()
    a.wave, a.wave_all, a.n_samples, a.n_total = wave_all.data_ptr(), wave_all.data_ptr(), n, wave_all.numel()
    a.mel_weights, a.mel, a.wave_sum, a.out = mel_weights.data_ptr(), mel.data_ptr(), wsum.data_ptr(), out.data_ptr()
    a.n_mel, a.target, a.mean, a.std_ = n_mel, target, mean, std
    a.offsets = (ctypes.c_int32 * 3)(*[int(o) for o in offsets])
    check(lib().missm_audio_fbank(ctypes.byref(a), stream_ptr()), "audio_fbank")
    return out, nf


# ------------------------------------------------------------------- fp32 verification mode
# (MISSM_PRECISION=fp32; csrc/fp32_mode.cu)  Not a performance path: every GEMM expands both fp32 operands into
# six bf16 pieces along the contraction dimension and runs ONE tcgen05 bf16 GEMM over K' = 6 K.
def expand6(x, which, stack_rows, cols_pad=None):
    """fp32 [R, C] (row-major, possibly a strided view) -> bf16 [R, 6 * cols_pad] (stack_rows=False) or
    [6 * R, C] (stack_rows=True); which = 0: A-operand piece order, 1: B-operand piece order."""
    assert x.dtype == F32 and x.is_cuda and x.dim() == 2 and x.stride(1) == 1
    R, C = x.shape
    cp = C if cols_pad is None else cols_pad
    out = torch.empty((6 * R, C) if stack_rows else (R, 6 * cp), device=x.device, dtype=BF16)
    check(lib().missm_expand6_bf16(_p(x), _ld(x), R, C, _p(out), _ld(out), cp, which, int(stack_rows), stream_ptr()),
          "expand6_bf16")
    return out


def gemm_f32(a, b, *, a_mn=False, b_mn=False, out=None, bias=None, epilogue=EPI_LINEAR, aux_in=None,
             scale_cols=0, col_scale=1.0, patch_P=0, out_rows=None, split_k=0):
    """fp32-grade C = epilogue(A B^T) with fp32 operands and an fp32 result (see module comment above).
    Layout flags as `gemm`.  Only the LINEAR / RESID / PATCH epilogues (fp32 outputs) are meaningful here."""
    assert a.dtype == F32 and b.dtype == F32
    K = a.shape[0] if a_mn else a.shape[1]
    if not a_mn and not b_mn:
        kp = (K + 7) // 8 * 8          # both K-major: pad every piece alike (TMA needs 16-byte row pitches)
    else:
        assert a_mn and b_mn or K % 8 == 0, "mixed operand layouts need K % 8 == 0"
        kp = K
    ea = expand6(a, 0, a_mn, None if a_mn else kp)
    eb = expand6(b, 1, b_mn, None if b_mn else kp)
    return gemm(ea, eb, a_mn=a_mn, b_mn=b_mn, out=out, out_dtype=F32, bias=bias, epilogue=epilogue, aux_in=aux_in,
                scale_cols=scale_cols, col_scale=col_scale, patch_P=patch_P, out_rows=out_rows, split_k=split_k)


def gelu_f32_fwd(u):
    assert u.dtype == F32 and u.is_contiguous()
    a = torch.empty_like(u)
    check(lib().missm_gelu_f32_fwd(_p(u), _p(a), u.numel(), stream_ptr()), "gelu_f32_fwd")
    return a


def gelu_f32_bwd(d_a, u):
    assert d_a.dtype == F32 and u.dtype == F32 and d_a.is_contiguous() and u.is_contiguous()
    d_u = torch.empty_like(u)
    check(lib().missm_gelu_f32_bwd(_p(d_a), _p(u), _p(d_u), u.numel(), stream_ptr()), "gelu_f32_bwd")
    return d_u


def attention_f32_fwd(qkv, lay, H, *, causal=False, key_mask=None, mask_rows=None, mask_div=1):
    """qkv f32 [rows, 3D] (q pre-scaled) -> (out f32 [rows, D], lse f32 [n_seq, H, N])."""
    assert qkv.dtype == F32 and qkv.is_cuda
    D = qkv.shape[1] // 3
    out = torch.empty((qkv.shape[0], D), device=qkv.device, dtype=F32)
    lse = torch.empty((lay.n_seq, H, lay.N), device=qkv.device, dtype=F32)
    a = _attn_args(qkv, out, lse, lay, H, causal, key_mask, mask_rows, mask_div)
    check(lib().missm_attention_f32_fwd(ctypes.byref(a), stream_ptr()), "attention_f32_fwd")
    return out, lse


def attention_f32_bwd(qkv, out, lse, d_out, lay, H, q_scale, *, causal=False, key_mask=None, mask_rows=None,
                      mask_div=1):
    """-> dqkv f32 [rows, 3D]; the q block is the gradient w.r.t. the UN-scaled projection."""
    assert d_out.dtype == F32 and d_out.shape == out.shape and _ld(d_out) == _ld(out)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    a = _attn_args(qkv, out, lse, lay, H, causal, key_mask, mask_rows, mask_div)
    a.d_out, a.delta, a.dqkv, a.q_scale = d_out.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), q_scale
    check(lib().missm_attention_f32_bwd(ctypes.byref(a), stream_ptr()), "attention_f32_bwd")
    return dqkv


def patchify_f32(pixels, ps, Kpad, T=1, sample_index=None, n_samples=None):
    """pixels f32 -> f32 [Bn * T * gh * gw, Kpad] patches of the present samples (zero padded columns)."""
    assert pixels.dtype == F32 and pixels.is_contiguous()
    C, H, W = pixels.shape[1], pixels.shape[-2], pixels.shape[-1]
    Bn = n_samples if n_samples is not None else pixels.shape[0]
    out = torch.zeros((Bn * T * (H // ps) * (W // ps), Kpad), device=pixels.device, dtype=F32)
    check(lib().missm_patchify_f32(_p(pixels), _p(sample_index), _p(out), Bn, C, T, H, W, ps, Kpad, stream_ptr()),
          "patchify_f32")
    return out


def colsum_f32(x):
    """f32 [M, N] contiguous -> f32 [N] column sums."""
    return colsum_grouped(x, 1, 1)[0]
