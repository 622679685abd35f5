"""Host side of the residual-block drivers (include/missm_b200.h: missm_attn_block_*, missm_mlp_block_*).

One binding call issues every kernel of a residual block into two caller-provided arenas (`saved`: forward ->
backward, `scratch`: backward only) plus one flat fp32 gradient buffer, so a ViT-L layer costs 2 calls forward and 2
backward instead of ~27 kernel-level calls with their own output allocations (round 1: the step was host-bound).

The four functions `attn_fwd / attn_bwd / mlp_fwd / mlp_bwd` are the only places that touch the library; everything
above them (autograd.EncoderLayerFn) is tensor bookkeeping.  Reference arithmetic: CLIPEncoderLayer.forward,
languagebind/image/modeling_image.py:105-151 (video: video/modeling_video.py:211-257).
"""
import ctypes

import torch

from ._lib import check, lib, stream_ptr

BF16, F32 = torch.bfloat16, torch.float32
_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float


class AttnBlockArgs(ctypes.Structure):
    """Mirror of `missm_attn_block_args`."""
    _fields_ = [
        ("M", _I), ("D", _I), ("H", _I), ("eps", _F),
        ("seq_outer", _L), ("seq_inner", _L), ("tok_stride", _L),
        ("N", _I), ("n_seq", _I), ("s_in", _I), ("causal", _I), ("mask_div", _I),
        ("key_mask", _P), ("mask_rows", _P),
        ("add_rows", _P), ("add_period", _I), ("add_div", _I),
        ("ln_w", _P), ("ln_b", _P), ("w_qkv", _P), ("b_qkv", _P), ("w_o", _P), ("b_o", _P),
        ("ldw_qkv", _I), ("ldw_o", _I), ("lora_r", _I),
        ("wb_qkv", _P), ("wb_o", _P),
        ("x", _P), ("out", _P), ("saved", _P),
        ("d_out", _P), ("d_out_bf16", _P), ("d_out_colsum_given", _I), ("wgrad", _I),
        ("dx", _P), ("dx_bf16", _P), ("grads", _P), ("scratch", _P),
    ]


class MlpBlockArgs(ctypes.Structure):
    """Mirror of `missm_mlp_block_args`."""
    _fields_ = [
        ("M", _I), ("D", _I), ("F", _I), ("eps", _F),
        ("ln_w", _P), ("ln_b", _P), ("w1", _P), ("b1", _P), ("w2", _P), ("b2", _P),
        ("x", _P), ("out", _P), ("saved", _P),
        ("d_out", _P), ("d_out_bf16", _P), ("d_out_colsum_given", _I), ("wgrad", _I),
        ("dx", _P), ("dx_bf16", _P), ("grads", _P), ("scratch", _P),
    ]


def pad8(n):
    """LoRA rank groups are padded to 64 columns = one 128-byte K block of the GEMM tiles (see include/missm_b200.h)."""
    return (n + 63) // 64 * 64


class AttnWeights:
    """Operands of one attention block as the driver wants them (built by autograd.attn_weights from the module's
    fp32 master parameters; bf16 copies are cached there by parameter version)."""
    __slots__ = ("ln_w", "ln_b", "w_qkv", "b_qkv", "w_o", "b_o", "temb", "lora_r", "wb_qkv", "wb_o")

    def __init__(self, ln_w, ln_b, w_qkv, b_qkv, w_o, b_o, temb=None, lora_r=0, wb_qkv=None, wb_o=None):
        self.ln_w, self.ln_b, self.w_qkv, self.b_qkv, self.w_o, self.b_o = ln_w, ln_b, w_qkv, b_qkv, w_o, b_o
        self.temb, self.lora_r, self.wb_qkv, self.wb_o = temb, lora_r, wb_qkv, wb_o


class MlpWeights:
    __slots__ = ("ln_w", "ln_b", "w1", "b1", "w2", "b2")

    def __init__(self, ln_w, ln_b, w1, b1, w2, b2):
        self.ln_w, self.ln_b, self.w1, self.b1, self.w2, self.b2 = ln_w, ln_b, w1, b1, w2, b2


class _State:
    """What a block keeps between its forward and its backward: the argument struct (pointers) and the tensors
    those pointers refer to."""
    __slots__ = ("args", "keep", "sizes", "meta", "w")

    def __init__(self, args, keep, sizes, meta, w):
        self.args, self.keep, self.sizes, self.meta, self.w = args, keep, sizes, meta, w


_SIZES = {}


def _sizes(kind, key, args):
    hit = _SIZES.get((kind, key))
    if hit is None:
        out = (ctypes.c_int64 * 3)()
        fn = lib().missm_attn_block_sizes if kind == "attn" else lib().missm_mlp_block_sizes
        check(fn(ctypes.byref(args), out), f"{kind}_block_sizes")
        hit = _SIZES[(kind, key)] = (int(out[0]), int(out[1]), int(out[2]))
    return hit


def _arena(nbytes, dev):
    return torch.empty((nbytes,), device=dev, dtype=torch.uint8)


# ----------------------------------------------------------------------------------------- attention block
def attn_fwd(meta, x, w):
    """x f32 [M, D] -> (out f32 [M, D], state).  meta: autograd.AttnMeta; w: AttnWeights."""
    M, D = x.shape
    lay = meta.layout
    a = AttnBlockArgs()
    a.M, a.D, a.H, a.eps = M, D, meta.H, meta.eps
    a.seq_outer, a.seq_inner, a.tok_stride = lay.seq_outer, lay.seq_inner, lay.tok_stride
    a.N, a.n_seq, a.s_in, a.causal, a.mask_div = lay.N, lay.n_seq, lay.s_in, int(meta.causal), meta.mask_div
    if meta.key_mask is not None:
        a.key_mask = meta.key_mask.data_ptr()
    if meta.mask_rows is not None:
        a.mask_rows = meta.mask_rows.data_ptr()
    if w.temb is not None:
        a.add_rows, a.add_period, a.add_div = w.temb.data_ptr(), meta.add_period, meta.add_div
    a.ln_w, a.ln_b = w.ln_w.data_ptr(), w.ln_b.data_ptr()
    a.w_qkv, a.b_qkv, a.w_o, a.b_o = w.w_qkv.data_ptr(), w.b_qkv.data_ptr(), w.w_o.data_ptr(), w.b_o.data_ptr()
    a.ldw_qkv, a.ldw_o, a.lora_r = w.w_qkv.stride(0), w.w_o.stride(0), w.lora_r
    if w.lora_r:
        a.wb_qkv, a.wb_o = w.wb_qkv.data_ptr(), w.wb_o.data_ptr()
    sizes = _sizes("attn", (M, D, meta.H, lay.n_seq, lay.N, w.temb is not None, meta.add_period, w.lora_r), a)
    out = torch.empty_like(x)
    saved = _arena(sizes[0], x.device)
    a.x, a.out, a.saved = x.data_ptr(), out.data_ptr(), saved.data_ptr()
    check(lib().missm_attn_block_fwd(ctypes.byref(a), stream_ptr()), "attn_block_fwd")
    return out, _State(a, (saved, x), sizes, meta, w)


def attn_bwd(st, d_out, d_out_bf16, colsum_given, wgrad):
    """-> (dx f32, dx_bf16, G) with G = dict of fp32 gradient views (names as in the header's `grads` order)."""
    a, w, meta = st.args, st.w, st.meta
    M, D = a.M, a.D
    dev = d_out.device
    dx = torch.empty((M, D), device=dev, dtype=F32)
    dx_b = torch.empty((M, D), device=dev, dtype=BF16)
    grads = torch.empty((st.sizes[2],), device=dev, dtype=F32)
    scratch = _arena(st.sizes[1], dev)
    a.d_out = d_out.data_ptr()
    a.d_out_bf16 = d_out_bf16.data_ptr() if d_out_bf16 is not None else None
    a.d_out_colsum_given, a.wgrad = int(colsum_given), int(wgrad)
    a.dx, a.dx_bf16, a.grads, a.scratch = dx.data_ptr(), dx_b.data_ptr(), grads.data_ptr(), scratch.data_ptr()
    check(lib().missm_attn_block_bwd(ctypes.byref(a), stream_ptr()), "attn_block_bwd")
    names = [("ln_w", (D,)), ("ln_b", (D,)), ("dx_colsum", (D,)), ("w_qkv", (3 * D, D)), ("b_qkv", (3 * D,)),
             ("w_o", (D, D)), ("b_o", (D,))]
    if w.temb is not None:
        names.append(("temb", (meta.add_period, D)))
    if w.lora_r:
        R3, R1 = pad8(3 * w.lora_r), pad8(w.lora_r)
        names += [("a_cat", (R3, D)), ("sb_cat", (3 * D, R3)), ("a_o", (R1, D)), ("sb_o", (D, R1))]
    return dx, dx_b, _views(grads, names)


# ----------------------------------------------------------------------------------------------- MLP block
def mlp_fwd(eps, x, w):
    M, D = x.shape
    Fd = w.w1.shape[0]
    a = MlpBlockArgs()
    a.M, a.D, a.F, a.eps = M, D, Fd, eps
    a.ln_w, a.ln_b = w.ln_w.data_ptr(), w.ln_b.data_ptr()
    a.w1, a.b1, a.w2, a.b2 = w.w1.data_ptr(), w.b1.data_ptr(), w.w2.data_ptr(), w.b2.data_ptr()
    sizes = _sizes("mlp", (M, D, Fd), a)
    out = torch.empty_like(x)
    saved = _arena(sizes[0], x.device)
    a.x, a.out, a.saved = x.data_ptr(), out.data_ptr(), saved.data_ptr()
    check(lib().missm_mlp_block_fwd(ctypes.byref(a), stream_ptr()), "mlp_block_fwd")
    return out, _State(a, (saved, x), sizes, None, w)


def mlp_bwd(st, d_out, d_out_bf16, colsum_given, wgrad):
    a = st.args
    M, D, Fd = a.M, a.D, a.F
    dev = d_out.device
    dx = torch.empty((M, D), device=dev, dtype=F32)
    dx_b = torch.empty((M, D), device=dev, dtype=BF16)
    grads = torch.empty((st.sizes[2],), device=dev, dtype=F32)
    scratch = _arena(st.sizes[1], dev)
    a.d_out = d_out.data_ptr()
    a.d_out_bf16 = d_out_bf16.data_ptr() if d_out_bf16 is not None else None
    a.d_out_colsum_given, a.wgrad = int(colsum_given), int(wgrad)
    a.dx, a.dx_bf16, a.grads, a.scratch = dx.data_ptr(), dx_b.data_ptr(), grads.data_ptr(), scratch.data_ptr()
    check(lib().missm_mlp_block_bwd(ctypes.byref(a), stream_ptr()), "mlp_block_bwd")
    names = [("ln_w", (D,)), ("ln_b", (D,)), ("dx_colsum", (D,)), ("w1", (Fd, D)), ("b1", (Fd,)), ("w2", (D, Fd)),
             ("b2", (D,))]
    return dx, dx_b, _views(grads, names)


def _views(flat, names):
    out, o = {}, 0
    for n, shape in names:
        cnt = 1
        for s in shape:
            cnt *= s
        out[n] = flat[o:o + cnt].view(shape)
        o += cnt
    assert o == flat.numel(), (o, flat.numel())
    return out
