"""ctypes binding of the C-ABI library `lib/libmissm_b200.so` (declared in include/missm_b200.h).

The product path has NO fallback: if the library is missing or a call fails, a RuntimeError is
raised.  `import torch` happens first so that the already-loaded libcudart.so.12 is shared.
"""
import ctypes
import os

import torch  # noqa: F401  (loads libcudart into the process before our library)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MISSM_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "lib", "libmissm_b200.so")   # (override: A/B builds)

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int32
c_float = ctypes.c_float


class AttnArgs(ctypes.Structure):
    """Mirror of `missm_attn_args` (include/missm_b200.h)."""
    _fields_ = [
        ("qkv", ctypes.c_void_p), ("out", ctypes.c_void_p), ("lse", ctypes.c_void_p),
        ("d_out", ctypes.c_void_p), ("delta", ctypes.c_void_p), ("dqkv", ctypes.c_void_p),
        ("key_mask", ctypes.c_void_p), ("mask_rows", ctypes.c_void_p),
        ("ld_qkv", ctypes.c_int64), ("ld_o", ctypes.c_int64),
        ("seq_outer", ctypes.c_int64), ("seq_inner", ctypes.c_int64), ("tok_stride", ctypes.c_int64),
        ("D", ctypes.c_int32), ("H", ctypes.c_int32), ("N", ctypes.c_int32), ("head_dim", ctypes.c_int32),
        ("n_seq", ctypes.c_int32), ("s_in", ctypes.c_int32),
        ("causal", ctypes.c_int32), ("mask_div", ctypes.c_int32),
        ("q_scale", ctypes.c_float),
        ("dqkv_colsum", ctypes.c_void_p), ("colsum_done", ctypes.c_int32),
    ]


class GemmArgs(ctypes.Structure):
    """Mirror of `missm_gemm_args` (include/missm_b200.h)."""
    _fields_ = [
        ("A", c_void_p), ("B", c_void_p), ("C", c_void_p),
        ("bias", c_void_p), ("aux_in", c_void_p), ("aux_out", c_void_p),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("lda", c_int), ("ldb", c_int), ("ldc", c_int),
        ("ld_aux_in", c_int), ("ld_aux_out", c_int),
        ("a_mn", c_int), ("b_mn", c_int),
        ("epilogue", c_int), ("out_f32", c_int),
        ("scale_cols", c_int), ("col_scale", c_float),
        ("patch_P", c_int), ("split_k", c_int), ("force_bn", c_int),
        ("colsum_out", c_void_p), ("colsum_part", c_void_p),
    ]


_lib = None


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"missm_b200: {LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ "
            f"as g; g.build()'` (or `make -C missm-benchmark_b200/csrc`). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    L.missm_version.restype = c_int
    L.missm_last_error.restype = ctypes.c_char_p
    from . import _abi
    if L.missm_version() != _abi.ABI_VERSION:
        # a stale build would read the by-pointer argument structs with another layout: refuse it
        raise RuntimeError(
            f"missm_b200: {LIB_PATH} was built for ABI v{L.missm_version()} but this package binds ABI "
            f"v{_abi.ABI_VERSION}. Rebuild it (`make -C missm-benchmark_b200/csrc` or __graft_entry__.build()).")
    L.missm_launch_count.restype = ctypes.c_int64
    L.missm_launch_count.argtypes = [c_int]
    _declare(L)
    _lib = L
    return L


def _declare(L):
    from . import _abi
    for name, argtypes in _abi.SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = c_int
        fn.argtypes = argtypes


CALLS = [0]      # binding calls made so far (one check() per call): bench.py reports them per step


_DEBUG_EVENTS = os.environ.get("MISSM_DEBUG_EVENTS") == "2"     # debugging aid, see missm_debug_dump
_LABELS = {}


def check(rc, what=""):
    CALLS[0] += 1
    if _DEBUG_EVENTS:
        lab = _LABELS.setdefault(what, what.encode())
        lib().missm_debug_crumb(lab, stream_ptr())
    if rc != 0:
        msg = lib().missm_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"missm_b200 {what} failed (rc={rc}): {msg}")


_RAW_STREAM = [None]        # None: not probed yet; False: unavailable; else torch._C._cuda_getCurrentRawStream


def stream_ptr():
    """Raw cudaStream_t of torch's current stream (0 = legacy default stream).  Called once per kernel launch
    (~2 500 times per training step), so the fast C accessor is used when this torch build has it: it skips the
    torch.cuda.Stream object that `current_stream()` builds (~2-3 us each).  The first call checks it against the
    public API and falls back for good if they disagree."""
    fast = _RAW_STREAM[0]
    if fast:
        return ctypes.c_void_p(fast(torch.cuda.current_device()))
    slow = torch.cuda.current_stream().cuda_stream
    if fast is None:
        f = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        try:
            ok = f is not None and int(f(torch.cuda.current_device())) == int(slow)
        except Exception:
            ok = False
        _RAW_STREAM[0] = f if ok else False
    return ctypes.c_void_p(slow)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
