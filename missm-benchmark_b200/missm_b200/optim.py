"""Drop-in for the optimizer the reference's training script builds (train_ddp.py:205
`optim.Adam(model.parameters(), lr=args.learning_rate, weight_decay=args.weight_decay)`, stepped at :254,
`optimizer.zero_grad()` at :221): SURVEY.md section 8(f) rank 2.

`FusedAdam` has torch.optim.Adam's constructor, `param_groups`, `state` layout (`step`, `exp_avg`, `exp_avg_sq`
per parameter -- state dicts are interchangeable with torch.optim.Adam's) and arithmetic, but `step()` is ONE
launch of `missm_adam_multi` (csrc/optim.cu) per parameter group over a device-resident table of every tensor,
instead of torch's ~10 multi-tensor launches with temporaries: 28 B / parameter of HBM traffic.

The unchanged script reaches it through `install()` (or MISSM_FUSED_ADAM=1 before `import languagebind`), which
rebinds `torch.optim.Adam`; schedulers (`ReduceLROnPlateau`, train_ddp.py:206) keep working because they only
touch `param_groups[i]['lr']`.  No CPU fallback: parameters must be fp32 CUDA tensors.
"""
import ctypes
import math
import os

import torch

from ._lib import check, lib, stream_ptr
from . import autograd as ag
from . import ops

CHUNK_ELEMS = 32768


class AdamArgs(ctypes.Structure):
    """Mirror of `missm_adam_args` (include/missm_b200.h)."""
    _fields_ = [
        ("params", ctypes.c_void_p), ("grads", ctypes.c_void_p), ("exp_avg", ctypes.c_void_p),
        ("exp_avg_sq", ctypes.c_void_p), ("bf16_out", ctypes.c_void_p), ("numel", ctypes.c_void_p),
        ("step_size", ctypes.c_void_p), ("bc2_sqrt", ctypes.c_void_p), ("chunk_tensor", ctypes.c_void_p),
        ("chunk_offset", ctypes.c_void_p), ("chunk_elems", ctypes.c_int64),
        ("n_tensors", ctypes.c_int32), ("n_chunks", ctypes.c_int32),
        ("beta1", ctypes.c_double), ("beta2", ctypes.c_double),
        ("eps", ctypes.c_float), ("weight_decay", ctypes.c_float), ("zero_grads", ctypes.c_int32),
    ]


def chunk_table(numels, chunk_elems=CHUNK_ELEMS):
    """(chunk_tensor, chunk_offset) host lists: tensor t contributes ceil(numel_t / chunk_elems) chunks."""
    tensor_of, offset_of = [], []
    for t, n in enumerate(numels):
        for off in range(0, n, chunk_elems):
            tensor_of.append(t)
            offset_of.append(off)
    return tensor_of, offset_of


def step_scalars(step, lr, beta1, beta2):
    """(lr / bias_correction1, sqrt(bias_correction2)) in double, as torch's _single_tensor_adam computes them."""
    return lr / (1.0 - beta1 ** step), math.sqrt(1.0 - beta2 ** step)


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *,
                 maximize=False, foreach=None, capturable=False, differentiable=False, fused=None):
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if amsgrad or maximize or capturable or differentiable:
            raise NotImplementedError("FusedAdam builds what train_ddp.py:205 uses: amsgrad / maximize / capturable / "
                                      "differentiable are not built")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False)
        super().__init__(params, defaults)
        self._tables = {}
        self.launches = 0
        self.refresh_operands = os.environ.get("MISSM_ADAM_REFRESH", "1") != "0"

    # ------------------------------------------------------------------------------ device tables
    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st['step'] = torch.tensor(0.0, dtype=torch.float32)
            st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables.clear()               # the moments were re-allocated

    def _table(self, gi, group):
        """Static part of a group's table (parameter / moment pointers, sizes, chunk list), validated and built
        once; rebuilt when a parameter was re-allocated (`.to()`), the trainable set changed, or a state dict was
        loaded.  The per-step host work is then one pass over the gradients."""
        plist = [p for p in group['params'] if p.requires_grad or p.grad is not None]
        ptrs = [p.data_ptr() for p in plist]
        hit = self._tables.get(gi)
        if hit is not None and hit['ptrs'] == ptrs:
            return hit
        if not plist:
            return None
        for p in plist:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("missm_b200.optim.FusedAdam: parameters must be contiguous fp32 CUDA tensors "
                                   f"(got {p.dtype} on {p.device}); there is no CPU fallback")
        states = [self._init_state(p) for p in plist]
        dev = plist[0].device
        i64 = lambda v: torch.tensor(v, dtype=torch.int64).to(dev)
        tensor_of, offset_of = chunk_table([p.numel() for p in plist])
        tab = dict(ptrs=ptrs, plist=plist, dev=dev, n=len(plist), n_chunks=len(tensor_of),
                   steps=[float(st['step']) for st in states], step_tensors=[st['step'] for st in states],
                   params=i64(ptrs), exp_avg=i64([st['exp_avg'].data_ptr() for st in states]),
                   exp_avg_sq=i64([st['exp_avg_sq'].data_ptr() for st in states]),
                   numel=i64([p.numel() for p in plist]),
                   chunk_tensor=torch.tensor(tensor_of, dtype=torch.int32).to(dev), chunk_offset=i64(offset_of))
        self._tables[gi] = tab
        return tab

    def _prepare(self, gi, group):
        """Host side of one step of one group -> (AdamArgs, tensors to keep alive until the launch) or None."""
        tab = self._table(gi, group)
        if tab is None:
            return None
        plist, steps = tab['plist'], tab['steps']
        beta1, beta2 = group['betas']
        lr = float(group['lr'])
        gptr, ssz, bc2, keep, scal, stepped, touched = [], [], [], [], {}, [], []
        for i, p in enumerate(plist):
            g = p.grad
            if g is None:                          # torch skips it: no update, its step count does not advance
                gptr.append(0), ssz.append(0.0), bc2.append(1.0)
                continue
            if g.dtype != torch.float32 or g.device != p.device or g.is_sparse:
                raise RuntimeError("FusedAdam: gradients must be dense fp32 tensors on the parameter's device")
            if not g.is_contiguous():
                g = g.contiguous()
                keep.append(g)
            k = steps[i] = steps[i] + 1.0
            sc = scal.get(k)
            if sc is None:                         # almost always ONE distinct step count per group
                sc = scal[k] = step_scalars(k, lr, beta1, beta2)
            gptr.append(g.data_ptr()), ssz.append(sc[0]), bc2.append(sc[1])
            stepped.append(tab['step_tensors'][i])
            touched.append(p)
        if not stepped:
            return None
        torch._foreach_add_(stepped, 1.0)          # the state's own `step` tensors (CPU scalars), one call
        dev = tab['dev']
        # pageable sources: the driver stages them before returning, so the host lists may die right away
        d_g = torch.tensor(gptr, dtype=torch.int64).to(dev, non_blocking=True)
        d_sb = torch.tensor(ssz + bc2, dtype=torch.float32).to(dev, non_blocking=True)
        a = AdamArgs()
        a.params, a.grads = tab['params'].data_ptr(), d_g.data_ptr()
        a.exp_avg, a.exp_avg_sq = tab['exp_avg'].data_ptr(), tab['exp_avg_sq'].data_ptr()
        # bf16 GEMM-operand copies the forward path keeps of these parameters: rewritten by the same pass
        # (saves the 6 B / element cast kernels of the next forward); anything not eligible rebuilds itself
        a.bf16_out, entries, dst = None, [], {}
        if self.refresh_operands:
            dst, entries = ag.plan_operand_refresh(touched)
            if dst:
                d_w = torch.tensor([dst[id(p)].data_ptr() if id(p) in dst else 0 for p in plist],
                                   dtype=torch.int64).to(dev, non_blocking=True)
                a.bf16_out = d_w.data_ptr()
                keep.append(d_w)
        a.numel, a.step_size, a.bc2_sqrt = tab['numel'].data_ptr(), d_sb.data_ptr(), d_sb.data_ptr() + 4 * tab['n']
        a.chunk_tensor, a.chunk_offset = tab['chunk_tensor'].data_ptr(), tab['chunk_offset'].data_ptr()
        a.chunk_elems, a.n_tensors, a.n_chunks = CHUNK_ELEMS, tab['n'], tab['n_chunks']
        a.beta1, a.beta2, a.eps, a.weight_decay = beta1, beta2, group['eps'], group['weight_decay']
        a.zero_grads = 0
        keep += [d_g, d_sb]
        tab['touched'], tab['entries'], tab['rewritten'] = touched, entries, list(dst.values())
        return a, keep, tab

    def _launch(self, a, tab):
        with torch.cuda.device(tab['dev']):
            check(lib().missm_adam_multi(ctypes.byref(a), stream_ptr()), "adam_multi")
        self.launches += 1

    # -------------------------------------------------------------------------------------- step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            prep = self._prepare(gi, group)
            if prep is None:
                continue
            a, keep, tab = prep
            self._launch(a, tab)
            # the kernel wrote through raw pointers: advance the parameters' version counters (host-only, no launch)
            # so that everything keyed on them -- the cached bf16 GEMM-operand copies of autograd.cached_weight,
            # autograd's saved-tensor checks -- sees the update exactly as after an in-place torch op
            # (the rewritten bf16 copies too: a backward over a graph recorded BEFORE this step must fail loudly, as it
            #  does with torch's own in-place update, instead of silently using the new weights)
            touched = tuple(tab['touched']) + tuple(tab['rewritten'])
            torch._C._autograd._unsafe_set_version_counter(touched, tuple(t._version + 1 for t in touched))
            ag.restamp(tab['entries'])
        return loss


_ORIGINAL_ADAM = [None]


def install():
    """Rebind torch.optim.Adam to FusedAdam so that the UNCHANGED train_ddp.py:205 builds it."""
    if _ORIGINAL_ADAM[0] is None:
        _ORIGINAL_ADAM[0] = torch.optim.Adam
        torch.optim.Adam = FusedAdam


def uninstall():
    if _ORIGINAL_ADAM[0] is not None:
        torch.optim.Adam = _ORIGINAL_ADAM[0]
        _ORIGINAL_ADAM[0] = None
