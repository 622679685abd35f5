"""The encoder bank: drop-in for `languagebind.LanguageBind` (languagebind/__init__.py:54-85)
with missing-modality mask compaction in front of every tower."""
import contextlib
import os
import time

import torch
from torch import nn

from . import autograd as ag
from . import ops
from . import towers as T
from . import config as C

# src/model/baseline.py:8 -- depth / thermal have no code in the reference; BASELINE.json configs 2
# and 4 use those towers, so the map is extended without touching codes 0-4.
MISSING_TYPE_INDEX = {'language': 1, 'video': 2, 'audio': 3, 'image': 4, 'depth': 5, 'thermal': 6}

# seconds the host has spent blocked in the per-step compaction read-back (it waits for the GPU to drain the previous
# step: bench.py subtracts it from the loop's wall time to report what the host needs to ISSUE a step)
HOST_WAIT_S = [0.0]

config_dict = {
    'thermal': C.LanguageBindThermalConfig, 'image': C.LanguageBindImageConfig,
    'video': C.LanguageBindVideoConfig, 'depth': C.LanguageBindDepthConfig,
    'audio': C.LanguageBindAudioConfig,
}
model_dict = {
    'thermal': T.LanguageBindThermal, 'image': T.LanguageBindImage, 'video': T.LanguageBindVideo,
    'depth': T.LanguageBindDepth, 'audio': T.LanguageBindAudio,
}


def _ddp_backward_sms():
    """Optional policy for data parallelism (MISSM_DDP_SMS=n, default off): during the BACKWARD pass the persistent
    kernels spread over n SMs only, leaving the rest to the all-reduce CTAs DDP launches while the backward is
    still running (a one-CTA-per-SM grid owns every SM's registers and shared memory, so an NCCL CTA can only
    start when a whole GEMM ends, and the next GEMM then runs with late CTAs).  Measured at 2 x B200, ms / step:
    off 117.0-117.2; 132 SMs for the whole step + NCCL_MAX_NCHANNELS=16 113.5; 132 in the backward only 116.5;
    124: 119.0; 140: 119.7; 144 + 4 channels 129.8 (all-reduce too slow).  Within run-to-run noise of one
    another, hence off by default; a dynamic tile scheduler is the real fix."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return 0
    return int(os.environ.get("MISSM_DDP_SMS", "0"))


_CORESIDENT = [None]


def _coresident_policy():
    """The persistent kernels run in their co-resident variants (include/missm_b200.h: missm_set_coresident) when the
    process is one rank of several: DDP's reducer (train_ddp.py:189) then issues a small copy kernel per parameter
    gradient while the backward is running, and those must not wait for the gaps between two one-CTA-per-SM kernels
    (profiles/r03c_*, r03e_* timelines).  Alone on the GPU the full-register variants are ~1.6 % faster."""
    import torch.distributed as dist
    want = bool(dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
    if _CORESIDENT[0] != want:
        ops.set_coresident(want)
        _CORESIDENT[0] = want


def _require_cuda_index(mi, mdev):
    """missing_index on the host is uploaded when the model lives on a CUDA device; there is no CPU path."""
    if not mi.is_cuda:
        if mdev.type != 'cuda':
            raise RuntimeError("missm_b200: missing_index is on the CPU; the B200 path has no CPU fallback "
                               "(move the model and its inputs to a CUDA device)")
        mi = mi.to(mdev, non_blocking=True)
    return mi


class _ZeroTower(torch.autograd.Function):
    """A tower that saw zero present samples on this rank still has to hand DDP a gradient for
    every parameter (find_unused_parameters=False, train_ddp.py:189): emit zero embeddings whose
    backward returns explicit zero gradients."""

    @staticmethod
    def forward(ctx, B, P, device, *params):
        ctx.shapes = [p.shape for p in params]
        ctx.dev = device
        return torch.zeros((B, P), device=device, dtype=torch.float32)

    @staticmethod
    def backward(ctx, d_out):
        return (None, None, None) + tuple(torch.zeros(s, device=ctx.dev, dtype=torch.float32) for s in ctx.shapes)


class LanguageBind(nn.Module):
    supports_compaction = True

    def __init__(self, clip_type, use_temp=True, cache_dir='./cache_dir'):
        super().__init__()
        self.use_temp = use_temp
        self.modality_encoder = {}
        self.modality_proj = {}
        self.modality_scale = {}
        self.modality_config = {}
        model = None
        for k, v in clip_type.items():
            pretrained_ckpt = f'LanguageBind/{v}'
            model = model_dict[k].from_pretrained(pretrained_ckpt, cache_dir=cache_dir)
            self.modality_encoder[k] = model.vision_model
            self.modality_proj[k] = model.visual_projection
            self.modality_scale[k] = model.logit_scale
            self.modality_config[k] = model.config
        # the text tower comes from the LAST loaded model (languagebind/__init__.py:69-70)
        self.modality_encoder['language'] = model.text_model
        self.modality_proj['language'] = model.text_projection
        self.modality_encoder = nn.ModuleDict(self.modality_encoder)
        self.modality_proj = nn.ModuleDict(self.modality_proj)
        self.compaction = os.environ.get("MISSM_COMPACTION", "1") != "0"
        self.tower_streams = os.environ.get("MISSM_TOWER_STREAMS", "1") != "0"
        self.lockstep = os.environ.get("MISSM_LOCKSTEP", "1") != "0"

    @classmethod
    def from_models(cls, models, use_temp=True):
        """Build a bank from already constructed LanguageBind* models (dict modality -> model)."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.use_temp = use_temp
        enc, proj, self.modality_scale, self.modality_config = {}, {}, {}, {}
        model = None
        for k, model in models.items():
            enc[k], proj[k] = model.vision_model, model.visual_projection
            self.modality_scale[k], self.modality_config[k] = model.logit_scale, model.config
        enc['language'], proj['language'] = model.text_model, model.text_projection
        self.modality_encoder, self.modality_proj = nn.ModuleDict(enc), nn.ModuleDict(proj)
        self.compaction = os.environ.get("MISSM_COMPACTION", "1") != "0"
        self.tower_streams = os.environ.get("MISSM_TOWER_STREAMS", "1") != "0"
        self.lockstep = os.environ.get("MISSM_LOCKSTEP", "1") != "0"
        return self

    def _side_streams(self, n, device):
        """One CUDA stream per tower.  The towers are independent until the fusion head, and every hot
        kernel is a persistent one-CTA-per-SM grid whose last wave leaves SMs idle (e.g. 236 pair tiles on
        74 CTA pairs = 3.2 waves); with the towers on separate streams the next tower's CTAs fill those
        SMs.  autograd replays each tower's backward on the stream its forward ran on."""
        pool = self.__dict__.setdefault('_stream_pool', {})
        key = (device.index, n)
        if key not in pool:
            pool[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
        return pool[key]

    def _param_device(self):
        p = next(self.parameters(), None)
        return p.device if p is not None else torch.device('cpu')

    def _scale(self, key):
        if self.use_temp and key != 'language':
            return float(self.modality_scale[key].detach().exp())
        return 1.0

    def forward(self, inputs, missing_index=None):
        """inputs: {modal: {'pixel_values': ...} | {'input_ids', 'attention_mask'}} -> {modal: [B, P]}.
        `missing_index` (int64 [B], optional) enables compaction: a tower only runs the samples whose
        code differs from its own; rows of missing samples come back as zeros."""
        mdev = self._param_device()
        if mdev.type == 'cuda':
            _coresident_policy()
        policy = _ddp_backward_sms()
        if policy:
            ops.set_persistent_sms(0)          # forward (also the no_grad evaluation after an epoch): all SMs
        ddp_sms = policy if torch.is_grad_enabled() else 0
        keys = list(inputs.keys())
        plan = {}
        # HOST inputs are accepted when the model lives on a CUDA device: every tower uploads its own tensors on
        # its own stream (pinned memory -> asynchronous), so the copy of tower k+1 overlaps the compute of tower k.
        # The caller must not overwrite a pinned buffer before the step's result has been read back.
        mdev = self._param_device()
        if missing_index is not None and self.compaction and len(keys) > 0:
            mi = missing_index.reshape(-1).to(torch.int64).contiguous()
            mi_host = mi if not mi.is_cuda else None
            mi = _require_cuda_index(mi, mdev)
            codes = [MISSING_TYPE_INDEX.get(k, -1) for k in keys]
            idx, slot, counts = ops.compact_mask(mi, codes)
            if mi_host is not None:
                # the index came from the host: the towers' batch sizes are counted there, no device read-back
                counts = [int((mi_host != c).sum()) for c in codes]
            else:
                t0 = time.perf_counter()
                counts = counts.tolist()      # the one host sync of the step: sizes of the towers' batches
                HOST_WAIT_S[0] += time.perf_counter() - t0
            B = mi.numel()
            for i, k in enumerate(keys):
                if counts[i] < B:
                    plan[k] = (idx[i], slot[i], counts[i], B)
        outputs = {}
        dev = mdev if mdev.type == 'cuda' else None
        for v in inputs.values():
            for t in v.values():
                if torch.is_tensor(t) and t.is_cuda:
                    dev = t.device
        use_streams = self.tower_streams and len(keys) > 1 and dev is not None
        main = torch.cuda.current_stream(dev) if use_streams else None
        streams = self._side_streams(len(keys), dev) if use_streams else None
        # Lockstep issue (MISSM_LOCKSTEP=0 turns it off): the towers advance one encoder layer at a time in round-robin
        # order, each on its own stream.  The GPU work is the same, but (a) every stream has work from the first
        # microseconds of the step instead of after the host has issued the whole previous tower, and (b) autograd
        # replays the backward in the reverse of this order -- layer 23 of every tower, then layer 22, ... -- which is
        # the order the gradients really become ready in.  DDP (train_ddp.py:189) rebuilds its buckets after the first
        # step in gradient-arrival order and launches bucket k+1 only after bucket k: with tower-by-tower issue the
        # buckets of the second and third tower queued behind the FIRST layer of the first tower (ready at the very end
        # of the backward), which left two thirds of the all-reduce exposed after the backward.
        lockstep = self.lockstep and len(keys) > 1      # (without streams: same kernels, interleaved on one stream)

        def tower(i, key):
            """Generator: the forward of tower i (yields between layers); returns its [B, P] output."""
            value = inputs[key]
            enc, proj = self.modality_encoder[key], self.modality_proj[key]
            scale = self._scale(key)
            if mdev.type == 'cuda':
                value = {k: (t.to(mdev, non_blocking=True) if torch.is_tensor(t) and not t.is_cuda else t)
                         for k, t in value.items()}
            if key in plan:
                pidx, slot, n, B = plan[key]
                if n == 0:
                    # only parameters that take a gradient (a peft-frozen ViT-L base would otherwise cost
                    # 1.2 GB of zero-filled gradients that autograd throws away)
                    params = [p for p in list(enc.parameters()) + list(proj.parameters()) if p.requires_grad]
                    out = _ZeroTower.apply(B, proj.weight.shape[0], proj.weight.device, *params)
                else:
                    y = (yield from enc.forward_steps(**value, present_idx=pidx, n_present=n, proj=proj,
                                                      scale=scale))[1]
                    out = ag.ScatterZeroFn.apply(y, slot, pidx, n, B)
            else:
                out = (yield from enc.forward_steps(**value, proj=proj, scale=scale))[1]
            if ddp_sms and out.requires_grad:
                out = ag.BackwardSmsFn.apply(out, ddp_sms)
            return out

        def on_stream(i):
            if use_streams:
                return torch.cuda.stream(streams[i])
            return contextlib.nullcontext()

        if use_streams:
            for st in streams:
                st.wait_stream(main)
        gens = {key: tower(i, key) for i, key in enumerate(keys)}
        live = list(enumerate(keys))
        while live:
            still = []
            for i, key in live:
                with on_stream(i):
                    try:
                        if lockstep:
                            next(gens[key])
                            still.append((i, key))
                        else:
                            outputs[key] = T.run_steps(gens[key])
                    except StopIteration as stop:
                        outputs[key] = stop.value
                if key in outputs and use_streams:
                    outputs[key].record_stream(main)
            live = still
        if use_streams:
            for st in streams:
                main.wait_stream(st)
        return outputs


def to_device(x, device):
    return {k: v.to(device) for k, v in x.items()}
