"""Opt-in integration with torch DDP for the UNCHANGED call of train_ddp.py:189
(`DDP(model, device_ids=[local_rank], broadcast_buffers=True, find_unused_parameters=False)`).

`install()` (run by `import languagebind` when MISSM_DDP_BUCKET_VIEW=1) makes that call default to
`gradient_as_bucket_view=True`: the reducer then no longer copies every parameter's gradient OUT of its all-reduce
buckets after the reduction (1 152 small device-to-device copies, 3.6 GB, issued after the last all-reduce: ~4 ms at
the very end of a three-tower ViT-L step, with nothing left to overlap them with -- profiles/r03c_*timeline*).  `.grad`
tensors then alias the buckets, which the script's `optimizer.zero_grad()` / `optimizer.step()` handle as usual.
It also raises the default `bucket_cap_mb` from torch's 25 to MISSM_DDP_BUCKET_MB (default 200): the persistent
one-CTA-per-SM kernels of the backward leave an all-reduce kernel (32 CTAs that each want a whole SM) only the gaps
between two kernels, so every NCCL launch costs ~0.5 ms of waiting whatever its size; 3.6 GB of gradients in 25 MB
buckets are 112 launches, in 200 MB buckets 19 (measured: 2 x B200 1113 -> 1128-1132, 4 x B200 2210 -> 2242 samples/s).
A keyword the caller passes explicitly always wins.  Rebinding a torch name is not something an import should do
silently, hence the switch."""
import functools
import os

_INSTALLED = [False]


def install():
    if _INSTALLED[0]:
        return
    import torch.nn.parallel as tnp

    ddp_init = tnp.DistributedDataParallel.__init__

    @functools.wraps(ddp_init)
    def init(self, *args, **kwargs):
        kwargs.setdefault("gradient_as_bucket_view", True)
        kwargs.setdefault("bucket_cap_mb", int(os.environ.get("MISSM_DDP_BUCKET_MB", "200")))
        ddp_init(self, *args, **kwargs)

    tnp.DistributedDataParallel.__init__ = init
    _INSTALLED[0] = True
