"""missm_b200 -- host-side Python of the B200-native MissM-Benchmark hot path."""
import os as _os

# MISSM_DDP_SMS=n under torchrun (WORLD_SIZE > 1): the backward pass leaves 148 - n SMs to NCCL
# (bank._ddp_backward_sms); more all-reduce channels than that would only queue behind the persistent kernels.
# Must be in the environment before the process group is created; a value set by the user wins.
if int(_os.environ.get("WORLD_SIZE", "1") or "1") > 1:
    _sms = int(_os.environ.get("MISSM_DDP_SMS", "0") or "0")
    if 0 < _sms < 148:
        _os.environ.setdefault("NCCL_MAX_NCHANNELS", str(148 - _sms))
