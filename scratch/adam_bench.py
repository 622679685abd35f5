"""HBM roofline check of the multi-tensor Adam (csrc/optim.cu): image-tower-sized parameter set (303 M fp32
parameters in the tower's own tensor shapes), CUDA events around K steps, algorithmic 28 B / parameter."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "missm-benchmark_b200"))
from missm_b200 import optim  # noqa: E402

D, F, L = 1024, 4096, 24
shapes = []
for _ in range(L):
    shapes += [(D, D)] * 4 + [(D,)] * 4 + [(F, D), (F,), (D, F), (D,)] + [(D,)] * 4
shapes += [(D, 3, 14, 14), (257, D), (D,), (768, D)]
params = [torch.nn.Parameter(torch.randn(s, device="cuda") * 0.02) for s in shapes]
n = sum(p.numel() for p in params)
for p in params:
    p.grad = torch.randn_like(p)
res = {"params": n, "tensors": len(params)}
QUICK = "--quick" in sys.argv          # the ncu target: FusedAdam only, a few steps
for name, cls in ((("fused", optim.FusedAdam),) if QUICK else (("fused", optim.FusedAdam), ("torch_foreach", torch.optim.Adam))):
    opt = cls(params, lr=1e-4)
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    K = 3 if QUICK else 10
    for _ in range(K):
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    res[name] = {"ms_per_step": ms, "host_ms_per_step": (time.perf_counter() - t0) * 1e3 / K,
                 "algorithmic_GBps": 28.0 * n / ms / 1e6}
    if name == "fused":
        # the launch alone (host side prepared once): what the kernel does when the host is not the limit
        a, keep, tab = opt._prepare(0, opt.param_groups[0])
        for _ in range(3):
            opt._launch(a, tab)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(K):
            opt._launch(a, tab)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        res["fused_kernel_only"] = {"ms_per_launch": ms, "algorithmic_GBps": 28.0 * n / ms / 1e6}
    del opt
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(peaks):
    res["measured_peaks"] = json.load(open(peaks))
print(json.dumps(res))
