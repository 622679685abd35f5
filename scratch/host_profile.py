"""Host-side cost of ISSUING one training step, measured without a GPU: the C-ABI library is replaced by no-op
entry points (every call returns 0 at once), tensors live on the CPU (torch.empty does not touch the pages), so what
remains is exactly the Python / ctypes / allocator work the product does per step -- the time the host needs before the
GPU could possibly be the limit.  `python scratch/host_profile.py [--profile] [--layers 24] [--batch 64]`."""
import argparse
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "missm-benchmark_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--profile", action="store_true")
ap.add_argument("--layers", type=int, default=24)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--lora-r", type=int, default=0)
a = ap.parse_args()

from missm_b200 import _lib, bank, blocks, fusion_ops, ops, towers  # noqa: E402
import missm_b200.autograd as ag  # noqa: E402


class _Noop:
    def __getattr__(self, name):
        if name == "missm_ln_bwd_num_partials":
            return lambda M: 8
        if name == "missm_colsum_num_partials":
            return lambda M: 8
        if name in ("missm_attn_block_sizes", "missm_mlp_block_sizes"):
            def sizes(args, out):
                a = args._obj
                n = a.M * a.D
                out[0], out[1] = 64 * n, 64 * n
                out[2] = (4 * a.D * a.D + 16 * a.D) if name.startswith("missm_attn") else (2 * a.D * a.F + 8 * a.D + a.F)
                return 0
            return sizes
        return lambda *args: 0


noop = _Noop()
_lib._lib = noop
_lib.lib = lambda: noop
ops.lib = lambda: noop
blocks.lib = lambda: noop
blocks.stream_ptr = lambda: None
fusion_ops.lib = lambda: noop
blocks.lib = lambda: noop
blocks.stream_ptr = lambda: None
ops.stream_ptr = lambda: None
fusion_ops.stream_ptr = lambda: None


def _masked_sum_norm(embs, weights, biases, codes, missing_index, gamma, beta, eps):      # skip the CUDA-only guard
    n = len(embs)
    return fusion_ops._MaskedSumNorm.apply(n, [int(c) for c in codes], missing_index.reshape(-1), float(eps), gamma,
                                           beta, *[e.float() for e in embs], *weights, *biases)


fusion_ops.masked_sum_norm = _masked_sum_norm
towers._require_cuda = lambda t, what: None
bank._require_cuda_index = lambda mi, mdev: mi

import bench  # noqa: E402
import restatement as R  # noqa: E402
from missm_b200 import shapes  # noqa: E402

cfgs, tcfg = bench.full_configs(a.layers, a.lora_r)
with torch.device("cpu"):
    model = shapes.build_finetune(cfgs, tcfg, bench.MODALS, 'sum', 3, 768, 256, dropout_prob=0.1)
for n, p in model.named_parameters():
    if 'language' in n:
        p.requires_grad_(False)
model.train()
B = a.batch
data = {m: {'pixel_values': torch.empty(B, 3, 224, 224)} for m in bench.MODALS}
mi = R.synth_missing_index(B, 0.3, bench.MODALS)
labels = torch.arange(B) % 3
crit = torch.nn.CrossEntropyLoss()
# compact_mask's outputs are read on the host (counts.tolist()): give the no-op library something sensible to return
real_compact = ops.compact_mask


def fake_compact(missing_index, codes):
    Bn, T = missing_index.numel(), len(codes)
    idx = torch.zeros((T, Bn), dtype=torch.int32)
    slot = torch.zeros((T, Bn), dtype=torch.int32)
    counts = torch.tensor([int((missing_index != c).sum()) for c in codes], dtype=torch.int32)
    return idx, slot, counts


ops.compact_mask = fake_compact


def step():
    model.zero_grad(set_to_none=True)
    loss = crit(model(data, mi), labels)
    loss.backward()


step()
t0 = time.perf_counter()
for _ in range(a.steps):
    step()
dt = (time.perf_counter() - t0) / a.steps
print(f"host issue time per step: {dt * 1e3:.1f} ms  (layers {a.layers}, B {B}, lora_r {a.lora_r}, "
      f"{_lib.CALLS[0] // (a.steps + 1)} binding calls)")
if a.profile:
    pr = cProfile.Profile()
    pr.enable()
    step()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
