"""The "library kernels" bar of SURVEY.md 8(d): the reference arithmetic (oracle/restatement.py, a functional
torch restatement of the reference path -- the reference itself cannot travel to the GPU box) run EAGERLY by
torch on one B200, in fp32 as written (TF32 off and on) and under bf16 autocast, on the bench workload
(image + depth + thermal, `sum` head, fwd+bwd, every tower on the full batch as the reference does).
This is a measurement aid, not a product path; bench.py does not call it."""
import argparse
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "missm-benchmark_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import restatement as R
from missm_b200 import shapes, config as C

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
MODALS = ['image', 'depth', 'thermal']
v = {k: val for k, val in C.VIT_L14.items() if k != 'lora_r'}
cfgs = {m: R.vision_config(**v) for m in MODALS}
tcfg = R.text_config(**dict(C.CLIP_TEXT))
named = shapes.reference_named_shapes(cfgs, tcfg, MODALS, 'sum')
sd = R.synth_state_dict([(k, s) for k, s in named if 'language' not in k])
sd = {k: t.cuda().requires_grad_(t.is_floating_point()) for k, t in sd.items()}
B = a.batch
data = {m: {'pixel_values': x['pixel_values'].cuda()} for m, x in R.synth_inputs(MODALS, B, cfgs, tcfg, seed=0).items()}
mi = R.synth_missing_index(B, 0.3, MODALS).cuda()
labels = (torch.arange(B, device='cuda') % 3)
scales = {m: 2.6592 for m in MODALS}


def step(autocast):
    for t in sd.values():
        t.grad = None
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
        logits, _ = R.finetune_forward(sd, 'sum', MODALS, data, mi, cfgs, tcfg, scales)
    loss = torch.nn.functional.cross_entropy(logits.float(), labels)
    loss.backward()


out = {"batch": B, "steps": a.steps, "workload": "image+depth+thermal ViT-L/14 + sum head, fwd+bwd, full batch per tower"}
for name, tf32, autocast in (("fp32", False, False), ("tf32", True, False), ("bf16_autocast", True, True)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    try:
        for _ in range(2):
            step(autocast)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step(autocast)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out[name] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
    except Exception as exc:  # noqa: BLE001  (an OOM here must not lose the other modes)
        out[name] = {"error": repr(exc)[:200]}
        torch.cuda.empty_cache()
print(json.dumps(out))
