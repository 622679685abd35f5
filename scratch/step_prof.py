"""One fwd+bwd step of the bench workload at reduced depth (image+depth+thermal, B = 64, 30 % missing,
`--layers` encoder layers per tower) bracketed by cudaProfilerStart/Stop, so that
`ncu --profile-from-start off --set full` captures every kernel type of the step exactly as the step
launches it (same shapes, same operands).  Tower streams are off: one launch at a time."""
import argparse
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "missm-benchmark_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import restatement as R
from missm_b200 import shapes, config as C

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=1)
ap.add_argument("--modals", default="image,depth,thermal")
ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
MODALS = a.modals.split(",")
cfgs = {}
for m in MODALS:
    v = {k: val for k, val in C.VIT_L14.items() if k != 'lora_r'}
    v['num_hidden_layers'] = a.layers
    v.update(C.SYNTHETIC_PER_MODALITY[m])
    v['temporal_mlp'] = (m != 'video')
    cfgs[m] = R.vision_config(**v)
tcfg = R.text_config(**dict(C.CLIP_TEXT, num_hidden_layers=1))
model = shapes.build_finetune(cfgs, tcfg, MODALS, 'sum', 3, 768, 256, dropout_prob=0.1)
sd = R.synth_state_dict([(k, tuple(t.shape)) for k, t in model.state_dict().items()])
shapes.load_named(model, sd)
model = model.cuda().train()
model.encoder.tower_streams = False
for n, p in model.named_parameters():
    if 'language' in n:
        p.requires_grad_(False)
B = a.batch
data = {m: {'pixel_values': x['pixel_values'].cuda()} for m, x in R.synth_inputs(MODALS, B, cfgs, tcfg, seed=0).items()}
mi = R.synth_missing_index(B, 0.3, MODALS, seed=2025).cuda()
labels = (torch.arange(B, device='cuda') % 3)


def step():
    model.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(model(data, mi), labels)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(loss.detach()))
