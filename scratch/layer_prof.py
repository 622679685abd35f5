"""One image tower, 2 layers, fwd+bwd at the bench geometry (58 present samples): per-kernel times per layer."""
import sys, os
ROOT=os.path.join(os.path.dirname(__file__), "..")
sys.path.insert(0, os.path.join(ROOT, "missm-benchmark_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import restatement as R
from missm_b200 import shapes, config as C
v = {k: val for k, val in C.VIT_L14.items() if k != 'lora_r'}; v['num_hidden_layers'] = 2
cfgs = {'image': R.vision_config(**v)}
tcfg = R.text_config(**dict(C.CLIP_TEXT, num_hidden_layers=1))
model = shapes.build_finetune(cfgs, tcfg, ['image'], 'sum', 3, 768, 256, dropout_prob=0.1)
sd = R.synth_state_dict([(k, tuple(t.shape)) for k, t in model.state_dict().items()])
shapes.load_named(model, sd); model = model.cuda().train()
B = 58
data = {'image': {'pixel_values': torch.randn(B, 3, 224, 224, device='cuda')}}
mi = torch.zeros(B, dtype=torch.long, device='cuda'); labels = (torch.arange(B, device='cuda') % 3)
for _ in range(3):
    model.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(model(data, mi), labels); loss.backward()
torch.cuda.synchronize(); print("ok", float(loss))
