"""How well conditioned are the gradients the full-depth parity test compares?  CPU only (oracle): perturb the
towers' embeddings by the bf16 path's measured forward error (5e-3 relative, random direction) and look at how much
the ORACLE's own head / projection gradients move.  `python scratch/grad_conditioning.py`"""
import sys, time, torch
sys.path.insert(0, 'oracle'); sys.path.insert(0, 'missm-benchmark_b200')
import restatement as R
from missm_b200 import config as C, shapes
modal = ['image', 'depth', 'thermal']
v = {k: val for k, val in C.VIT_L14.items() if k != 'lora_r'}
cfgs = {m: R.vision_config(**v) for m in modal}; tcfg = R.text_config(**dict(C.CLIP_TEXT))
named = shapes.reference_named_shapes(cfgs, tcfg, modal, 'sum')
sd = R.synth_state_dict([(k, s) for k, s in named if 'language' not in k])
data = R.synth_inputs(modal, 4, cfgs, tcfg, seed=41)
mi = torch.tensor([0, 4, 5, 6]); labels = torch.tensor([0, 1, 2, 1])
t0 = time.time()
with torch.no_grad():
    _, emb = R.finetune_forward(sd, 'sum', modal, data, mi, cfgs, tcfg, {m: 2.6592 for m in cfgs})
print('oracle forward', time.time() - t0, 's')
for m in modal:
    e = emb[m]
    cos = torch.nn.functional.cosine_similarity(e[0:1], e[1:], dim=-1)
    print(m, 'norms', e.norm(dim=-1).tolist(), 'cos(sample0, others)', cos.tolist())

def head_grads(embs):
    hs = {k: t.clone().requires_grad_(True) for k, t in sd.items() if k.startswith('fusion.')}
    es = {m: embs[m].clone().requires_grad_(True) for m in modal}
    logits = R.fusion_forward(hs, 'sum', modal, es, mi)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    g = {k: t.grad for k, t in hs.items()}
    g.update({'d_emb/' + m: es[m].grad for m in modal})
    return loss.item(), g

l0, g0 = head_grads(emb)
torch.manual_seed(0)
for eps in (5e-3, 1e-3):
    pert = {m: e + eps * e.norm(dim=-1, keepdim=True) * torch.nn.functional.normalize(torch.randn_like(e), dim=-1) for m, e in emb.items()}
    l1, g1 = head_grads(pert)
    print(f'--- embeddings perturbed by {eps:g} relative: loss {l0:.6f} -> {l1:.6f}')
    for k in g0:
        print(f'   {k}: rel change {((g1[k] - g0[k]).norm() / g0[k].norm()).item():.3e}')
