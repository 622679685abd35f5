"""The workloads bench.py is paid on, at FULL width and depth, against the CPU oracle WITH gradients.

* BASELINE.json configs[1]: 3 x ViT-L/14 x 24 layers (image + depth + thermal) + `sum` head, B = 4 with missing codes
  {0, 4, 5, 6} (every tower drops one sample through mask compaction), fwd + bwd: loss, the three embeddings and 18
  gradient tensors -- first / middle / last layer, patch, position and class embeddings of every tower, the head.
* BASELINE.json configs[2]: audio (N = 593) + 8-frame video (temporal attention) towers at full width, 2 layers,
  fwd + bwd with gradients (temporal attention, temporal embedding, resized position table, ...).

The oracle (oracle/restatement.py, fp32 on the host cores) is the checker; it finishes these sizes in seconds.
Reference call sites: src/model/baseline.py:450-453, train_ddp.py:249-253."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import restatement as R  # noqa: E402  (the checker, never the thing measured)
from test_parity_gpu import DEV, TOL, TOL_GRAD, TOL_LOGIT, make, rel, to_dev  # noqa: E402


# bf16 mode, gradients of the 24-layer towers.  With name-seeded random weights the 24-layer towers map every input
# to nearly the same embedding (cosine 0.96 between samples), so the loss gradient lives in small differences and is
# ILL-CONDITIONED with respect to the forward tolerance: perturbing the oracle's OWN embeddings by 5e-3 (half of
# north_star's 1e-2; the bf16 path measures 5e-3) moves the oracle's own head / d_embedding gradients by 7-12 %
# (scratch/grad_conditioning.py, ReLU / LayerNorm nonlinearity of the head).  So the test bounds three things:
#   (a) fp32 verification mode vs the oracle <= 1e-4 on every tensor        -> the backward formulas are right;
#   (b) bf16 head gradients vs the oracle head evaluated AT THE PRODUCT'S embeddings <= 2e-3 -> nothing but the
#       forward difference explains the head's deviation;
#   (c) bf16 tower gradients vs the oracle: relative error <= 0.25 and cosine >= 0.96 (measured 0.09-0.21).
TOL_GRAD_DEEP = 0.25
COS_GRAD_DEEP = 0.96


def _oracle_step(sd, modal, data, mi, cfgs, tcfg, labels):
    sdg = {k: t.clone().requires_grad_(t.is_floating_point()) for k, t in sd.items()}
    logits, emb = R.finetune_forward(sdg, 'sum', modal, data, mi, cfgs, tcfg, {m: 2.6592 for m in cfgs})
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    return logits.detach(), {m: e.detach() for m, e in emb.items()}, loss.detach(), sdg


def _product_step(model, data, mi, labels):
    model.train()
    emb = {}

    def grab(mod, args, out):
        emb.update({m: e.detach() for m, e in out.items()})
    h = model.encoder.register_forward_hook(grab)
    logits = model(to_dev(data), mi.to(DEV))
    h.remove()
    loss = torch.nn.functional.cross_entropy(logits, labels.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    return logits.detach(), emb, loss.detach()


def _check_grads(model, sdg, names, tol, tag='', cos_min=None, loose=None):
    """`loose`: {substring: tolerance} for tensors with a known cancellation (stated at the call site)."""
    params = dict(model.named_parameters())
    errs, bad = [], []
    for n in names:
        ref = sdg[n].grad
        assert ref is not None and ref.norm() > 0, n
        g = params[n].grad.detach().float().cpu()
        e = rel(g, ref)
        c = torch.nn.functional.cosine_similarity(g.flatten(), ref.flatten(), dim=0).item()
        print(f'   {tag}grad {n}: rel {e:.2e} cos {c:.4f}')
        errs.append((e, n))
        t = tol
        for sub, lt in (loose or {}).items():
            if sub in n:
                t = lt
        if not e < t or (cos_min is not None and not c > cos_min):
            bad.append((e, c, n))
    assert not bad, bad
    return max(errs)


def test_config2_full_depth_fwd_bwd_vs_oracle():
    from missm_b200 import config as C
    modal = ['image', 'depth', 'thermal']
    v = {k: val for k, val in C.VIT_L14.items() if k != 'lora_r'}
    assert v['num_hidden_layers'] == 24 and v['hidden_size'] == 1024
    meta = dict(vision=v, text=dict(C.CLIP_TEXT), projection_dim=768, fusion_dim=256)
    model, cfgs, tcfg, sd = make(meta, modal, 'sum')
    B = 4
    data = R.synth_inputs(modal, B, cfgs, tcfg, seed=41)
    mi = torch.tensor([0, 4, 5, 6])                      # image / depth / thermal each miss one sample
    labels = torch.tensor([0, 1, 2, 1])
    sd = {k: t for k, t in sd.items() if 'language' not in k}
    ref_logits, ref_emb, ref_loss, sdg = _oracle_step(sd, modal, data, mi, cfgs, tcfg, labels)
    logits, emb, loss = _product_step(model, data, mi, labels)
    code = {'image': 4, 'depth': 5, 'thermal': 6}
    for m in modal:
        present = mi != code[m]
        e = rel(emb[m][present.to(DEV)], ref_emb[m][present])
        print(f'config2 full depth: emb {m} rel {e:.2e}')
        assert e < TOL, (m, e)
        assert emb[m][(~present).to(DEV)].abs().max().item() == 0.0      # skipped samples come back as zero rows
    print(f'config2 full depth: logits rel {rel(logits, ref_logits):.2e}, loss {loss.item():.6f} vs {ref_loss.item():.6f}')
    assert rel(logits, ref_logits) < TOL_LOGIT
    assert abs(loss.item() - ref_loss.item()) < TOL * abs(ref_loss.item())
    names = []
    for m in modal:
        p = f'encoder.modality_encoder.{m}.'
        names += [p + 'encoder.layers.0.self_attn.q_proj.weight', p + 'embeddings.patch_embedding.weight',
                  p + 'embeddings.position_embedding.weight', p + 'embeddings.class_embedding',
                  p + 'encoder.layers.11.mlp.fc1.weight', p + 'encoder.layers.23.self_attn.out_proj.bias']
    names += ['fusion.modal_proj.image.weight', 'fusion.head.head.0.weight', 'encoder.modality_proj.depth.weight']
    # (a) fp32 verification mode first: the same 24-layer step with fp32-grade arithmetic must reproduce the oracle's
    #     gradients to 1e-4 -- every backward formula / operand layout of the path is right at full size
    from missm_b200 import autograd as ag
    old = ag.set_precision("fp32")
    try:
        model.zero_grad(set_to_none=True)
        logits32, emb32, loss32 = _product_step(model, data, mi, labels)
        for m in modal:
            present = mi != code[m]
            assert rel(emb32[m][present.to(DEV)], ref_emb[m][present]) < 1e-5, m
        assert abs(loss32.item() - ref_loss.item()) < 1e-5 * abs(ref_loss.item())
        worst32 = _check_grads(model, sdg, names, 1e-4, 'fp32 mode ')
    finally:
        ag.set_precision(old)
    print('config2 full depth, fp32 mode: worst gradient', worst32)
    # (b) the product's bf16 mode: head gradients against the oracle head evaluated at the product's OWN embeddings
    model.zero_grad(set_to_none=True)
    _, emb_p, _ = _product_step(model, data, mi, labels)
    hs = {k: t.detach().clone().requires_grad_(True) for k, t in sd.items() if k.startswith('fusion.')}
    lg = R.fusion_forward(hs, 'sum', modal, {m: emb_p[m].float().cpu() for m in modal}, mi)
    torch.nn.functional.cross_entropy(lg, labels).backward()
    params = dict(model.named_parameters())
    for k, t in hs.items():
        e = rel(params[k].grad, t.grad)
        print(f'   head grad at the product embeddings {k}: rel {e:.2e}')
        assert e < 2e-3, (k, e)
    # (c) every tensor against the full oracle
    worst = _check_grads(model, sdg, names, TOL_GRAD_DEEP, cos_min=COS_GRAD_DEEP)
    print('config2 full depth: worst gradient', worst)


def test_config3_audio_video_gradients_vs_oracle():
    from test_configs_gpu import FULL
    modal = ['audio', 'video']
    model, cfgs, tcfg, sd = make(FULL, modal, 'sum')
    B = 3
    data = R.synth_inputs(modal, B, cfgs, tcfg, seed=43)
    mi = torch.tensor([0, 3, 2])                         # sample 1 misses audio (3), sample 2 misses video (2)
    labels = torch.tensor([2, 0, 1])
    sd = {k: t for k, t in sd.items() if 'language' not in k}
    ref_logits, ref_emb, ref_loss, sdg = _oracle_step(sd, modal, data, mi, cfgs, tcfg, labels)
    logits, emb, loss = _product_step(model, data, mi, labels)
    code = {'audio': 3, 'video': 2}
    for m in modal:
        present = mi != code[m]
        e = rel(emb[m][present.to(DEV)], ref_emb[m][present])
        print(f'config3: emb {m} rel {e:.2e}')
        assert e < TOL, (m, e)
    assert rel(logits, ref_logits) < TOL_LOGIT
    assert abs(loss.item() - ref_loss.item()) < TOL * abs(ref_loss.item())
    a, v = 'encoder.modality_encoder.audio.', 'encoder.modality_encoder.video.'
    names = [a + 'embeddings.position_embedding.weight', a + 'embeddings.patch_embedding.weight',
             a + 'encoder.layers.0.self_attn.k_proj.weight', a + 'encoder.layers.1.mlp.fc2.weight',
             v + 'encoder.layers.0.temporal_attn.q_proj.weight', v + 'encoder.layers.0.temporal_embedding',
             v + 'encoder.layers.1.temporal_attn.out_proj.weight', v + 'encoder.layers.0.temporal_layer_norm1.weight',
             v + 'encoder.layers.0.self_attn.v_proj.bias', v + 'embeddings.patch_embedding.weight',
             v + 'embeddings.position_embedding.weight', 'encoder.modality_proj.video.weight',
             'fusion.modal_proj.audio.weight']
    from missm_b200 import autograd as ag
    old = ag.set_precision("fp32")
    try:
        model.zero_grad(set_to_none=True)
        _product_step(model, data, mi, labels)
        worst32 = _check_grads(model, sdg, names, 1e-4, 'fp32 mode ')
    finally:
        ag.set_precision(old)
    print('config3, fp32 mode: worst gradient', worst32)
    model.zero_grad(set_to_none=True)
    _product_step(model, data, mi, labels)
    # temporal attention runs over T = 8 frames of nearly identical synthetic tokens: its softmax is nearly uniform and
    # dS = P o (dP - delta) is a difference of nearly equal numbers -- the q / k projection gradients of that block carry
    # the bf16 rounding of P amplified (measured 0.16; 1e-4 in the fp32 mode above); everything else <= 3e-2
    worst = _check_grads(model, sdg, names, TOL_GRAD, loose={'temporal_attn.q_proj': 0.25, 'temporal_attn.k_proj': 0.25})
    print('config3: worst gradient', worst)
