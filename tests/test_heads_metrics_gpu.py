"""SURVEY.md section 8(f) rank 4 on the GPU: the graph fusion heads end to end through the CUDA towers (full batch --
they are not skip-safe) against the CPU oracle with gradients, and the device-resident evaluation metrics against
the reference's own recipe (sklearn on the host, train_ddp.py:117-133)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import restatement as R  # noqa: E402  (the checker, never the thing measured)
from test_parity_gpu import DEV, TOL, TOL_GRAD, TOL_LOGIT, make, rel, to_dev  # noqa: E402

META = dict(vision=dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, patch_size=14,
                        image_size=56),
            text=dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, vocab_size=1000,
                      max_position_embeddings=77),
            per={'video': dict(add_time_attn=True, num_frames=4), 'audio': dict(num_mel_bins=28, target_length=70)},
            projection_dim=768, fusion_dim=256)     # fusion_gcn() hard-codes 768 / 256 (baseline.py:253,293-294)


@pytest.mark.parametrize("fusion", ["graph_fusion", "unified_graph"])
def test_graph_heads_on_gpu_vs_oracle(fusion):
    modal = ['language', 'video', 'audio', 'image']
    model, cfgs, tcfg, sd = make(META, modal, fusion)
    for k in sd:
        if k.endswith('att_l') or k.endswith('att_r'):       # zeros by default: make the attention non-trivial
            sd[k] = R.synth_param(k, sd[k].shape, 0.3)
    from missm_b200 import shapes
    shapes.load_named(model, sd)
    model = model.to(DEV).train()
    assert model.fusion.compaction_safe is False
    B = 6
    data = R.synth_inputs(modal, B, cfgs, tcfg, seed=3)
    mi = torch.tensor([0, 1, 2, 3, 4, 0])
    labels = torch.tensor([0, 1, 2, 0, 1, 2])
    sdg = {k: t.clone().requires_grad_(t.is_floating_point()) for k, t in sd.items()}
    ref, ref_emb = R.finetune_forward(sdg, fusion, modal, data, mi, cfgs, tcfg, {m: 2.6592 for m in cfgs})
    ref_loss = torch.nn.functional.cross_entropy(ref, labels)
    ref_loss.backward()
    out = model(to_dev(data), mi.to(DEV))
    loss = torch.nn.functional.cross_entropy(out, labels.to(DEV))
    loss.backward()
    assert rel(out, ref) < TOL_LOGIT, rel(out, ref)
    assert abs(loss.item() - ref_loss.item()) < TOL * abs(ref_loss.item())
    params = dict(model.named_parameters())
    checked = 0
    for n, t in sdg.items():
        towers = ('encoder.modality_encoder.image.encoder.layers.0.self_attn.q_proj.weight',
                  'encoder.modality_encoder.video.encoder.layers.1.temporal_attn.v_proj.weight',
                  'encoder.modality_encoder.language.encoder.layers.0.mlp.fc1.weight')
        if not (n.startswith('fusion.') or n in towers) or t.grad is None or t.grad.norm() < 1e-8:
            continue
        e = rel(params[n].grad, t.grad)
        print(f'{fusion}: grad {n} rel {e:.2e}')
        # the head is fp32 torch on embeddings that carry the towers' bf16 error (5e-3): node-attention softmax and
        # LeakyReLU amplify it a little more than the plain heads do
        assert e < 5e-2, (n, e)
        checked += 1
    assert checked >= 10, checked


@pytest.mark.parametrize("C", [3, 2, 7])
def test_eval_accumulator_matches_the_scripts_recipe(C):
    from sklearn.metrics import accuracy_score, f1_score, roc_auc_score
    from missm_b200.metrics import EvalAccumulator
    g = torch.Generator().manual_seed(C)
    acc = EvalAccumulator(C, DEV)
    crit = torch.nn.CrossEntropyLoss()
    total_loss, preds, probs, labels_all = 0.0, [], [], []
    for step, B in enumerate((64, 64, 37)):
        logits = torch.randn(B, C, generator=g) * 2
        if step == 0:
            logits[5, :] = 1.25                                   # a tie: torch.argmax takes the first maximum
        labels = torch.randint(0, C, (B,), generator=g)
        acc.update(logits.to(DEV), labels.to(DEV))
        # the script's way (train_ddp.py:105-123)
        total_loss += crit(logits, labels).item()
        preds.extend(torch.argmax(logits, dim=1).numpy())
        probs.extend(torch.softmax(logits, dim=-1).numpy())
        labels_all.extend(labels.numpy())
    m = acc.compute()
    import numpy as np
    probs = np.array(probs)
    want = {'loss': total_loss / 3, 'accuracy': accuracy_score(labels_all, preds),
            'f1': f1_score(labels_all, preds, average='macro'),
            'auc': roc_auc_score(labels_all, probs if C > 2 else probs[:, 1], multi_class='ovo')}
    assert int(acc.n_seen) == 165 and int(acc.confusion.sum()) == 165
    assert m['accuracy'] == pytest.approx(want['accuracy'], abs=1e-12)          # integer counts: exact
    assert m['f1'] == pytest.approx(want['f1'], abs=1e-12)
    assert m['loss'] == pytest.approx(want['loss'], rel=1e-5)
    assert m['auc'] == pytest.approx(want['auc'], abs=1e-6)
