"""Pins oracle/restatement.py against outputs of the UNMODIFIED reference (tests/golden/*.pt,
generated in the build container by oracle/make_golden.py through the import shim).  CPU only."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import restatement as R  # noqa: E402

GOLD = os.path.join(HERE, "golden")


def build_problem(meta, modals_with_lang, fusion, n_classes=3):
    """Re-create configs, name-seeded weights and inputs exactly as make_golden.py did, but
    WITHOUT the reference: parameter names/shapes come from the product's module tree."""
    from missm_b200 import shapes
    per = meta.get('per', {})
    vis = [m for m in modals_with_lang if m != 'language']
    cfgs = {}
    for m in vis:
        d = {k: v for k, v in meta['vision'].items() if k != 'lora_r'}
        d.update(per.get(m, {}))
        d['temporal_mlp'] = (m != 'video')
        cfgs[m] = R.vision_config(**d)
    tcfg = R.text_config(**meta['text'])
    named = shapes.reference_named_shapes(cfgs, tcfg, modals_with_lang, fusion,
                                          projection_dim=meta.get('projection_dim', 768),
                                          fusion_dim=meta.get('fusion_dim', 256), n_classes=n_classes)
    sd = R.synth_state_dict(named)
    return cfgs, tcfg, sd


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def tiny():
    return torch.load(os.path.join(GOLD, "tiny_bank.pt"), weights_only=False)


FUSIONS = ['sum', 'concat', 'regression', 'retrieval', 'intra_attention', 'inter_attention',
           'dedicated_dnn', 'Distill_tea', 'self_distill']


@pytest.mark.parametrize("fusion", FUSIONS)
def test_oracle_matches_reference_tiny(tiny, fusion):
    meta = tiny['meta']
    modal_types = ['language'] + meta['modals']
    cfgs, tcfg, sd = build_problem(meta, modal_types, fusion)
    data = R.synth_inputs(modal_types, meta['B'], cfgs, tcfg, seed=0)
    scales = {m: 2.6592 for m in meta['modals']}
    with torch.no_grad():
        res, emb = R.finetune_forward(sd, fusion, modal_types, data, tiny['missing_index'], cfgs, tcfg, scales)
    logits = res[-1] if isinstance(res, tuple) else res
    assert rel(logits, tiny[f'logits/{fusion}']) < 2e-5
    if fusion == 'sum':
        for m in modal_types:
            assert rel(emb[m], tiny[f'emb/{m}']) < 2e-5, m


def test_oracle_gradients_match_reference_tiny(tiny):
    meta = tiny['meta']
    modal_types = ['language'] + meta['modals']
    cfgs, tcfg, sd = build_problem(meta, modal_types, 'sum')
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    data = R.synth_inputs(modal_types, meta['B'], cfgs, tcfg, seed=0)
    scales = {m: 2.6592 for m in meta['modals']}
    logits, _ = R.finetune_forward(sd, 'sum', modal_types, data, tiny['missing_index'], cfgs, tcfg, scales)
    loss = torch.nn.functional.cross_entropy(logits, tiny['labels'])
    assert abs(loss.item() - tiny['loss/sum'].item()) < 1e-5
    loss.backward()
    checked = 0
    for k, v in tiny.items():
        if k.startswith('grad/'):
            g = sd[k[5:]].grad
            # (k_proj.bias has a mathematically zero gradient: softmax is invariant to it)
            assert g is not None and (g - v).norm() <= 5e-4 * v.norm() + 1e-8, k
            checked += 1
    assert checked >= 15
    for name, gn in tiny['grad_norms'].items():
        if gn is not None and name in sd and sd[name].grad is not None and gn > 1e-8:
            assert abs(sd[name].grad.norm().item() - gn) / gn < 2e-3, name


def test_resize_pos_matches_reference(tiny):
    out = R.resize_pos(tiny['resize_pos/in'], [2, 5])
    assert torch.allclose(out, tiny['resize_pos/out'], atol=1e-6)


def test_missing_index_generator_matches_reference_algorithm():
    """src/utils/generate_missing.py:23-38 with the global `random` module."""
    import random
    B, ratio, modal = 64, 0.3, ['image', 'depth', 'thermal']
    random.seed(2025)
    idxs = random.sample(range(B), int(B * ratio))
    exp = [0] * B
    for i in idxs:
        exp[i] = random.choice([R.MISSING_TYPE_INDEX[m] for m in modal])
    assert R.synth_missing_index(B, ratio, modal).tolist() == exp
    assert sum(1 for e in exp if e) == 19


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "config1_full.pt")), reason="full golden not generated")
def test_oracle_matches_reference_config1_full():
    """BASELINE.json config 1 (image ViT-L/14 + text, B = 8, one image-missing sample, fp32 CPU)."""
    g = torch.load(os.path.join(GOLD, "config1_full.pt"), weights_only=False)
    meta = g['meta']
    modal_types = ['language', 'image']
    cfgs, tcfg, sd = build_problem(dict(meta, modals=['image']), modal_types, 'sum')
    data = R.synth_inputs(modal_types, meta['B'], cfgs, tcfg, seed=0)
    with torch.no_grad():
        logits, emb = R.finetune_forward(sd, 'sum', modal_types, data, g['missing_index'], cfgs, tcfg,
                                         {'image': 2.6592})
    assert rel(logits, g['logits/sum']) < 1e-4
    for m in modal_types:
        assert rel(emb[m], g[f'emb/{m}']) < 1e-4
