"""Parity of the CUDA product path (drop-in `languagebind` + `src.model.baseline`) against
(a) golden vectors produced by the unmodified reference (tests/golden/) and (b) the CPU oracle
(oracle/restatement.py) on the same seeded inputs.  Tolerance in bf16 mode (BASELINE.json
north_star): <= 1e-2 relative error on embeddings / logits / loss; compaction indices bit-exact.
Run with `-m gpu` on a B200; /root/reference is NOT needed."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import restatement as R  # noqa: E402  (the checker, never the thing measured)

GOLD = os.path.join(HERE, "golden")
DEV = "cuda"
TOL = 1e-2          # bf16 mode, relative (norm-wise): embeddings and loss (north_star)
TOL_LOGIT = 3e-2    # logits of the tiny random-weight heads are differences of O(1) features
TOL_GRAD = 3e-2     # gradients through 2-24 bf16 layers


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def to_dev(data):
    return {k: {kk: vv.to(DEV) for kk, vv in v.items()} for k, v in data.items()}


def make(meta, modal_types, fusion, dropout=0.0):
    from missm_b200 import shapes
    per = meta.get('per', {})
    cfgs = {}
    for m in modal_types:
        if m == 'language':
            continue
        d = {k: v for k, v in meta['vision'].items() if k != 'lora_r'}
        d.update(per.get(m, {}))
        d['temporal_mlp'] = (m != 'video')
        cfgs[m] = R.vision_config(**d)
    tcfg = R.text_config(**meta['text'])
    pd, fd = meta.get('projection_dim', 768), meta.get('fusion_dim', 256)
    model = shapes.build_finetune(cfgs, tcfg, modal_types, fusion, 3, pd, fd, dropout_prob=dropout)
    sd = R.synth_state_dict([(k, tuple(v.shape)) for k, v in model.state_dict().items()])
    shapes.load_named(model, sd)
    return model.to(DEV), cfgs, tcfg, sd


@pytest.fixture(scope="module")
def tiny():
    return torch.load(os.path.join(GOLD, "tiny_bank.pt"), weights_only=False)


FUSIONS = ['sum', 'concat', 'regression', 'retrieval', 'intra_attention', 'inter_attention',
           'dedicated_dnn', 'Distill_tea', 'self_distill']


@pytest.mark.parametrize("fusion", FUSIONS)
def test_forward_matches_reference_golden(tiny, fusion):
    """All five towers + text (tiny widths, head_dim 64) + each skip-safe fusion head, with one
    sample missing each of language/video/audio/image/thermal: logits vs the reference's."""
    meta = tiny['meta']
    modal_types = ['language'] + meta['modals']
    model, cfgs, tcfg, _ = make(meta, modal_types, fusion)
    model.eval()
    data = to_dev(R.synth_inputs(modal_types, meta['B'], cfgs, tcfg, seed=0))
    mi = tiny['missing_index'].to(DEV)
    with torch.no_grad():
        res = model(data, mi)
        emb = model.encoder(data)                      # full batch, no compaction (test.py:107 path)
        emb_c = model.encoder(data, missing_index=mi)  # compacted
    logits = res[-1] if isinstance(res, tuple) else res
    errs = {m: rel(emb[m], tiny[f'emb/{m}']) for m in modal_types}
    print(fusion, 'logits', rel(logits, tiny[f'logits/{fusion}']), 'emb', errs)
    assert rel(logits, tiny[f'logits/{fusion}']) < TOL_LOGIT
    for m in modal_types:
        assert errs[m] < TOL, (m, errs)
        miss = (mi == R.MISSING_TYPE_INDEX[m])
        assert torch.equal(emb_c[m][miss], torch.zeros_like(emb_c[m][miss]))          # zero-filled rows
        assert rel(emb_c[m][~miss], tiny[f'emb/{m}'][~miss.cpu()]) < TOL, m


def test_backward_matches_reference_golden(tiny):
    meta = tiny['meta']
    modal_types = ['language'] + meta['modals']
    model, cfgs, tcfg, _ = make(meta, modal_types, 'sum')
    model.train()
    data = to_dev(R.synth_inputs(modal_types, meta['B'], cfgs, tcfg, seed=0))
    logits = model(data, tiny['missing_index'].to(DEV))
    loss = torch.nn.functional.cross_entropy(logits, tiny['labels'].to(DEV))
    assert abs(loss.item() - tiny['loss/sum'].item()) < TOL * abs(tiny['loss/sum'].item())
    loss.backward()
    params = dict(model.named_parameters())
    for n, p in params.items():
        assert p.grad is not None, f"{n} got no gradient (DDP needs one for every parameter)"
        assert torch.isfinite(p.grad).all(), n
    worst = {}
    for k, v in tiny.items():
        if k.startswith('grad/'):
            g = params[k[5:]].grad
            err = (g.float().cpu() - v).norm().item() / (v.norm().item() + 1e-8)
            worst[k] = err
            if v.norm() > 1e-6:
                assert err < TOL_GRAD, (k, err)
    gn = tiny['grad_norms']
    bad = []
    for n, p in params.items():
        ref = gn.get(n)
        if ref is not None and ref > 1e-6:
            e = abs(p.grad.norm().item() - ref) / ref
            if e > 5e-2:
                bad.append((n, e))
    assert not bad, bad[:10]


def test_compaction_matches_full_batch(tiny):
    """Skipping missing samples must not change the logits and must give
    every parameter a gradient; a tower with NO present sample still returns zero grads."""
    meta = tiny['meta']
    modal_types = ['image', 'depth', 'thermal']
    model, cfgs, tcfg, _ = make(meta, modal_types, 'sum')
    model.eval()
    B = 8
    data = to_dev(R.synth_inputs(modal_types, B, cfgs, tcfg, seed=3))
    mi = torch.tensor([0, 4, 5, 6, 4, 0, 6, 6], device=DEV)
    with torch.no_grad():
        a = model(data, mi)
        model.encoder.compaction = False
        b = model(data, mi)
        model.encoder.compaction = True
    assert rel(a, b) < 5e-3
    model.train()
    mi_all = torch.full((B,), 5, device=DEV)            # depth missing everywhere -> empty tower
    out = model(data, mi_all)
    out.sum().backward()
    for n, p in model.named_parameters():
        # the text tower is registered but unused here, exactly as in the reference
        # (languagebind/__init__.py:69-70): it gets no gradient there either
        if 'language' not in n:
            assert p.grad is not None, n
    dg = [p.grad.abs().max().item() for n, p in model.named_parameters() if 'modality_encoder.depth' in n]
    assert max(dg) == 0.0


def test_host_inputs_match_device_inputs(tiny):
    """finetune_model.forward(pinned HOST tensors, host missing_index) -- the end-to-end call bench.py times --
    uploads every tower's input on that tower's stream; results must be identical to device-resident inputs."""
    meta = tiny['meta']
    modal_types = ['image', 'depth', 'thermal']
    model, cfgs, tcfg, _ = make(meta, modal_types, 'sum')
    model.train()
    B = 8
    host = R.synth_inputs(modal_types, B, cfgs, tcfg, seed=5)
    host = {m: {k: t.pin_memory() for k, t in v.items()} for m, v in host.items()}
    mi_host = torch.tensor([0, 4, 5, 6, 4, 0, 6, 0]).pin_memory()
    out_h = model(host, mi_host)
    out_h.sum().backward()
    g_h = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    out_d = model(to_dev(host), mi_host.to(DEV))
    out_d.sum().backward()
    assert torch.equal(out_h, out_d)
    for n, p in model.named_parameters():
        if p.grad is not None:
            assert rel(g_h[n], p.grad) < 1e-5, n       # split-K wgrad atomics reorder fp32 sums, nothing else differs


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "config1_full.pt")), reason="full golden not generated")
def test_config1_full_size_matches_reference_golden():
    """BASELINE.json config 1: ViT-L/14 224 image tower + text tower + sum head, B = 8, one
    image-missing sample -- embeddings / logits vs the reference run on CPU fp32."""
    g = torch.load(os.path.join(GOLD, "config1_full.pt"), weights_only=False)
    meta = dict(g['meta'], modals=['image'])
    modal_types = ['language', 'image']
    model, cfgs, tcfg, _ = make(meta, modal_types, 'sum')
    model.eval()
    data = to_dev(R.synth_inputs(modal_types, meta['B'], cfgs, tcfg, seed=0))
    mi = g['missing_index'].to(DEV)
    with torch.no_grad():
        logits = model(data, mi)
        emb = model.encoder(data)
    for m in modal_types:
        print('config1', m, rel(emb[m], g[f'emb/{m}']))
        assert rel(emb[m], g[f'emb/{m}']) < TOL, (m, rel(emb[m], g[f'emb/{m}']))
    print('config1 logits', rel(logits, g['logits/sum']))
    assert rel(logits, g['logits/sum']) < TOL_LOGIT
    # the same full-size forward in the fp32 VERIFICATION mode (3-way split tcgen05 GEMMs, fp32 attention):
    # north_star's fp32 tolerance, <= 1e-5 relative on the embeddings, through all 24 / 12 layers
    from missm_b200 import autograd as ag
    old = ag.set_precision("fp32")
    try:
        with torch.no_grad():
            emb32 = model.encoder(data)
            logits32 = model(data, mi)
    finally:
        ag.set_precision(old)
    for m in modal_types:
        print('config1 fp32 mode', m, rel(emb32[m], g[f'emb/{m}']))
    print('config1 fp32 mode logits', rel(logits32, g['logits/sum']))
    for m in modal_types:
        assert rel(emb32[m], g[f'emb/{m}']) < 1e-5, (m, rel(emb32[m], g[f'emb/{m}']))
    assert rel(logits32, g['logits/sum']) < 1e-4
    # B = 4 training step: loss and gradient norms
    model.train()
    data4 = {k: {kk: vv[:4] for kk, vv in v.items()} for k, v in data.items()}
    lg = model(data4, mi[:4])
    loss = torch.nn.functional.cross_entropy(lg, g['labels4'].to(DEV))
    assert abs(loss.item() - g['loss4'].item()) < TOL * abs(g['loss4'].item())
    loss.backward()
    params = dict(model.named_parameters())
    for k in ('fusion.modal_proj.image.weight', 'encoder.modality_encoder.image.embeddings.class_embedding'):
        assert rel(params[k].grad, g[f'grad/{k}']) < TOL_GRAD, k
    bad = []
    for n, ref in g['grad_norms4'].items():
        if ref > 1e-6:
            e = abs(params[n].grad.norm().item() - ref) / ref
            if e > 5e-2:
                bad.append((n, e))
    assert len(bad) <= 0.02 * len(params), bad[:10]


def test_oracle_parity_random_shapes():
    """Product vs the CPU oracle on fresh seeds / ragged sizes (audio non-square grid, B = 3)."""
    meta = dict(vision=dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2,
                            patch_size=14, image_size=84),
                text=dict(hidden_size=128, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                          vocab_size=500, max_position_embeddings=77),
                per={'audio': dict(num_mel_bins=42, target_length=98)}, projection_dim=128, fusion_dim=128)
    modal_types = ['language', 'audio', 'image']
    model, cfgs, tcfg, sd = make(meta, modal_types, 'sum')
    model.eval()
    data = R.synth_inputs(modal_types, 3, cfgs, tcfg, seed=11)
    mi = torch.tensor([3, 0, 1])
    with torch.no_grad():
        ref, ref_emb = R.finetune_forward(sd, 'sum', modal_types, data, mi, cfgs, tcfg, {m: 2.6592 for m in cfgs})
        out = model(to_dev(data), mi.to(DEV))
        emb = model.encoder(to_dev(data))
    for m in modal_types:
        assert rel(emb[m], ref_emb[m]) < TOL, (m, rel(emb[m], ref_emb[m]))
    assert rel(out, ref) < TOL_LOGIT
