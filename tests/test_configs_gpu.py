"""BASELINE.json configs[2] (audio + 8-frame video towers with missing-modality masking) and configs[4]
(test.py-style eval sweep over missing rates 0-90 %) on the CUDA path, at the configs' FULL token /
width geometry (ViT-L/14 width 1024, 16 heads, N = 593 audio tokens, T = 8 frames x 257 tokens; depth
cut to 2 layers so the CPU oracle finishes in seconds), checked (a) against the CPU oracle on the same
seeded inputs and (b) through size-independent properties: compaction on == compaction off, batch
permutation equivariance (bit-exact: every kernel computes a row from that row alone), mask-compaction
indices bit-exact vs torch.nonzero.  Run with `-m gpu` on a B200."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import restatement as R  # noqa: E402  (the checker, never the thing measured)
from test_parity_gpu import DEV, TOL, TOL_LOGIT, make, rel, to_dev  # noqa: E402

FULL = dict(vision=dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=2, num_attention_heads=16,
                        patch_size=14, image_size=224),
            text=dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=1, num_attention_heads=12,
                      vocab_size=49408, max_position_embeddings=77),
            per={'audio': dict(num_mel_bins=112, target_length=1036), 'video': dict(add_time_attn=True, num_frames=8)},
            projection_dim=768, fusion_dim=256)


def test_config3_audio_video_full_geometry_vs_oracle():
    modal = ['audio', 'video']
    model, cfgs, tcfg, sd = make(FULL, modal, 'sum')
    model.eval()
    B = 2
    data = R.synth_inputs(modal, B, cfgs, tcfg, seed=21)
    assert data['audio']['pixel_values'].shape == (B, 3, 112, 1036)          # 8 x 74 patches + CLS = 593 tokens
    assert data['video']['pixel_values'].shape == (B, 3, 8, 224, 224)
    mi = torch.tensor([0, 3])                                                 # sample 1 misses audio
    with torch.no_grad():
        ref, ref_emb = R.finetune_forward(sd, 'sum', modal, data, mi, cfgs, tcfg, {m: 2.6592 for m in cfgs})
        out = model(to_dev(data), mi.to(DEV))
        emb = model.encoder(to_dev(data))
    for m in modal:
        assert rel(emb[m], ref_emb[m]) < TOL, (m, rel(emb[m], ref_emb[m]))
    assert rel(out, ref) < TOL_LOGIT


def test_config3_properties_and_backward():
    modal = ['audio', 'video']
    model, cfgs, tcfg, _ = make(FULL, modal, 'sum')
    B = 6
    data = to_dev(R.synth_inputs(modal, B, cfgs, tcfg, seed=22))
    mi = torch.tensor([0, 3, 2, 0, 3, 0], device=DEV)          # codes: audio = 3, video = 2
    model.eval()
    with torch.no_grad():
        a = model(data, mi)
        model.encoder.compaction = False
        b = model(data, mi)
        model.encoder.compaction = True
        emb = model.encoder(data)
        perm = torch.tensor([4, 2, 5, 0, 1, 3], device=DEV)
        emb_p = model.encoder({m: {'pixel_values': v['pixel_values'][perm].contiguous()} for m, v in data.items()})
    assert rel(a, b) < 5e-3                                      # skipping missing samples changes nothing
    for m in modal:                                              # rows are computed from that row alone
        assert torch.equal(emb_p[m], emb[m][perm]), m
    model.train()
    labels = torch.arange(B, device=DEV) % 3
    loss = torch.nn.functional.cross_entropy(model(data, mi), labels)
    loss.backward()
    assert torch.isfinite(loss)
    for n, p in model.named_parameters():
        if 'language' in n:
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    g = dict(model.named_parameters())
    assert g['encoder.modality_encoder.video.encoder.layers.0.temporal_attn.q_proj.weight'].grad.abs().max() > 0
    assert g['encoder.modality_encoder.audio.embeddings.position_embedding.weight'].grad.shape[0] == 593


@pytest.mark.parametrize("ratio", [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9])
def test_config5_eval_sweep(ratio):
    """test.py-style no_grad evaluation at every missing rate of the sweep (data_loader.py:348,354)."""
    from missm_b200 import ops
    from missm_b200.bank import MISSING_TYPE_INDEX
    meta = dict(FULL, vision=dict(FULL['vision'], hidden_size=256, intermediate_size=512, num_attention_heads=4))
    modal = ['image', 'depth', 'thermal']
    model, cfgs, tcfg, sd = make(meta, modal, 'sum')
    model.eval()
    B = 20
    host = R.synth_inputs(modal, B, cfgs, tcfg, seed=5)
    mi = R.synth_missing_index(B, ratio, modal, seed=2025)
    assert int((mi != 0).sum()) == int(B * ratio)
    codes = [MISSING_TYPE_INDEX[m] for m in modal]
    idx, slot, counts = ops.compact_mask(mi.to(DEV), codes)
    for t, c in enumerate(codes):                                 # bit-exact vs torch.nonzero
        want = torch.nonzero(mi != c).flatten().to(torch.int32)
        assert int(counts[t]) == want.numel()
        assert torch.equal(idx[t, :want.numel()].cpu(), want)
    with torch.no_grad():
        out = model(to_dev(host), mi.to(DEV))
        model.encoder.compaction = False
        full = model(to_dev(host), mi.to(DEV))
        ref, _ = R.finetune_forward(sd, 'sum', modal, host, mi, cfgs, tcfg, {m: 2.6592 for m in cfgs})
    assert rel(out, full) < 5e-3
    assert rel(out, ref) < TOL_LOGIT
