"""BASELINE.json configs[1] -- the bench workload -- at FULL size on the CUDA path, through size-independent properties
(no CPU oracle finishes 3 x ViT-L/14 x 24 layers x B = 64 in seconds).  Kept in its own file, last in collection order:
it was written after the round's last GPU call, so its first run is the driver's, and under `pytest -x` a surprise here
must not hide the other GPU tests."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import restatement as R  # noqa: E402  (the checker, never the thing measured)
from test_parity_gpu import DEV, TOL, TOL_LOGIT, make, rel, to_dev  # noqa: E402


def test_config2_full_size_properties():
    """BASELINE.json configs[1] -- the bench workload -- at its FULL size (three ViT-L/14 towers, 24 layers, B = 64,
    30 % of the samples missing one modality, codes drawn as generate_missing.py does): no CPU oracle finishes that in
    seconds, so parity goes through size-independent properties: skipping the missing samples changes nothing,
    their embedding rows come back as exact zeros, the batch order does not matter, and the compaction indices are
    bit-exact against torch.nonzero."""
    from missm_b200 import config as C, ops
    modal = ['image', 'depth', 'thermal']
    v = {k: val for k, val in C.VIT_L14.items() if k != 'lora_r'}
    meta = dict(vision=v, text=dict(C.CLIP_TEXT), projection_dim=768, fusion_dim=256)
    model, cfgs, tcfg, _ = make(meta, modal, 'sum')
    model.eval()
    B = 64
    data = to_dev(R.synth_inputs(modal, B, cfgs, tcfg, seed=31))
    mi = R.synth_missing_index(B, 0.3, modal).to(DEV)
    assert int((mi != 0).sum()) == 19
    codes = [R.MISSING_TYPE_INDEX[m] for m in modal]
    idx, slot, counts = ops.compact_mask(mi, codes)
    for t, c in enumerate(codes):                                 # bit-exact indices at the bench batch
        present = torch.nonzero(mi != c).reshape(-1).to(torch.int32)
        assert int(counts[t]) == present.numel() and torch.equal(idx[t, :present.numel()], present)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(3)).to(DEV)
    with torch.no_grad():
        a = model(data, mi)
        emb_c = model.encoder(data, missing_index=mi)
        model.encoder.compaction = False
        b = model(data, mi)
        model.encoder.compaction = True
        emb = model.encoder(data)
        emb_p = model.encoder({m: {'pixel_values': x['pixel_values'][perm].contiguous()} for m, x in data.items()})
    # skipping missing samples changes nothing beyond bf16 noise (another row count picks other GEMM tile shapes)
    assert torch.isfinite(a).all() and rel(a, b) < TOL_LOGIT
    exact = True
    for m, c in zip(modal, codes):
        gone = mi == c
        assert gone.any() and emb_c[m][gone].abs().max().item() == 0.0          # zero rows, not garbage
        assert rel(emb_c[m][~gone], emb[m][~gone]) < TOL, m
        assert rel(emb_p[m], emb[m][perm]) < 1e-3, m                            # batch order does not matter
        exact = exact and torch.equal(emb_p[m], emb[m][perm])
    print('config 2 full size: permutation equivariance bit-exact =', exact, '; compaction on/off logits', rel(a, b))
