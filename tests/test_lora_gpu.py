"""GPU parity of the LoRA row: the CUDA path with peft-style adapters on the
attention projections against (a) the golden of the reference's own convert_to_lora (tests/golden/tiny_lora.pt,
made by oracle/make_golden.py) and (b) the CPU oracle on one full-width encoder layer whose row count sends the
GEMMs to the CTA-pair kernel (K' = D + 8 contraction tails, N = 8 skinny outputs, pitched operand views).
Tolerances as the other parity tests: bf16 mode <= 1e-2 on embeddings / loss, fp32 mode <= 1e-5."""
import os
import sys
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
sys.path.insert(0, HERE)
import restatement as R  # noqa: E402  (the checker, never the thing measured)
from test_lora_host import lora_cfgs, make_model  # noqa: E402

GOLD = os.path.join(HERE, "golden")
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "tiny_lora.pt"), weights_only=False)


@pytest.mark.parametrize("precision,tol_emb,tol_logit,tol_grad", [("bf16", 1e-2, 3e-2, 5e-2), ("fp32", 1e-5, 1e-4, 2e-4)])
def test_lora_towers_match_reference_golden(gold, precision, tol_emb, tol_logit, tol_grad):
    from missm_b200 import autograd as ag
    meta = gold['meta']
    model, _, cfgs, tcfg, modal_types = make_model(meta)
    model = model.to(DEV).train()
    data = R.synth_inputs(modal_types, meta['B'], cfgs, tcfg, seed=meta['seed'])
    data = {k: {kk: vv.to(DEV) for kk, vv in v.items()} for k, v in data.items()}
    mi = gold['missing_index'].to(DEV)
    old = ag.set_precision(precision)
    try:
        with torch.no_grad():
            emb = model.encoder(data)
        logits = model(data, mi)
        loss = torch.nn.functional.cross_entropy(logits, gold['labels'].to(DEV))
        loss.backward()
        torch.cuda.synchronize()
    finally:
        ag.set_precision(old)
    errs = {m: rel(emb[m], gold[f'emb/{m}']) for m in modal_types}
    print('lora', precision, 'emb', errs, 'logits', rel(logits, gold['logits/sum']))
    for m in modal_types:
        assert errs[m] < tol_emb, errs
    assert rel(logits, gold['logits/sum']) < tol_logit
    assert abs(loss.item() - gold['loss/sum'].item()) < max(tol_emb, 1e-5) * abs(gold['loss/sum'].item())
    params = dict(model.named_parameters())
    worst = ('', 0.0)
    for k, v in gold.items():
        if k.startswith('grad/') and v.norm() > 1e-7:
            e = rel(params[k[5:]].grad, v)
            worst = max(worst, (k, e), key=lambda t: t[1])
            assert e < tol_grad, (k, e)
    print('lora', precision, 'worst gradient', worst)
    for n, ref in gold['grad_norms'].items():
        if ref is None:
            assert params[n].grad is None, n           # frozen by peft: no wgrad was computed


@pytest.mark.parametrize("D,H,N,S,r", [(1024, 16, 257, 6, 2), (128, 2, 77, 3, 4)])
def test_lora_layer_full_width_vs_oracle(D, H, N, S, r):
    """One encoder layer with adapters (rank r, padded to 8 columns per group) at M = S * N rows: M >= 1024 goes
    through the cta_group::2 GEMM, M = 231 through the one-CTA kernel; forward and every adapter gradient vs the
    fp32 oracle (bf16 operands: <= 1e-2 forward, <= 3e-2 gradients)."""
    from missm_b200 import autograd as ag, config as C, ops, towers as T
    torch.manual_seed(D + N)
    cfg = C.CLIPVisionConfig(hidden_size=D, intermediate_size=2 * D, num_hidden_layers=1, num_attention_heads=H,
                             patch_size=14, image_size=224, lora_r=r, lora_alpha=16)
    layer = T.CLIPEncoderLayer(cfg)
    sd = R.synth_state_dict([(k, tuple(v.shape)) for k, v in layer.state_dict().items()])
    layer.load_state_dict(sd)
    for n, p in layer.named_parameters():
        p.requires_grad = '.lora_' in n
    layer = layer.to(DEV)
    x = torch.randn(S * N, D)
    g = torch.randn(S * N, D)
    xd = x.clone().to(DEV).requires_grad_(True)
    meta = ag.AttnMeta(H, cfg.layer_norm_eps, ops.SeqLayout.spatial(S, N))
    y = layer.run(xd, meta, None)
    y.backward(g.to(DEV))
    torch.cuda.synchronize()
    # oracle: the same layer, fp32, CPU
    ocfg = types.SimpleNamespace(num_attention_heads=H, layer_norm_eps=cfg.layer_norm_eps, hidden_act='quick_gelu',
                                 add_time_attn=False, lora_r=r, lora_alpha=16)
    sdr = {k: v.clone().requires_grad_('.lora_' in k) for k, v in sd.items()}
    xr = x.view(S, N, D).clone().requires_grad_(True)
    yr = R.encoder_layer(sdr, '', xr, ocfg)
    yr.backward(g.view(S, N, D))
    e_y, e_x = rel(y, yr.reshape(S * N, D)), rel(xd.grad, xr.grad.reshape(S * N, D))
    print('lora layer', (D, H, N, S, r), 'y', e_y, 'dx', e_x)
    assert e_y < 1e-2 and e_x < 3e-2
    params = dict(layer.named_parameters())
    for k, v in sdr.items():
        if '.lora_' in k:
            e = rel(params[k].grad, v.grad)
            print('   ', k, e)
            assert e < 3e-2, (k, e)
        else:
            assert params[k].grad is None, k
