"""CPU tests of the LoRA row (reference: convert_to_lora, languagebind/image/
modeling_image.py:775-793, config default lora_r = 2, configuration_image.py:200): parameter names / shapes / the
trainable set against the reference (run through its own convert_to_lora + the peft restatement of
oracle/ref_shim.py -> tests/golden/tiny_lora.pt), the oracle against that golden, the checkpoint-key layouts, and
the HOST algebra of autograd.LoraAttnBlockFn (adapters carried through the contraction dimension of the GEMMs)
over torch stand-ins of the C ABI (tests/ops_emulation.py -- test infrastructure, the product has no CPU path).
The CUDA kernels behind the same calls are checked on the B200 by tests/test_lora_gpu.py."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
sys.path.insert(0, HERE)
import restatement as R  # noqa: E402
import ops_emulation as E  # noqa: E402

GOLD = os.path.join(HERE, "golden")


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "tiny_lora.pt"), weights_only=False)


def lora_cfgs(meta):
    cfgs = {}
    for m in meta['modals']:
        d = dict(meta['vision'])
        d.update(meta['per'].get(m, {}))
        d['temporal_mlp'] = (m != 'video')
        cfgs[m] = R.vision_config(**d)
    return cfgs, R.text_config(**meta['text'])


def make_model(meta):
    from missm_b200 import shapes
    cfgs, tcfg = lora_cfgs(meta)
    modal_types = ['language'] + meta['modals']
    model = shapes.build_finetune(cfgs, tcfg, modal_types, 'sum', meta['n_classes'], meta['projection_dim'],
                                  meta['fusion_dim'], dropout_prob=0.0)
    sd = R.synth_state_dict([(k, tuple(v.shape)) for k, v in model.state_dict().items()])
    shapes.load_named(model, sd)
    return model, sd, cfgs, tcfg, modal_types


def test_names_shapes_and_trainable_set_match_reference(gold):
    model, _, _, _, _ = make_model(gold['meta'])
    own = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert own == gold['names']
    assert [n for n, p in model.named_parameters() if p.requires_grad] == gold['trainable']
    # spatial LoRA on the image tower, temporal-attention LoRA only on the video tower (modeling_image.py:778-783)
    assert 'encoder.modality_encoder.image.encoder.base_model.model.layers.0.self_attn.q_proj.lora_A.default.weight' in own
    assert 'encoder.modality_encoder.video.encoder.base_model.model.layers.0.temporal_attn.out_proj.lora_B.default.weight' in own
    assert not any('video' in k and 'self_attn' in k and 'lora_' in k for k in own)


def test_oracle_matches_reference_lora_golden(gold):
    meta = gold['meta']
    cfgs, tcfg = lora_cfgs(meta)
    modal_types = ['language'] + meta['modals']
    sd = R.synth_state_dict(list(gold['names'].items()))
    data = R.synth_inputs(modal_types, meta['B'], cfgs, tcfg, seed=meta['seed'])
    scales = {m: float(sd[f'encoder.modality_scale.{m}']) if f'encoder.modality_scale.{m}' in sd else 2.6592
              for m in meta['modals']}
    with torch.no_grad():
        logits, emb = R.finetune_forward(sd, 'sum', modal_types, data, gold['missing_index'], cfgs, tcfg, scales)
    for m in modal_types:
        assert rel(emb[m], gold[f'emb/{m}']) < 2e-5, m
    assert rel(logits, gold['logits/sum']) < 2e-5


def _step(model, data, gold):
    logits = model(data, gold['missing_index'])
    loss = torch.nn.functional.cross_entropy(logits, gold['labels'])
    loss.backward()
    return logits, loss


def _check(model, logits, loss, gold, tol_logits, tol_loss, tol_grad, tol_norm):
    assert rel(logits, gold['logits/sum']) < tol_logits, rel(logits, gold['logits/sum'])
    assert abs(loss.item() - gold['loss/sum'].item()) < tol_loss * abs(gold['loss/sum'].item())
    params = dict(model.named_parameters())
    n_lora = 0
    for k, v in gold.items():
        if k.startswith('grad/'):
            g = params[k[5:]].grad
            assert g is not None, k
            if v.norm() > 1e-7:
                assert rel(g, v) < tol_grad, (k, rel(g, v))
            n_lora += '.lora_' in k
    assert n_lora == 2 * 4 * 2 * 2            # (A, B) x (q, k, v, out) x 2 layers x 2 towers
    for n, ref in gold['grad_norms'].items():
        if ref is None:                        # frozen by peft: the product must not produce a gradient either
            assert params[n].grad is None, n
        elif ref > 1e-6:
            assert abs(params[n].grad.norm().item() - ref) < tol_norm * ref, n


def test_lora_block_host_algebra_exact(gold):
    """Product blocks (autograd.LoraAttnBlockFn, frozen-encoder dgrad-only paths) with every "bf16" buffer widened
    to fp32: pitched operand views, the [X | T] / [W | sB] / [W ; A] packing and every adapter-gradient formula must
    reproduce the reference's fwd + bwd at fp32 accuracy."""
    model, _, cfgs, tcfg, modal_types = make_model(gold['meta'])
    model.train()
    data = R.synth_inputs(modal_types, gold['meta']['B'], cfgs, tcfg, seed=gold['meta']['seed'])
    with E.emulated_fp32_mode(precision="bf16", wide_bf16=True):
        with torch.no_grad():
            emb = model.encoder(data)
        logits, loss = _step(model, data, gold)
    for m in modal_types:
        assert rel(emb[m], gold[f'emb/{m}']) < 1e-5, (m, rel(emb[m], gold[f'emb/{m}']))
    _check(model, logits, loss, gold, 1e-4, 1e-5, 2e-4, 2e-4)


def test_lora_block_bf16_rounding_within_tolerance(gold):
    """The same with real bf16 operand rounding (what the tcgen05 GEMMs see): north_star's <= 1e-2 on embeddings."""
    model, _, cfgs, tcfg, modal_types = make_model(gold['meta'])
    model.train()
    data = R.synth_inputs(modal_types, gold['meta']['B'], cfgs, tcfg, seed=gold['meta']['seed'])
    with E.emulated_fp32_mode(precision="bf16"):
        with torch.no_grad():
            emb = model.encoder(data)
        logits, loss = _step(model, data, gold)
    for m in modal_types:
        assert rel(emb[m], gold[f'emb/{m}']) < 1e-2, (m, rel(emb[m], gold[f'emb/{m}']))
    _check(model, logits, loss, gold, 3e-2, 1e-2, 5e-2, 5e-2)


def test_lora_fp32_verification_mode(gold):
    """MISSM_PRECISION=fp32: the adapter folded into W + sBA, gradients carried back to A and B by autograd."""
    model, _, cfgs, tcfg, modal_types = make_model(gold['meta'])
    model.train()
    data = R.synth_inputs(modal_types, gold['meta']['B'], cfgs, tcfg, seed=gold['meta']['seed'])
    with E.emulated_fp32_mode():
        logits, loss = _step(model, data, gold)
    _check(model, logits, loss, gold, 1e-4, 1e-5, 2e-4, 2e-4)


def test_adapter_refresh_after_update(gold):
    """The packed operands follow an in-place parameter update (optimizer step) of the adapters only."""
    model, _, cfgs, tcfg, modal_types = make_model(gold['meta'])
    model.eval()
    data = R.synth_inputs(modal_types, gold['meta']['B'], cfgs, tcfg, seed=gold['meta']['seed'])
    with E.emulated_fp32_mode(precision="bf16", wide_bf16=True):
        with torch.no_grad():
            e0 = model.encoder(data)['image'].clone()
            for n, p in model.named_parameters():
                if 'image' in n and 'lora_B' in n:
                    p.zero_()
            e1 = model.encoder(data)['image'].clone()
    assert rel(e0, gold['emb/image']) < 1e-5
    sd = {k: (torch.zeros_like(v) if 'lora_B' in k and '.image.' in k else v)
          for k, v in R.synth_state_dict(list(gold['names'].items())).items()}
    with torch.no_grad():
        ref = R.vision_tower(sd, 'encoder.modality_encoder.image.', data['image']['pixel_values'], cfgs['image'])
        ref = torch.nn.functional.linear(ref, sd['encoder.modality_proj.image.weight'])
        ref = ref / ref.norm(dim=-1, keepdim=True) * float(torch.tensor(2.6592).exp())
    assert rel(e1, ref) < 1e-5 and rel(e0, e1) > 1e-2


def test_checkpoint_key_layouts():
    from missm_b200 import config as C, towers as T
    vc = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2, patch_size=14,
              image_size=28)
    tc = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2, vocab_size=100)
    torch.manual_seed(0)
    plain = T.LanguageBindImage(C.LanguageBindImageConfig(text_config=tc, vision_config=dict(vc, lora_r=0),
                                                         projection_dim=32))
    lora = T.LanguageBindImage(C.LanguageBindImageConfig(text_config=tc, vision_config=dict(vc, lora_r=2),
                                                        projection_dim=32))
    # (1) a plain (merged / lora_r = 0) checkpoint into a LoRA-configured model: base weights land under
    #     encoder.base_model.model, adapters stay at the no-op initialisation
    lora.load_reference_state_dict(plain.state_dict())
    w = 'vision_model.encoder.base_model.model.layers.0.self_attn.q_proj.'
    assert torch.equal(lora.state_dict()[w + 'weight'], plain.state_dict()['vision_model.encoder.layers.0.self_attn.q_proj.weight'])
    assert lora.state_dict()[w + 'lora_B.default.weight'].abs().max() == 0
    # (2) the peft >= 0.6 spelling of the wrapped layout (`base_layer.`)
    sd = {}
    for k, v in lora.state_dict().items():
        if any(k.endswith(p + s) for p in ('q_proj.', 'k_proj.', 'v_proj.', 'out_proj.') for s in ('weight', 'bias')) \
                and 'base_model' in k:
            k = k.rsplit('.', 1)[0] + '.base_layer.' + k.rsplit('.', 1)[1]
        sd[k] = v.clone() + 1.0
    lora.load_reference_state_dict(sd)
    assert torch.equal(lora.state_dict()[w + 'weight'], sd[w + 'base_layer.weight'])
    # (3) adapters into a model built without them: refuse loudly
    with pytest.raises(RuntimeError):
        plain.load_reference_state_dict(lora.state_dict())
    # frozen set = everything inside the wrapped encoder except the adapters
    for n, p in lora.named_parameters():
        inside = n.startswith('vision_model.encoder.')
        assert p.requires_grad == (not inside or '.lora_' in n), n


def test_finetune_checkpoint_round_trips_both_peft_layouts(gold):
    """`final_model/<ds>_<fusion>.pth` files go through a plain, STRICT finetune_model.load_state_dict in the scripts
    (test.py:92, train_ddp.py:193,317).  The reference does not pin peft: a checkpoint written under peft >= 0.6
    spells the frozen Linear `...q_proj.base_layer.weight`, one written under 0.4 / 0.5 `...q_proj.weight`.  Both
    must load strictly, and a checkpoint of the unwrapped encoder must load non-strictly (adapters stay no-ops)."""
    model, sd, _, _, _ = make_model(gold['meta'])
    own = {k: v.clone() for k, v in model.state_dict().items()}
    new_layout = {}
    for k, v in own.items():
        head, leaf = k.rsplit('.', 1)
        if 'base_model.model' in k and head.rsplit('.', 1)[-1] in ('q_proj', 'k_proj', 'v_proj', 'out_proj') and \
                leaf in ('weight', 'bias') and (head + '.lora_A.default.weight') in own:
            k = head + '.base_layer.' + leaf
        new_layout[k] = v.clone() + 0.5
    assert any('.base_layer.' in k for k in new_layout)
    res = model.load_state_dict(new_layout)                      # strict, as the scripts call it
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in model.state_dict().items():
        assert torch.equal(v, own[k] + 0.5) or not v.is_floating_point(), k
    res = model.load_state_dict({k: v - 0.5 for k, v in model.state_dict().items()})     # 0.4 / 0.5 layout: strict too
    assert not res.missing_keys and not res.unexpected_keys
    # unwrapped-encoder keys (`...encoder.layers.N...`): accepted, only the adapters are reported missing
    plain = {k.replace('.encoder.base_model.model.', '.encoder.'): v for k, v in model.state_dict().items() if '.lora_' not in k}
    res = model.load_state_dict(plain, strict=False)
    assert not res.unexpected_keys and res.missing_keys and all('.lora_' in k for k in res.missing_keys)
