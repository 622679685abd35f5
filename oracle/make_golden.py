"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt by executing the UNMODIFIED reference
(/root/reference, through oracle/ref_shim.py) in the build container.

    python oracle/make_golden.py [--full]

The reference has no golden vectors of its own; these files pin the oracle
(oracle/restatement.py) and, through it and directly, the CUDA path.  Weights are the
name-seeded synthetic tensors of restatement.synth_state_dict, inputs are
restatement.synth_inputs, so a test can rebuild the exact same problem without /root/reference.
"""
import argparse
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import restatement as R  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

TINY_V = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
              patch_size=14, image_size=56, lora_r=0)
TINY_T = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
              vocab_size=1000, max_position_embeddings=77)
TINY_PER = {'video': dict(add_time_attn=True, num_frames=4),
            'audio': dict(num_mel_bins=28, target_length=70)}
TINY_MODALS = ['video', 'audio', 'image', 'depth', 'thermal']
FUSIONS = ['sum', 'concat', 'regression', 'retrieval', 'intra_attention', 'inter_attention',
           'dedicated_dnn', 'Distill_tea', 'self_distill']

FULL_V = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
              patch_size=14, image_size=224, lora_r=0)
FULL_T = dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
              vocab_size=49408, max_position_embeddings=77)


def load_synth(model):
    named = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    sd = R.synth_state_dict(named)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all('position_ids' in m for m in missing), (missing, unexpected)
    return sd


def cfg_objects(vdict, tdict, per, modals):
    cfgs = {}
    for m in modals:
        d = {k: v for k, v in vdict.items() if k != 'lora_r'}
        d.update(per.get(m, {}))
        d['temporal_mlp'] = (m != 'video')
        cfgs[m] = R.vision_config(**d)
    return cfgs, R.text_config(**tdict)


def tiny():
    torch.manual_seed(0)
    bank = ref_shim.build_reference_bank(TINY_MODALS, TINY_V, TINY_T, projection_dim=64,
                                         per_modality_cfg=TINY_PER)
    cfgs, tcfg = cfg_objects(TINY_V, TINY_T, TINY_PER, TINY_MODALS)
    modal_types = ['language'] + TINY_MODALS
    B = 6
    data = R.synth_inputs(modal_types, B, cfgs, tcfg, seed=0)
    missing_index = torch.tensor([0, 4, 2, 1, 6, 3], dtype=torch.long)
    out = {'meta': dict(vision=TINY_V, text=TINY_T, per=TINY_PER, modals=TINY_MODALS, projection_dim=64,
                        B=B, fusion_dim=32, n_classes=3), 'missing_index': missing_index}
    extra = {'depth': 5, 'thermal': 6}
    names = {}
    for fusion in FUSIONS:
        model = ref_shim.build_reference_model(bank, fusion, modal_types, 3, feature_dims=64, fusion_dim=32,
                                               dropout_prob=0.0, extra_missing_codes=extra)
        names[fusion] = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        load_synth(model)
        model.eval()
        with torch.no_grad():
            emb = model.encoder({k: dict(v) for k, v in data.items()})
            res = model({k: dict(v) for k, v in data.items()}, missing_index)
        logits = res[-1] if isinstance(res, tuple) else res
        out[f'logits/{fusion}'] = logits.clone()
        if fusion == 'sum':
            for k, v in emb.items():
                out[f'emb/{k}'] = v.clone()
    # gradients for the default head (train mode, dropout_prob = 0 so it is deterministic)
    model = ref_shim.build_reference_model(bank, 'sum', modal_types, 3, feature_dims=64, fusion_dim=32,
                                           dropout_prob=0.0, extra_missing_codes=extra)
    load_synth(model)
    model.train()
    labels = torch.tensor([0, 1, 2, 0, 1, 2])
    logits = model({k: dict(v) for k, v in data.items()}, missing_index)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    out['labels'] = labels
    out['loss/sum'] = loss.detach().clone()
    keep = ['fusion.modal_proj.image.weight', 'fusion.norm.weight', 'fusion.head.head.3.bias',
            'encoder.modality_proj.image.weight',
            'encoder.modality_encoder.image.post_layernorm.weight',
            'encoder.modality_encoder.image.encoder.layers.1.mlp.fc2.weight',
            'encoder.modality_encoder.image.encoder.layers.0.self_attn.q_proj.weight',
            'encoder.modality_encoder.image.encoder.layers.0.self_attn.k_proj.bias',
            'encoder.modality_encoder.image.encoder.layers.0.layer_norm1.weight',
            'encoder.modality_encoder.image.embeddings.class_embedding',
            'encoder.modality_encoder.image.embeddings.position_embedding.weight',
            'encoder.modality_encoder.image.embeddings.patch_embedding.weight',
            'encoder.modality_encoder.audio.embeddings.patch_embedding.weight',
            'encoder.modality_encoder.video.encoder.layers.0.temporal_embedding',
            'encoder.modality_encoder.video.encoder.layers.1.temporal_attn.v_proj.weight',
            'encoder.modality_encoder.video.encoder.layers.0.temporal_layer_norm1.bias',
            'encoder.modality_encoder.language.embeddings.position_embedding.weight',
            'encoder.modality_encoder.language.encoder.layers.0.mlp.fc1.weight',
            'encoder.modality_encoder.language.final_layer_norm.weight']
    gn = {}
    for n, p in model.named_parameters():
        gn[n] = float(p.grad.norm()) if p.grad is not None else None
        if n in keep:
            out[f'grad/{n}'] = p.grad.detach().clone()
    out['grad_norms'] = gn
    # resize_pos golden: 4x4 grid table -> 2x5 (audio) through the reference's own method
    from transformers.models.clip import modeling_clip as mc
    import languagebind as lb
    acfg = lb.config_dict['audio'](text_config=dict(TINY_T), vision_config=dict(TINY_V, **TINY_PER['audio']),
                                   projection_dim=64)
    sq = lb.config_dict['image'](text_config=dict(TINY_T), vision_config=dict(TINY_V), projection_dim=64)
    emb_mod = mc.CLIPVisionEmbeddings(sq.vision_config)
    table = R.synth_param('resize_pos_test', (17, 128), 0.5)
    with torch.no_grad():
        emb_mod.position_embedding.weight.copy_(table)
    amodel = lb.model_dict['audio'](acfg)
    amodel.resize_pos(emb_mod, acfg.vision_config)
    out['resize_pos/in'] = table
    out['resize_pos/out'] = emb_mod.position_embedding.weight.detach().clone()
    torch.save({'meta': dict(vision=TINY_V, text=TINY_T, per=TINY_PER, modals=TINY_MODALS), 'fusions': names},
               os.path.join(GOLD, 'reference_param_shapes.pt'))
    torch.save(out, os.path.join(GOLD, 'tiny_bank.pt'))
    print('tiny golden written:', {k: (tuple(v.shape) if torch.is_tensor(v) else type(v).__name__)
                                  for k, v in out.items() if not k.startswith('grad/')})


LORA_V = dict(TINY_V, lora_r=2, lora_alpha=16)
LORA_PER = {'video': dict(add_time_attn=True, num_frames=4)}
LORA_MODALS = ['video', 'image']


def lora():
    """peft-wrapped encoders (the reference config DEFAULT is lora_r = 2,
    configuration_image.py:200).  The reference's own convert_to_lora (modeling_image.py:775-793) runs against the
    peft restatement of ref_shim.py: spatial LoRA on the image tower, temporal-attention LoRA on the video tower,
    frozen encoder weights, a trained (non-zero B) adapter; fwd + bwd of the `sum` head."""
    torch.manual_seed(0)
    bank = ref_shim.build_reference_bank(LORA_MODALS, LORA_V, TINY_T, projection_dim=64, per_modality_cfg=LORA_PER)
    cfgs = {}
    for m in LORA_MODALS:
        d = dict(LORA_V)
        d.update(LORA_PER.get(m, {}))
        d['temporal_mlp'] = (m != 'video')
        cfgs[m] = R.vision_config(**d)
    tcfg = R.text_config(**TINY_T)
    modal_types = ['language'] + LORA_MODALS
    B = 5
    data = R.synth_inputs(modal_types, B, cfgs, tcfg, seed=3)
    missing_index = torch.tensor([0, 4, 2, 0, 1], dtype=torch.long)
    model = ref_shim.build_reference_model(bank, 'sum', modal_types, 3, feature_dims=64, fusion_dim=32,
                                           dropout_prob=0.0, extra_missing_codes={'depth': 5, 'thermal': 6})
    names = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    load_synth(model)
    out = {'meta': dict(vision=LORA_V, text=TINY_T, per=LORA_PER, modals=LORA_MODALS, projection_dim=64, B=B,
                        fusion_dim=32, n_classes=3, seed=3),
           'missing_index': missing_index, 'names': names,
           'trainable': [n for n, p in model.named_parameters() if p.requires_grad]}
    model.eval()
    with torch.no_grad():
        emb = model.encoder({k: dict(v) for k, v in data.items()})
        out['logits/sum'] = model({k: dict(v) for k, v in data.items()}, missing_index).clone()
    for k, v in emb.items():
        out[f'emb/{k}'] = v.clone()
    model.train()
    labels = torch.tensor([0, 1, 2, 0, 1])
    loss = torch.nn.functional.cross_entropy(model({k: dict(v) for k, v in data.items()}, missing_index), labels)
    loss.backward()
    out['labels'], out['loss/sum'] = labels, loss.detach().clone()
    gn = {}
    for n, p in model.named_parameters():
        gn[n] = float(p.grad.norm()) if p.grad is not None else None
        if p.grad is not None and ('lora_' in n or 'embeddings' in n or 'layrnorm' in n):
            out[f'grad/{n}'] = p.grad.detach().clone()
    out['grad_norms'] = gn
    torch.save(out, os.path.join(GOLD, 'tiny_lora.pt'))
    print('lora golden written: %d trainable / %d parameters, %d gradient tensors' %
          (len(out['trainable']), len(gn), sum(k.startswith('grad/') for k in out)))


PREPROC_CASES = [('image', 11, 300, 400), ('image', 12, 517, 231), ('image', 13, 100, 180), ('thermal', 14, 224, 224),
                 ('image', 15, 1080, 1920), ('depth', 16, 480, 640), ('depth', 17, 150, 97)]


def preproc():
    """the reference's OWN transforms (get_image_transform / get_thermal_transform /
    get_depth_transform, run here with this image's torchvision) on synthetic decoded images; every 7th output pixel
    is stored (32 x 32 x 3 per case) -- enough to pin resize geometry, crop offsets and normalisation."""
    import numpy as np
    import torchvision
    from PIL import Image
    ref_shim.install()
    from languagebind.image.processing_image import get_image_transform
    from languagebind.thermal.processing_thermal import get_thermal_transform
    from languagebind.depth.processing_depth import get_depth_transform
    import languagebind as lb
    cfg = lb.config_dict['depth'](text_config=dict(TINY_T), vision_config=dict(TINY_V), projection_dim=64)
    tfs = {'image': get_image_transform(cfg), 'thermal': get_thermal_transform(cfg), 'depth': get_depth_transform(cfg)}
    out = {'meta': dict(cases=PREPROC_CASES, stride=7, torchvision=torchvision.__version__,
                        max_depth=float(cfg.vision_config.max_depth))}
    for kind, seed, H, W in PREPROC_CASES:
        if kind == 'depth':
            y = tfs[kind](R.synth_depth(seed, H, W))
        else:
            y = tfs[kind](Image.fromarray(R.synth_image(seed, H, W)))
        assert tuple(y.shape) == (3, 224, 224), y.shape
        out[f'{kind}/{seed}'] = y[:, ::7, ::7].clone()
    torch.save(out, os.path.join(GOLD, 'preproc.pt'))
    print('preproc golden written with torchvision', torchvision.__version__)


def full():
    """BASELINE.json config 1: image ViT-L/14 224 + text, forward, B = 8, one image-missing
    sample, CPU fp32; plus a B = 4 fwd+bwd step (loss and gradient norms)."""
    torch.manual_seed(0)
    bank = ref_shim.build_reference_bank(['image'], FULL_V, FULL_T, projection_dim=768)
    cfgs, tcfg = cfg_objects(FULL_V, FULL_T, {}, ['image'])
    modal_types = ['language', 'image']
    model = ref_shim.build_reference_model(bank, 'sum', modal_types, 3, dropout_prob=0.0)
    load_synth(model)
    model.eval()
    B = 8
    data = R.synth_inputs(modal_types, B, cfgs, tcfg, seed=0)
    missing_index = torch.zeros(B, dtype=torch.long)
    missing_index[3] = 4
    t0 = time.time()
    with torch.no_grad():
        emb = model.encoder({k: dict(v) for k, v in data.items()})
        logits = model({k: dict(v) for k, v in data.items()}, missing_index)
    dt = (time.time() - t0) / 2
    out = {'meta': dict(vision=FULL_V, text=FULL_T, B=B, seconds_per_forward=dt, threads=torch.get_num_threads()),
           'missing_index': missing_index, 'logits/sum': logits.clone()}
    for k, v in emb.items():
        out[f'emb/{k}'] = v.clone()
    model.train()
    data4 = {k: {kk: vv[:4] for kk, vv in v.items()} for k, v in data.items()}
    labels = torch.tensor([0, 1, 2, 0])
    t0 = time.time()
    lg = model({k: dict(v) for k, v in data4.items()}, missing_index[:4])
    loss = torch.nn.functional.cross_entropy(lg, labels)
    loss.backward()
    out['meta']['seconds_fwd_bwd_b4'] = time.time() - t0
    out['labels4'] = labels
    out['loss4'] = loss.detach().clone()
    out['logits4'] = lg.detach().clone()
    out['grad_norms4'] = {n: float(p.grad.norm()) for n, p in model.named_parameters() if p.grad is not None}
    out['grad/fusion.modal_proj.image.weight'] = model.fusion.modal_proj['image'].weight.grad.clone()
    out['grad/encoder.modality_encoder.image.embeddings.class_embedding'] = \
        model.encoder.modality_encoder['image'].embeddings.class_embedding.grad.clone()
    torch.save(out, os.path.join(GOLD, 'config1_full.pt'))
    print('full golden written; forward %.2fs, fwd+bwd(B=4) %.2fs' % (dt, out['meta']['seconds_fwd_bwd_b4']))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--full', action='store_true')
    ap.add_argument('--only-lora', action='store_true')
    ap.add_argument('--only-preproc', action='store_true')
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    preproc()
    if a.only_preproc:
        sys.exit(0)
    lora()
    if a.only_lora:
        sys.exit(0)
    tiny()
    if a.full:
        full()
