"""CPU-only tests of the host side: the drop-in module tree carries the reference's parameter
names and shapes, configs mirror the reference defaults, the product refuses to run on the CPU
(no fallback), and the benchmark's synthetic problem follows the reference's generators."""
import os
import sys
import types

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import restatement as R  # noqa: E402

GOLD = os.path.join(HERE, "golden")


def test_dropin_import_surface():
    import languagebind as lb
    from src.model.baseline import finetune_model, missing_type_index  # noqa: F401
    for name in ("LanguageBind", "to_device", "transform_dict", "LanguageBindImageTokenizer", "model_dict",
                 "config_dict"):
        assert hasattr(lb, name), name
    assert set(lb.model_dict) == {'thermal', 'image', 'video', 'depth', 'audio'} == set(lb.transform_dict)
    assert {k: missing_type_index[k] for k in ('language', 'video', 'audio', 'image')} == \
        {'language': 1, 'video': 2, 'audio': 3, 'image': 4}           # src/model/baseline.py:8 unchanged


def test_parameter_names_and_shapes_match_reference():
    """tests/golden/reference_param_shapes.pt holds state-dict names/shapes of the UNMODIFIED
    reference modules (oracle/make_golden.py); the product's module tree must be identical."""
    path = os.path.join(GOLD, "reference_param_shapes.pt")
    if not os.path.exists(path):
        pytest.skip("name list not generated")
    from missm_b200 import shapes
    ref = torch.load(path, weights_only=False)
    meta = ref['meta']
    cfgs = {}
    for m in meta['modals']:
        d = {k: v for k, v in meta['vision'].items() if k != 'lora_r'}
        d.update(meta['per'].get(m, {}))
        cfgs[m] = R.vision_config(**d)
    tcfg = R.text_config(**meta['text'])
    for fusion, names in ref['fusions'].items():
        mine = dict(shapes.reference_named_shapes(cfgs, tcfg, ['language'] + meta['modals'], fusion,
                                                  projection_dim=64, fusion_dim=32))
        theirs = {k: tuple(v) for k, v in names.items() if not k.endswith('position_ids')}
        mine = {k: v for k, v in mine.items() if not k.endswith('position_ids')}
        assert set(mine) == set(theirs), (fusion, sorted(set(mine) ^ set(theirs))[:8])
        for k in theirs:
            assert mine[k] == theirs[k], (fusion, k, mine[k], theirs[k])


def test_config_defaults_mirror_reference():
    from missm_b200 import config as C
    v, t = C.CLIPVisionConfig(), C.CLIPTextConfig()
    assert (v.hidden_size, v.intermediate_size, v.num_hidden_layers, v.patch_size, v.hidden_act, v.lora_r,
            v.layer_norm_eps, v.add_time_attn, v.num_frames) == (768, 3072, 12, 32, "quick_gelu", 2, 1e-5, False, 1)
    assert (t.vocab_size, t.hidden_size, t.max_position_embeddings, t.eos_token_id) == (49408, 512, 77, 49407)
    c = C.LanguageBindAudioConfig(vision_config=dict(num_mel_bins=112, target_length=1036, patch_size=14))
    assert c.logit_scale_init_value == 2.6592 and c.vision_config.target_length == 1036
    with pytest.raises(ValueError):
        from missm_b200 import towers
        towers.LanguageBindImage(types.SimpleNamespace(text_config={}, vision_config={}, projection_dim=8,
                                                       logit_scale_init_value=1.0, initializer_factor=1.0))


def test_audio_grid_and_position_table_shapes():
    from missm_b200 import towers, config as C
    with torch.device('meta'):
        m = towers.LanguageBindAudio(towers.LanguageBindAudio.synthetic_config())
    assert m.vision_model.embeddings.position_embedding.weight.shape == (593, 1024)   # 8 x 74 + 1
    table = R.synth_param('t', (17, 128), 0.5)
    assert torch.allclose(towers.resize_pos_table(table, [2, 5]), R.resize_pos(table, [2, 5]))


def test_product_has_no_cpu_fallback():
    from missm_b200 import shapes
    v = dict(hidden_size=128, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2, patch_size=14,
             image_size=28)
    t = dict(hidden_size=128, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2, vocab_size=64)
    model = shapes.build_finetune({'image': v}, t, ['image'], 'sum', 3, 64, 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model({'image': {'pixel_values': torch.randn(2, 3, 28, 28)}}, torch.zeros(2, dtype=torch.long))


def test_unknown_heads_are_refused():
    from src.model.baseline import finetune_model
    args = types.SimpleNamespace(fusion_type='no_such_head', modality_types=['image'], feature_dims=8, fusion_dim=8,
                                 dropout_prob=0.0)
    with pytest.raises(ValueError):
        finetune_model(args, 3, torch.nn.Identity())


@pytest.mark.parametrize("fusion", ["graph_fusion", "unified_graph"])
def test_graph_heads_dense_form_matches_edge_list_oracle(fusion):
    """The graph heads (reference baseline.py:240-331) as dense per-sample masked attention vs the oracle's edge-list
    restatement of torch_geometric's SuperGATConv: every missing pattern incl. an isolated node; not skip-safe, so the
    bank must NOT be asked to compact for them.  (torch_geometric is absent: parity unpinned for that third party.)"""
    from src.model import baseline as B
    mods = ['language', 'video', 'audio', 'image']
    P, Fd = 768, 256                                   # fusion_gcn() hard-codes 256 / 768 (:253, :293-294)
    args = types.SimpleNamespace(fusion_type=fusion, modality_types=mods, feature_dims=P, fusion_dim=Fd, dropout_prob=0.0)
    head = B._FUSIONS[fusion](args, 3).eval()
    assert head.compaction_safe is False
    sd = R.synth_state_dict([(k, tuple(v.shape)) for k, v in head.state_dict().items()])
    for k in sd:
        if k.endswith('att_l') or k.endswith('att_r'):
            sd[k] = R.synth_param(k, sd[k].shape, 0.3)
    head.load_state_dict(sd)
    g = torch.Generator().manual_seed(0)
    emb = {m: torch.randn(6, P, generator=g) for m in mods}
    mi = torch.tensor([0, 1, 2, 3, 4, 0])
    with torch.no_grad():
        out = head({k: v.clone() for k, v in emb.items()}, mi)
        ref = R.fusion_forward({'fusion.' + k: v for k, v in sd.items()}, fusion, mods, emb, mi)
    assert ((out - ref).norm() / ref.norm()).item() < 1e-5
    # a missing modality's embedding still matters (self-loop): changing it changes the logits of that sample only
    emb2 = {k: v.clone() for k, v in emb.items()}
    emb2['video'][2] += 1.0
    with torch.no_grad():
        out2 = head(emb2, mi)
    changed = (out2 - out).abs().max(dim=1).values > 1e-6
    # (unified_graph replaces the missing embedding by the completion GCN's output for that node, :316-318 -- but
    #  a node without present neighbours only has its self-loop, so that output is a function of the same embedding)
    assert changed.tolist() == [False, False, True, False, False, False]
    # names follow the reference's module tree
    keys = set(head.state_dict())
    pre = 'gcn.' if fusion == 'graph_fusion' else 'complete_gcn.'
    assert {pre + 'gat1.att_l', pre + 'gat1.att_r', pre + 'gat1.bias', pre + 'gat1.lin.weight', pre + 'gat2.lin.weight'} <= keys


def test_return_arity_per_fusion_type():
    """train_ddp.py:232-249 unpacks tuples for the distillation heads."""
    from src.model import baseline as B
    args = types.SimpleNamespace(modality_types=['a', 'b'], feature_dims=8, fusion_dim=8, dropout_prob=0.0)
    B.missing_type_index.update({'a': 11, 'b': 12})
    try:
        batch = {'a': torch.randn(3, 8), 'b': torch.randn(3, 8)}
        mi = torch.tensor([0, 11, 12])
        d = B.modal_distillation(args, 3)
        feats, logits = d(batch, mi)
        assert feats.shape == (3, 16) and logits.shape == (3, 3) and feats[1, :8].abs().sum() == 0
        s = B.modal_self_distillation(args, 3).train()
        masks, stu, tea, logits = s(batch, mi)
        assert len(masks) == 2 and len(stu) == 2 and tea.shape == (3, 8) and logits.shape == (3, 3)
        assert s.eval()(batch, mi).shape == (3, 3)
        c = B.modal_concat(args, 3)
        c.set_statistics({'a': [1.0] * 8, 'b': [2.0] * 8}, ['a', 'b'])
        assert c.statistics_a.tolist() == [1.0] * 8 and c(batch, mi).shape == (3, 3)
    finally:
        B.missing_type_index.pop('a'), B.missing_type_index.pop('b')


def test_synthetic_inputs_follow_the_loader_contract():
    cfgs = {'video': R.vision_config(patch_size=14, image_size=28, add_time_attn=True, num_frames=8),
            'audio': R.vision_config(patch_size=14, num_mel_bins=112, target_length=1036)}
    t = R.text_config()
    d = R.synth_inputs(['language', 'video', 'audio'], 3, cfgs, t)
    assert d['video']['pixel_values'].shape == (3, 3, 8, 28, 28)           # b c t h w
    assert d['audio']['pixel_values'].shape == (3, 3, 112, 1036)
    ids, am = d['language']['input_ids'], d['language']['attention_mask']
    assert ids.shape == (3, 77) and ids.dtype == torch.int64 and am[:, :21].all() and not am[:, 21:].any()
    assert (ids.argmax(-1) == 20).all()                                     # first EOT


def test_host_side_tower_batch_sizes_match_the_compaction_contract():
    """When `missing_index` arrives on the host (the end-to-end call with pinned host buffers) the bank counts the
    present samples per tower there instead of reading the kernel's counts back (one host sync less): the host count
    must be exactly what missm_compact_mask reports (contract: present = missing_index != code; tests/ops_emulation.py
    restates it, the kernel itself is checked bit-exactly in test_kernels_gpu.py)."""
    import ops_emulation as E
    from missm_b200.bank import MISSING_TYPE_INDEX
    g = torch.Generator().manual_seed(3)
    keys = ['image', 'depth', 'thermal', 'video', 'audio', 'language', 'unknown_modality']
    codes = [MISSING_TYPE_INDEX.get(k, -1) for k in keys]
    for B in (1, 7, 64, 257):
        mi = torch.randint(0, 7, (B,), generator=g, dtype=torch.int64)
        _, _, counts = E.compact_mask(mi, codes)
        assert [int((mi != c).sum()) for c in codes] == counts.tolist()
    assert codes[-1] == -1      # a modality without a code is never "missing": all samples present
