"""Builders shared by tests, bench.py and smoke(): construct the product's bank / finetune_model
from plain config dictionaries (no hub access), and enumerate its parameter names and shapes."""
import types

import torch

from . import config as C
from . import towers as T
from .bank import LanguageBind

_VISION_KEYS = ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "num_channels",
                "image_size", "patch_size", "hidden_act", "layer_norm_eps", "add_time_attn", "num_frames",
                "num_mel_bins", "target_length", "lora_r", "lora_alpha", "lora_dropout")
_TEXT_KEYS = ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
              "max_position_embeddings", "hidden_act", "layer_norm_eps")


def _as_dict(cfg, keys):
    if isinstance(cfg, dict):
        return {k: cfg[k] for k in keys if k in cfg}
    return {k: getattr(cfg, k) for k in keys if hasattr(cfg, k)}


def build_bank(vision_cfgs, text_cfg, projection_dim=768, use_temp=True):
    """vision_cfgs: {modality: dict | namespace}; the text tower is attached to the last model, as in
    languagebind/__init__.py:69-70."""
    models = {}
    mods = {'image': T.LanguageBindImage, 'video': T.LanguageBindVideo, 'audio': T.LanguageBindAudio,
            'depth': T.LanguageBindDepth, 'thermal': T.LanguageBindThermal}
    for m, vc in vision_cfgs.items():
        cls = mods[m]
        vd = _as_dict(vc, _VISION_KEYS)
        vd.setdefault("lora_r", 0)      # the builders default to the plain encoder (reference default: 2)
        cfg = cls.config_class(text_config=_as_dict(text_cfg, _TEXT_KEYS), vision_config=vd,
                               projection_dim=projection_dim)
        models[m] = cls(cfg)
    return LanguageBind.from_models(models, use_temp=use_temp)


def build_finetune(vision_cfgs, text_cfg, modality_types, fusion_type, n_classes=3, projection_dim=768,
                   fusion_dim=256, dropout_prob=0.0, use_temp=True):
    from src.model.baseline import finetune_model
    bank = build_bank(vision_cfgs, text_cfg, projection_dim, use_temp)
    args = types.SimpleNamespace(fusion_type=fusion_type, modality_types=list(modality_types),
                                 feature_dims=projection_dim, fusion_dim=fusion_dim, dropout_prob=dropout_prob)
    return finetune_model(args, n_classes, bank)


def reference_named_shapes(vision_cfgs, text_cfg, modality_types, fusion_type, projection_dim=768,
                           fusion_dim=256, n_classes=3):
    """(name, shape) of every entry of the product's state dict -- by construction the reference's
    names (`encoder.modality_encoder.<m>...`, `fusion...`); built on the meta device (no memory)."""
    vis = {m: vision_cfgs[m] for m in modality_types if m != 'language'}
    with torch.device('meta'):
        model = build_finetune(vis, text_cfg, modality_types, fusion_type, n_classes, projection_dim, fusion_dim)
    return [(k, tuple(v.shape)) for k, v in model.state_dict().items()]


def load_named(model, sd):
    """Copy a {name: tensor} dict into the model's parameters / buffers by name (strict)."""
    own = dict(model.state_dict())
    missing = [k for k in own if k not in sd and not k.endswith('position_ids')]
    if missing:
        raise KeyError(f"missing {missing[:4]}")
    with torch.no_grad():
        for k, v in own.items():
            if k in sd:
                v.copy_(sd[k].to(v.device))
