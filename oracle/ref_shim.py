"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference Python files.

Used (a) inside the build container (where /root/reference exists) by oracle/make_golden.py to generate golden
vectors and to pin oracle/restatement.py, and (b) by bench.py's CPU legs (`--impl reference`, `cpu_baseline`), which
time the reference's own code on the host cores: on the GPU box the reference is the byte-identical copy that
oracle/stage_reference.py staged into the git-ignored oracle/_ref/.  Nothing in the product path imports it.

Why a shim is needed: the reference binds third-party
names at import time (languagebind/image/modeling_image.py:5-15) that do not exist in this
image: `peft` (LoRA re-stated below), `decord`, `pytorchvideo`, `torch_geometric`, and four symbols of a
transformers-4.3x-era `modeling_clip` (`_expand_mask`, the 4.3x `CLIPAttention` calling
convention with `causal_attention_mask=`, and a `CLIPVisionEmbeddings` without the
square-input check).  The third-party arithmetic is re-stated here from its published
algorithm (transformers 4.31-4.34 `modeling_clip.py`; version unpinned by the reference).
"""
import os
import sys
import types

import torch
from torch import nn

_HERE = os.path.dirname(os.path.abspath(__file__))
# the reference tree itself in the build container, else the copy staged by oracle/stage_reference.py
REFERENCE_ROOT = os.environ.get("MISSM_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/languagebind") else os.path.join(_HERE, "_ref"))


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "languagebind")) and \
        os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "model", "baseline.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("stubbed third-party symbol (not installed in this image)")


def _expand_mask(mask, dtype, tgt_len=None):
    """transformers 4.3x `_expand_mask`: [B,S] -> [B,1,T,S]; masked positions = finfo.min."""
    bsz, src_len = mask.size()
    tgt_len = tgt_len if tgt_len is not None else src_len
    expanded = mask[:, None, None, :].expand(bsz, 1, tgt_len, src_len).to(dtype)
    inverted = 1.0 - expanded
    return inverted.masked_fill(inverted.to(torch.bool), torch.finfo(dtype).min)


class CLIPAttention(nn.Module):
    """transformers 4.3x CLIPAttention (call sites: modeling_image.py:69,81,121,140)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embed_dim = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = self.embed_dim // self.num_heads
        self.scale = self.head_dim ** -0.5
        self.dropout = config.attention_dropout
        self.k_proj = nn.Linear(self.embed_dim, self.embed_dim)
        self.v_proj = nn.Linear(self.embed_dim, self.embed_dim)
        self.q_proj = nn.Linear(self.embed_dim, self.embed_dim)
        self.out_proj = nn.Linear(self.embed_dim, self.embed_dim)

    def _shape(self, t, n, b):
        return t.view(b, n, self.num_heads, self.head_dim).transpose(1, 2).contiguous()

    def forward(self, hidden_states, attention_mask=None, causal_attention_mask=None,
                output_attentions=False):
        b, n, d = hidden_states.size()
        q = self.q_proj(hidden_states) * self.scale
        k = self._shape(self.k_proj(hidden_states), -1, b)
        v = self._shape(self.v_proj(hidden_states), -1, b)
        shp = (b * self.num_heads, -1, self.head_dim)
        q = self._shape(q, n, b).view(*shp)
        k = k.view(*shp)
        v = v.view(*shp)
        w = torch.bmm(q, k.transpose(1, 2))
        if causal_attention_mask is not None:
            w = (w.view(b, self.num_heads, n, n) + causal_attention_mask).view(b * self.num_heads, n, n)
        if attention_mask is not None:
            w = (w.view(b, self.num_heads, n, n) + attention_mask).view(b * self.num_heads, n, n)
        w = nn.functional.softmax(w, dim=-1)
        p = nn.functional.dropout(w, p=self.dropout, training=self.training)
        o = torch.bmm(p, v)
        o = o.view(b, self.num_heads, n, self.head_dim).transpose(1, 2).reshape(b, n, d)
        return self.out_proj(o), None


class CLIPVisionEmbeddings(nn.Module):
    """Same algorithm as the reference's own copy at video/modeling_video.py:19-51."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embed_dim = config.hidden_size
        self.image_size = config.image_size
        self.patch_size = config.patch_size
        self.class_embedding = nn.Parameter(torch.randn(self.embed_dim))
        self.patch_embedding = nn.Conv2d(config.num_channels, self.embed_dim,
                                         kernel_size=self.patch_size, stride=self.patch_size,
                                         bias=False)
        self.num_patches = (self.image_size // self.patch_size) ** 2
        self.num_positions = self.num_patches + 1
        self.position_embedding = nn.Embedding(self.num_positions, self.embed_dim)
        self.register_buffer("position_ids", torch.arange(self.num_positions).expand((1, -1)),
                             persistent=False)

    def forward(self, pixel_values):
        bsz = pixel_values.shape[0]
        pe = self.patch_embedding(pixel_values).flatten(2).transpose(1, 2)
        cls = self.class_embedding.expand(bsz, 1, -1)
        emb = torch.cat([cls, pe], dim=1)
        return emb + self.position_embedding(self.position_ids)


# ---------------------------------------------------------------------------------------------
# peft (third-party, unpinned by the reference, not installed here): LoRA re-stated from its published
# algorithm (peft 0.4/0.5 layout, contemporary with the transformers 4.31-4.34 API the reference needs):
#   get_peft_model(model, LoraConfig) -> PeftModel(base_model=LoraModel(model=model)); every nn.Linear
#   whose qualified name ends with a target is replaced by a Linear subclass that keeps `weight`/`bias`
#   and adds lora_A.default [r, in] (kaiming-uniform, a = sqrt 5), lora_B.default [out, r] (zeros):
#   y = W x + b + (lora_alpha / r) * B(A(dropout(x)));  bias="none": every parameter without "lora_" in its
#   name is frozen.  Call site: modeling_image.py:775-793 (`convert_to_lora`).
# ---------------------------------------------------------------------------------------------
class LoraConfig:
    def __init__(self, r=8, lora_alpha=8, target_modules=None, lora_dropout=0.0, bias="none",
                 modules_to_save=None, **kwargs):
        self.r, self.lora_alpha, self.target_modules = r, lora_alpha, list(target_modules or [])
        self.lora_dropout, self.bias, self.modules_to_save = lora_dropout, bias, modules_to_save


class LoraLinear(nn.Linear):
    def __init__(self, base, r, lora_alpha, lora_dropout):
        super().__init__(base.in_features, base.out_features, bias=base.bias is not None)
        self.weight, self.bias = base.weight, base.bias
        self.lora_dropout = nn.ModuleDict({'default': nn.Dropout(lora_dropout) if lora_dropout > 0 else nn.Identity()})
        self.lora_A = nn.ModuleDict({'default': nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({'default': nn.Linear(r, base.out_features, bias=False)})
        self.scaling = lora_alpha / r
        nn.init.kaiming_uniform_(self.lora_A['default'].weight, a=5 ** 0.5)
        nn.init.zeros_(self.lora_B['default'].weight)

    def forward(self, x):
        y = nn.functional.linear(x, self.weight, self.bias)
        return y + self.lora_B['default'](self.lora_A['default'](self.lora_dropout['default'](x))) * self.scaling


class _Delegate(nn.Module):
    _inner = None

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__(self._inner), name)

    def forward(self, *a, **k):
        return getattr(self, self._inner)(*a, **k)


class LoraModel(_Delegate):
    _inner = 'model'

    def __init__(self, model, config):
        super().__init__()
        self.model = model
        for name, mod in list(model.named_modules()):
            if isinstance(mod, nn.Linear) and any(name == t or name.endswith('.' + t) for t in config.target_modules):
                parent = model.get_submodule(name.rsplit('.', 1)[0]) if '.' in name else model
                setattr(parent, name.rsplit('.', 1)[-1], LoraLinear(mod, config.r, config.lora_alpha, config.lora_dropout))
        for n, p in model.named_parameters():
            if 'lora_' not in n:
                p.requires_grad = False


class PeftModel(_Delegate):
    _inner = 'base_model'

    def __init__(self, model, config):
        super().__init__()
        self.base_model = LoraModel(model, config)


def get_peft_model(model, config):
    return PeftModel(model, config)


_installed = False


def install():
    """Install stubs and re-stated symbols, put the reference on sys.path."""
    global _installed
    if _installed:
        return
    _stub("peft", LoraConfig=LoraConfig, get_peft_model=get_peft_model)
    _stub("decord", VideoReader=_Unavailable, cpu=lambda *a, **k: None,
          bridge=types.SimpleNamespace(set_bridge=lambda *a, **k: None))
    _stub("pytorchvideo")
    _stub("pytorchvideo.data")
    _stub("pytorchvideo.data.encoded_video", EncodedVideo=_Unavailable)
    _stub("pytorchvideo.transforms", ApplyTransformToKey=_Unavailable,
          ShortSideScale=_Unavailable, UniformTemporalSubsample=_Unavailable)
    _stub("torch_geometric")
    _stub("torch_geometric.nn", SuperGATConv=_Unavailable)
    _stub("torch_geometric.data", Batch=_Unavailable, Data=_Unavailable)
    try:
        import torchaudio
        if not hasattr(torchaudio, "set_audio_backend"):
            torchaudio.set_audio_backend = lambda *a, **k: None
    except Exception:  # torchaudio itself absent
        ta = _stub("torchaudio", set_audio_backend=lambda *a, **k: None)
        ta.compliance = types.SimpleNamespace(kaldi=types.SimpleNamespace())
        _stub("torchaudio.compliance")
        _stub("torchaudio.compliance.kaldi")
    from transformers import PretrainedConfig
    if not hasattr(PretrainedConfig, "_set_token_in_kwargs"):
        PretrainedConfig._set_token_in_kwargs = classmethod(lambda cls, kwargs, token=None: None)
    from transformers.models.clip import modeling_clip as mc
    mc._expand_mask = _expand_mask
    mc.CLIPAttention = CLIPAttention
    mc.CLIPVisionEmbeddings = CLIPVisionEmbeddings
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def build_reference_bank(modalities, vision_cfg, text_cfg, projection_dim=768, use_temp=True,
                         per_modality_cfg=None):
    """Run the reference's LanguageBind.__init__ (languagebind/__init__.py:55-73) unmodified,
    with `from_pretrained` replaced by construction from a synthetic config (no network)."""
    install()
    import languagebind as lb  # the reference package
    assert lb.__file__.startswith(REFERENCE_ROOT), lb.__file__
    per_modality_cfg = per_modality_cfg or {}
    for k in modalities:
        cls = lb.model_dict[k]
        vc = dict(vision_cfg)
        vc.update(per_modality_cfg.get(k, {}))

        def _fp(klass, name, cache_dir=None, _vc=vc, **kw):
            cfg = klass.config_class(text_config=dict(text_cfg), vision_config=_vc,
                                     projection_dim=projection_dim)
            return klass(cfg)

        cls.from_pretrained = classmethod(_fp)
    bank = lb.LanguageBind({k: f"synthetic_{k}" for k in modalities}, use_temp=use_temp)
    return bank


def build_reference_model(bank, fusion_type, modality_types, n_classes, feature_dims=768,
                          fusion_dim=256, dropout_prob=0.1, extra_missing_codes=None):
    """finetune_model (src/model/baseline.py:421-453) on top of a reference bank."""
    install()
    from src.model import baseline as ref_baseline
    assert ref_baseline.__file__.startswith(REFERENCE_ROOT)
    if extra_missing_codes:
        # depth / thermal have no code in the reference (baseline.py:8); BASELINE.json configs
        # 2 and 4 use them, so the map is extended without touching codes 0-4.
        ref_baseline.missing_type_index.update(extra_missing_codes)
    args = types.SimpleNamespace(fusion_type=fusion_type, modality_types=list(modality_types),
                                 feature_dims=feature_dims, fusion_dim=fusion_dim,
                                 dropout_prob=dropout_prob)
    return ref_baseline.finetune_model(args, n_classes, bank)
