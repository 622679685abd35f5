"""Module trees of the LanguageBind towers with the reference's parameter names and shapes
(so reference / hub state dicts load by name), driving the CUDA path of autograd.py.

Mirrors (by behaviour, not by code) languagebind/image/modeling_image.py:
  CLIPEncoderLayer :65-158, CLIPEncoder :337-437, CLIPTextTransformer :458-532,
  CLIPVisionTransformer :596-672, LanguageBindImage :734-773 and `_init_weights` :179-230;
the video differences of languagebind/video/modeling_video.py (:168-264: no temporal MLP);
and the transformers 4.3x CLIPAttention / CLIPMLP / CLIP*Embeddings parameter containers.
"""
import math
import os

import torch
import torch.nn.functional as F
from torch import nn

from . import autograd as ag
from . import config as C
from . import ops


class ModelOutput(tuple):
    """(last_hidden_state, pooler_output) -- indexable like the reference's return value
    (`self.modality_encoder[key](**value)[1]`, languagebind/__init__.py:78)."""

    def __new__(cls, last_hidden_state, pooler_output):
        return super().__new__(cls, (last_hidden_state, pooler_output))

    last_hidden_state = property(lambda self: self[0])
    pooler_output = property(lambda self: self[1])


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"missm_b200: {what} is on {t.device}; the B200 path has no CPU fallback "
                           f"(move the model and its inputs to a CUDA device)")


class LoraLinear(nn.Linear):
    """Parameter container of a peft-wrapped Linear (peft 0.4/0.5 layout: the wrapper IS the nn.Linear, so the
    frozen `weight` / `bias` keep their names, plus `lora_A.default.weight` [r, in] and `lora_B.default.weight`
    [out, r]); y = W x + b + (lora_alpha / r) B(A(x)).  Created by convert_to_lora below (reference:
    modeling_image.py:775-793).  Initialised as peft does: A kaiming-uniform(a = sqrt 5), B = 0."""

    def __init__(self, in_features, out_features, r, lora_alpha, lora_dropout):
        super().__init__(in_features, out_features)
        self.r, self.scaling, self.p_drop = int(r), float(lora_alpha) / float(r), float(lora_dropout)
        self.lora_A = nn.ModuleDict({'default': nn.Linear(in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({'default': nn.Linear(r, out_features, bias=False)})
        nn.init.kaiming_uniform_(self.lora_A['default'].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B['default'].weight)

    def _load_from_state_dict(self, state_dict, prefix, *args):
        # peft >= 0.6 keeps the frozen Linear under `.base_layer.`; this tree uses the 0.4 / 0.5 layout (the wrapper IS
        # the Linear).  The reference does not pin peft, so a `final_model/*.pth` written by either loads through the
        # plain nn.Module.load_state_dict the scripts call (test.py:92, train_ddp.py:193).
        for name in ('weight', 'bias'):
            k = prefix + 'base_layer.' + name
            if k in state_dict:
                state_dict[prefix + name] = state_dict.pop(k)
        super()._load_from_state_dict(state_dict, prefix, *args)

    @property
    def A(self):
        return self.lora_A['default'].weight

    @property
    def B(self):
        return self.lora_B['default'].weight


class _Wrapped(nn.Module):
    """peft's two wrapper levels, kept only for the parameter NAMES they produce
    (`encoder.base_model.model.layers.N...`): PeftModel.base_model = LoraModel, LoraModel.model = CLIPEncoder."""
    _inner = None

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            if name == self._inner:
                raise
            return getattr(super().__getattr__(self._inner), name)


class LoraModel(_Wrapped):
    _inner = 'model'

    def __init__(self, model):
        super().__init__()
        self.model = model


class PeftModel(_Wrapped):
    _inner = 'base_model'

    def __init__(self, model):
        super().__init__()
        self.base_model = LoraModel(model)

    def _load_from_state_dict(self, state_dict, prefix, *args):
        # a checkpoint of the UNWRAPPED encoder (`...encoder.layers.N...`: lora_r = 0, or adapters merged before
        # saving) loads into the wrapped tree: its keys move under `base_model.model.`
        inner = prefix + 'base_model.model.'
        for k in [k for k in state_dict if k.startswith(prefix) and not k.startswith(prefix + 'base_model.')]:
            state_dict[inner + k[len(prefix):]] = state_dict.pop(k)
        super()._load_from_state_dict(state_dict, prefix, *args)


class CLIPAttention(nn.Module):
    def __init__(self, config, lora=None):
        super().__init__()
        d = config.hidden_size
        self.embed_dim, self.num_heads = d, config.num_attention_heads
        self.head_dim = d // self.num_heads
        if self.head_dim * self.num_heads != d:
            raise ValueError(f"embed_dim must be divisible by num_heads (got {d} and {self.num_heads})")
        if config.attention_dropout != 0.0:
            raise NotImplementedError("attention_dropout != 0 is not built (reference default 0.0, "
                                      "configuration_image.py:193)")
        lin = (lambda: nn.Linear(d, d)) if not lora else (lambda: LoraLinear(d, d, *lora))
        self.has_lora = bool(lora)
        self.k_proj = lin()
        self.v_proj = lin()
        self.q_proj = lin()
        self.out_proj = lin()

    def lora_params(self):
        """(A, B) of q, k, v, out in the order autograd.LoraAttnBlockFn takes them, then the scaling."""
        out = []
        for m in (self.q_proj, self.k_proj, self.v_proj, self.out_proj):
            out += [m.A, m.B]
        return out + [self.q_proj.scaling]


class CLIPMLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        if config.hidden_act != "quick_gelu":
            raise NotImplementedError(f"hidden_act={config.hidden_act!r}: only quick_gelu is built "
                                      f"(reference default, configuration_image.py:79,191)")
        self.fc1 = nn.Linear(config.hidden_size, config.intermediate_size)
        self.fc2 = nn.Linear(config.intermediate_size, config.hidden_size)


class CLIPEncoderLayer(nn.Module):
    def __init__(self, config, temporal_mlp=True):
        super().__init__()
        d = config.hidden_size
        self.embed_dim = d
        self.eps = config.layer_norm_eps
        # convert_to_lora (modeling_image.py:775-793): with temporal attention the adapters sit on temporal_attn.*
        # (and temporal_mlp.fc1 / fc2 where that block exists), otherwise on every *.{q,k,v,out}_proj
        lora = None
        if getattr(config, 'lora_r', 0):
            lora = (config.lora_r, config.lora_alpha, config.lora_dropout)
        time_attn = bool(getattr(config, 'add_time_attn', False))
        self.self_attn = CLIPAttention(config, None if time_attn else lora)
        self.layer_norm1 = nn.LayerNorm(d, eps=config.layer_norm_eps)
        self.mlp = CLIPMLP(config)
        self.layer_norm2 = nn.LayerNorm(d, eps=config.layer_norm_eps)
        self.add_time_attn = bool(getattr(config, 'add_time_attn', False))
        self.has_temporal_mlp = False
        if self.add_time_attn:
            self.t = config.num_frames
            self.temporal_embedding = nn.Parameter(torch.zeros(1, config.num_frames, d))
            nn.init.normal_(self.temporal_embedding, std=d ** -0.5)
            self.temporal_attn = CLIPAttention(config, lora)
            self.temporal_layer_norm1 = nn.LayerNorm(d, eps=config.layer_norm_eps)
            if temporal_mlp and lora:
                raise NotImplementedError("LoRA on temporal_mlp.fc1 / fc2 (modeling_image.py:780-781) is not built: "
                                          "only the video tower enables temporal attention, and it has no temporal MLP")
            if temporal_mlp:   # image/audio/depth/thermal files keep it (:83-84), video dropped it
                self.has_temporal_mlp = True
                self.temporal_mlp = CLIPMLP(config)
                self.temporal_layer_norm2 = nn.LayerNorm(d, eps=config.layer_norm_eps)
        self._cache = {"sa": {}, "mlp": {}, "ta": {}, "tmlp": {}}

    @staticmethod
    def _attn_params(ln, a, temb=None):
        ps = [ln.weight, ln.bias, a.q_proj.weight, a.q_proj.bias, a.k_proj.weight, a.k_proj.bias,
              a.v_proj.weight, a.v_proj.bias, a.out_proj.weight, a.out_proj.bias]
        if temb is not None:
            ps.append(temb)
        if a.has_lora:
            ps += a.lora_params()[:-1]
        return ps

    @staticmethod
    def _mlp_params(ln, m):
        return [ln.weight, ln.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias]

    def _plan(self):
        """The layer as a chain of residual blocks (static) + its parameters in plan order."""
        hit = self.__dict__.get('_plan_cache')
        if hit is None:
            blocks, params = [], []

            def attn(which, cache, ln, a, temb):
                blocks.append(ag.BlockPlan("attn", cache, has_temb=temb is not None, lora=a.has_lora,
                                           scaling=a.q_proj.scaling if a.has_lora else 1.0, which=which))
                params.extend(self._attn_params(ln, a, temb))

            def mlp(cache, ln, m):
                blocks.append(ag.BlockPlan("mlp", cache, eps=self.eps))
                params.extend(self._mlp_params(ln, m))

            if self.add_time_attn:      # modeling_image.py:105-134 (video: no temporal MLP, modeling_video.py:235-240)
                attn("temporal", self._cache["ta"], self.temporal_layer_norm1, self.temporal_attn,
                     self.temporal_embedding if self.t != 1 else None)
                if self.has_temporal_mlp:
                    mlp(self._cache["tmlp"], self.temporal_layer_norm2, self.temporal_mlp)
            attn("spatial", self._cache["sa"], self.layer_norm1, self.self_attn, None)      # :137-146
            mlp(self._cache["mlp"], self.layer_norm2, self.mlp)                             # :148-151
            hit = self.__dict__['_plan_cache'] = (tuple(blocks), tuple(params))
        return hit

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop('_plan_cache', None)         # .to() / .cuda() may replace parameter objects
        return super()._apply(fn, *args, **kwargs)

    def run(self, x, spatial_meta, temporal_meta):
        """x: fp32 [M, D] residual stream, rows ordered (image, token)."""
        plan, params = self._plan()
        for a in (self.self_attn, getattr(self, 'temporal_attn', None)):
            if a is not None and a.has_lora and a.q_proj.p_drop != 0.0 and self.training:
                raise NotImplementedError("lora_dropout != 0 in training mode is not built (reference default 0.0, "
                                          "configuration_image.py:202)")
        return ag.encoder_layer(x, plan, {"spatial": spatial_meta, "temporal": temporal_meta}, params)


class CLIPEncoder(nn.Module):
    def __init__(self, config, temporal_mlp=True):
        super().__init__()
        self.config = config
        self.layers = nn.ModuleList([CLIPEncoderLayer(config, temporal_mlp)
                                     for _ in range(config.num_hidden_layers)])
        self.gradient_checkpointing = False


def vision_grid(config):
    """Token grid after the reference's resize_pos (modeling_image.py:797-803)."""
    if config.num_mel_bins and config.target_length:
        size = [int(config.num_mel_bins), int(config.target_length)]
    elif isinstance(config.image_size, (list, tuple)):
        size = list(config.image_size)
    else:
        size = [config.image_size, config.image_size]
    return size, [size[0] // config.patch_size, size[1] // config.patch_size]


def resize_pos_table(old, grid_size):
    """Setup-time port of the arithmetic of LanguageBind*.resize_pos (modeling_image.py:795-839):
    bicubic, antialiased resample of the patch rows of a position table to `grid_size`."""
    new_len = grid_size[0] * grid_size[1] + 1
    if new_len == old.shape[0]:
        return old
    tok, img = old[:1], old[1:]
    og = int(math.sqrt(len(img)))
    img = img.reshape(1, og, og, -1).permute(0, 3, 1, 2)
    img = F.interpolate(img.float(), size=list(grid_size), mode='bicubic', antialias=True, align_corners=False)
    img = img.permute(0, 2, 3, 1).reshape(grid_size[0] * grid_size[1], -1)
    return torch.cat([tok, img.to(old.dtype)], dim=0)


class CLIPVisionEmbeddings(nn.Module):
    def __init__(self, config, persistent_ids):
        super().__init__()
        self.config = config
        self.embed_dim = config.hidden_size
        self.patch_size = config.patch_size
        self.image_size, self.grid = vision_grid(config)
        self.class_embedding = nn.Parameter(torch.randn(self.embed_dim))
        self.patch_embedding = nn.Conv2d(config.num_channels, self.embed_dim, kernel_size=self.patch_size,
                                         stride=self.patch_size, bias=False)
        self.num_patches = self.grid[0] * self.grid[1]
        self.num_positions = self.num_patches + 1
        self.position_embedding = nn.Embedding(self.num_positions, self.embed_dim)
        self.register_buffer("position_ids", torch.arange(self.num_positions).expand((1, -1)),
                             persistent=persistent_ids)


class CLIPVisionTransformer(nn.Module):
    def __init__(self, config, temporal_mlp=True, persistent_ids=False):
        super().__init__()
        self.config = config
        if config.hidden_size // config.num_attention_heads != 64:
            raise NotImplementedError("the fused attention kernel is built for head_dim 64")
        self.embeddings = CLIPVisionEmbeddings(config, persistent_ids)
        self.pre_layrnorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.encoder = CLIPEncoder(config, temporal_mlp)
        if getattr(config, 'lora_r', 0):
            # what get_peft_model does to the encoder (modeling_image.py:793, bias="none"): adapters are the only
            # trainable parameters inside it, and its parameters move under encoder.base_model.model
            for n, p in self.encoder.named_parameters():
                if 'lora_' not in n:
                    p.requires_grad = False
            self.encoder = PeftModel(self.encoder)
        self.post_layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self._cache = {"embed": {}, "pool": {}}

    def forward(self, *args, **kwargs):
        """Returns (last_hidden_state, pooled).  With `proj` (the bank's projection Linear) the pooled
        output is already projected, L2-normalised and scaled (fused tail, languagebind/__init__.py
        :79-83).  `present_idx`/`n_present`: run only these samples (mask compaction)."""
        return run_steps(self.forward_steps(*args, **kwargs))

    def forward_steps(self, pixel_values=None, output_attentions=None, output_hidden_states=None, return_dict=None,
                      present_idx=None, n_present=None, proj=None, scale=1.0):
        """`forward` as a generator that yields after the embedding and after every encoder layer, so that the
        bank can issue the layers of several towers in lockstep (bank.LanguageBind.forward)."""
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")
        if output_attentions or output_hidden_states:
            raise NotImplementedError("attention maps / per-layer states are never materialised")
        if self.training and self.config.force_patch_dropout:
            raise NotImplementedError("force_patch_dropout != 0 is not built (default 0.0)")
        _require_cuda(pixel_values, "pixel_values")
        _require_cuda(self.pre_layrnorm.weight, "the vision tower")
        cfg, emb = self.config, self.embeddings
        if pixel_values.dim() == 7:      # modeling_image.py:630-634
            b, p, T, bs = pixel_values.shape[:4]
            pixel_values = pixel_values.reshape(b * p * bs * T, *pixel_values.shape[4:])
            B, frames_in_batch = b * p * bs, True
        elif pixel_values.dim() == 5:    # b c t h w
            B, T, frames_in_batch = pixel_values.shape[0], pixel_values.shape[2], False
        else:
            B, T, frames_in_batch = pixel_values.shape[0], 1, False
        if frames_in_batch:
            # already (b t) c h w: treat every frame as an image for patch extraction
            px, Tp, nB = pixel_values, 1, B * T
            if present_idx is not None:
                raise NotImplementedError("compaction with 7-D pixel_values")
        else:
            px, Tp, nB = pixel_values, T, B
        px = px.float().contiguous()
        if tuple(px.shape[-2:]) != tuple(emb.image_size):
            raise ValueError(f"pixel_values spatial size {tuple(px.shape[-2:])} != configured {emb.image_size}")
        n_samp = nB if n_present is None else n_present
        P, D = emb.num_patches, cfg.hidden_size
        N = P + 1
        geom = (cfg.patch_size, Tp, emb.grid[0], emb.grid[1], cfg.layer_norm_eps)
        x = ag.vision_embed(px, present_idx, n_samp, geom, self._cache["embed"], emb.class_embedding,
                                   emb.patch_embedding.weight, emb.position_embedding.weight,
                                   self.pre_layrnorm.weight, self.pre_layrnorm.bias)
        n_img = n_samp * Tp
        H = cfg.num_attention_heads
        spatial = ag.AttnMeta(H, cfg.layer_norm_eps, ops.SeqLayout.spatial(n_img, N))
        temporal = None
        if cfg.add_time_attn:
            t = cfg.num_frames
            if n_img % t:
                raise ValueError(f"{n_img} frames is not a multiple of num_frames={t}")
            temporal = ag.AttnMeta(H, cfg.layer_norm_eps, ops.SeqLayout.temporal(n_img // t, t, N),
                                   add_period=t, add_div=N)
        yield
        for layer in self.encoder.layers:
            x = layer.run(x, spatial, temporal)
            yield
        n_out = n_samp if not frames_in_batch else B
        T_pool = T
        rows = torch.arange(n_out * T_pool, device=x.device, dtype=torch.int32) * N
        pooled = _pool(self, x, rows, n_out, T_pool, self.post_layernorm, proj, scale, cfg.layer_norm_eps)
        return ModelOutput(x.view(n_img, N, D), pooled)


def run_steps(gen):
    """Drive a `forward_steps` generator to its end and hand back its return value."""
    try:
        while True:
            next(gen)
    except StopIteration as stop:
        return stop.value


def _pool(owner, x, rows, n_present, T, ln, proj, scale, eps):
    if proj is None:
        # standalone tower call: pooled = LayerNorm(rows) (mean over frames); identity "projection"
        raise NotImplementedError("call the tower through LanguageBind (projection is fused into the tail)")
    return ag.pool_proj(x, rows, n_present, T, eps, scale, owner._cache["pool"], ln.weight, ln.bias,
                               proj.weight)


class CLIPTextEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.token_embedding = nn.Embedding(config.vocab_size, config.hidden_size)
        self.position_embedding = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)),
                             persistent=False)


class CLIPTextTransformer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        if config.hidden_size // config.num_attention_heads != 64:
            raise NotImplementedError("the fused attention kernel is built for head_dim 64")
        self.embeddings = CLIPTextEmbeddings(config)
        self.encoder = CLIPEncoder(config)
        self.final_layer_norm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self._cache = {"pool": {}}

    def forward(self, *args, **kwargs):
        return run_steps(self.forward_steps(*args, **kwargs))

    def forward_steps(self, input_ids=None, attention_mask=None, position_ids=None, output_attentions=None,
                      output_hidden_states=None, return_dict=None, present_idx=None, n_present=None, proj=None,
                      scale=1.0):
        if input_ids is None:
            raise ValueError("You have to specify input_ids")
        if position_ids is not None:
            raise NotImplementedError("explicit position_ids")
        _require_cuda(input_ids, "input_ids")
        _require_cuda(self.final_layer_norm.weight, "the text tower")
        cfg = self.config
        ids = input_ids.reshape(-1, input_ids.shape[-1]).contiguous()
        B, L = ids.shape
        n_samp = B if n_present is None else n_present
        am = None
        if attention_mask is not None:
            am = attention_mask.reshape(B, L).to(torch.int64).contiguous()
        x = ag.TextEmbedFn.apply(ids, present_idx, n_samp, self.embeddings.token_embedding.weight,
                                 self.embeddings.position_embedding.weight)
        meta = ag.AttnMeta(cfg.num_attention_heads, cfg.layer_norm_eps, ops.SeqLayout.spatial(n_samp, L),
                           causal=True, key_mask=am, mask_rows=present_idx if am is not None else None)
        yield
        for layer in self.encoder.layers:
            x = layer.run(x, meta, None)
            yield
        rows = ops.argmax_rows(ids, sample_index=present_idx, n_samples=n_samp)
        pooled = _pool(self, x, rows, n_samp, 1, self.final_layer_norm, proj, scale, cfg.layer_norm_eps)
        return ModelOutput(None, pooled)


# ----------------------------------------------------------------------------------------------
# LanguageBind{Image,Video,Depth,Audio,Thermal}: containers the bank takes apart
# (languagebind/__init__.py:64-70)
# ----------------------------------------------------------------------------------------------
class _LanguageBindModel(nn.Module):
    config_class = C.LanguageBindImageConfig
    modality = 'image'
    temporal_mlp = True
    persistent_ids = False

    def __init__(self, config):
        super().__init__()
        if not isinstance(config.text_config, C.CLIPTextConfig):
            raise ValueError("config.text_config is expected to be of type CLIPTextConfig but is of type"
                             f" {type(config.text_config)}.")
        if not isinstance(config.vision_config, C.CLIPVisionConfig):
            raise ValueError("config.vision_config is expected to be of type CLIPVisionConfig but is of type"
                             f" {type(config.vision_config)}.")
        self.config = config
        tc, vc = config.text_config, config.vision_config
        self.projection_dim = config.projection_dim
        self.text_embed_dim, self.vision_embed_dim = tc.hidden_size, vc.hidden_size
        resized = bool(vc.num_mel_bins and vc.target_length)
        self.text_model = CLIPTextTransformer(tc)
        self.vision_model = CLIPVisionTransformer(vc, self.temporal_mlp, self.persistent_ids or resized)
        self.visual_projection = nn.Linear(self.vision_embed_dim, self.projection_dim, bias=False)
        self.text_projection = nn.Linear(self.text_embed_dim, self.projection_dim, bias=False)
        self.logit_scale = nn.Parameter(torch.tensor(float(config.logit_scale_init_value)))
        self.apply(self._init_weights)

    def _init_weights(self, m):
        """Same distributions as CLIPPreTrainedModel._init_weights (modeling_image.py:179-230)."""
        f = self.config.initializer_factor
        if isinstance(m, CLIPTextEmbeddings):
            m.token_embedding.weight.data.normal_(mean=0.0, std=f * 0.02)
            m.position_embedding.weight.data.normal_(mean=0.0, std=f * 0.02)
        elif isinstance(m, CLIPVisionEmbeddings):
            nn.init.normal_(m.class_embedding, mean=0.0, std=m.embed_dim ** -0.5 * f)
            nn.init.normal_(m.patch_embedding.weight, std=m.config.initializer_range * f)
            nn.init.normal_(m.position_embedding.weight, std=m.config.initializer_range * f)
        elif isinstance(m, CLIPAttention):
            n_layers = (self.config.vision_config if m.embed_dim == self.vision_embed_dim
                        else self.config.text_config).num_hidden_layers
            in_std = (m.embed_dim ** -0.5) * ((2 * n_layers) ** -0.5) * f
            for lin in (m.q_proj, m.k_proj, m.v_proj):
                nn.init.normal_(lin.weight, std=in_std)
            nn.init.normal_(m.out_proj.weight, std=(m.embed_dim ** -0.5) * f)
        elif isinstance(m, CLIPMLP):
            hidden = m.fc1.in_features
            n_layers = (self.config.vision_config if hidden == self.vision_embed_dim
                        else self.config.text_config).num_hidden_layers
            nn.init.normal_(m.fc1.weight, std=(2 * hidden) ** -0.5 * f)
            nn.init.normal_(m.fc2.weight, std=(hidden ** -0.5) * ((2 * n_layers) ** -0.5) * f)
        elif isinstance(m, _LanguageBindModel):
            nn.init.normal_(m.text_projection.weight, std=m.text_embed_dim ** -0.5 * f)
            nn.init.normal_(m.visual_projection.weight, std=m.vision_embed_dim ** -0.5 * f)
        if isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)
        if isinstance(m, nn.Linear) and m.bias is not None:
            m.bias.data.zero_()

    @classmethod
    def synthetic_config(cls):
        vc = dict(C.VIT_L14)
        vc.update(C.SYNTHETIC_PER_MODALITY[cls.modality])
        return cls.config_class(text_config=dict(C.CLIP_TEXT), vision_config=vc, projection_dim=768)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, cache_dir=None, **kwargs):
        """Local-only stand-in for PreTrainedModel.from_pretrained (called at
        languagebind/__init__.py:64).  MISSM_SYNTHETIC=1 builds the synthetic ViT-L/14 model of
        SURVEY.md section 8(d) with reference-style random init (there is no network here)."""
        if os.environ.get("MISSM_SYNTHETIC", "0") == "1":
            torch.manual_seed(1234)
            return cls(cls.synthetic_config())
        path = C.resolve_checkpoint_dir(pretrained_model_name_or_path, cache_dir)
        model = cls(cls.config_class.from_json_file(os.path.join(path, "config.json")))
        model.load_checkpoint_dir(path)
        return model

    def load_checkpoint_dir(self, path):
        sd = None
        st = os.path.join(path, "model.safetensors")
        if os.path.isfile(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            for name in ("pytorch_model.bin", "model.pth"):
                if os.path.isfile(os.path.join(path, name)):
                    sd = torch.load(os.path.join(path, name), map_location="cpu")
                    break
        if sd is None:
            raise FileNotFoundError(f"no weights file in {path}")
        self.load_reference_state_dict(sd)

    def load_reference_state_dict(self, sd):
        """Load a reference/hub state dict by name; position tables of another grid are resampled the
        way resize_pos does, `position_ids` buffers are ignored.  Encoder keys are accepted in the three layouts a
        reference checkpoint can have: plain (`vision_model.encoder.layers.N...`,
        lora_r = 0 or adapters merged before saving), peft-wrapped in the 0.4/0.5 layout this module tree uses
        (`vision_model.encoder.base_model.model.layers.N...q_proj.weight` + `.lora_A.default.weight`), and the
        peft >= 0.6 layout of the same thing (`...q_proj.base_layer.weight`).  A plain checkpoint loads into a
        LoRA-configured model with the adapters left at their no-op initialisation (B = 0)."""
        sd = {k.replace(".base_layer.", "."): v for k, v in sd.items() if not k.endswith("position_ids")}
        wrapped = isinstance(self.vision_model.encoder, PeftModel)
        plain_p, peft_p = "vision_model.encoder.", "vision_model.encoder.base_model.model."
        has_peft_keys = any(k.startswith(peft_p) for k in sd)
        keep_adapters = False
        if wrapped and not has_peft_keys:
            sd = {(peft_p + k[len(plain_p):] if k.startswith(plain_p) else k): v for k, v in sd.items()}
            keep_adapters = True
        elif has_peft_keys and not wrapped:
            raise RuntimeError("the checkpoint holds a peft-wrapped encoder (LoRA adapters) but the model was built "
                               "with lora_r = 0: build it from the checkpoint's own config.json")
        k = "vision_model.embeddings.position_embedding.weight"
        if k in sd and sd[k].shape[0] != self.vision_model.embeddings.num_positions:
            sd[k] = resize_pos_table(sd[k], self.vision_model.embeddings.grid)
        missing, unexpected = self.load_state_dict(sd, strict=False)
        missing = [m for m in missing if not m.endswith("position_ids") and not (keep_adapters and ".lora_" in m)]
        if missing or unexpected:
            raise RuntimeError(f"state dict mismatch: missing={missing[:5]} unexpected={unexpected[:5]}")


class LanguageBindImage(_LanguageBindModel):
    config_class, modality = C.LanguageBindImageConfig, 'image'


class LanguageBindDepth(_LanguageBindModel):
    config_class, modality = C.LanguageBindDepthConfig, 'depth'


class LanguageBindThermal(_LanguageBindModel):
    config_class, modality = C.LanguageBindThermalConfig, 'thermal'


class LanguageBindAudio(_LanguageBindModel):
    config_class, modality = C.LanguageBindAudioConfig, 'audio'


class LanguageBindVideo(_LanguageBindModel):
    # video/modeling_video.py: own CLIPVisionEmbeddings copy (persistent position_ids, :44),
    # temporal attention without the temporal MLP (:189-190, :235-240)
    config_class, modality = C.LanguageBindVideoConfig, 'video'
    temporal_mlp = False
    persistent_ids = True
