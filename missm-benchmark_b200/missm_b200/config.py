"""Configuration objects with the attribute names and defaults of the reference's
`languagebind/*/configuration_*.py` (CLIPTextConfig :16-123, CLIPVisionConfig :128-250 and the
composite LanguageBind*Config :253-413 of configuration_image.py; depth adds `max_depth`
(configuration_depth.py:205), audio adds `audio_sample_rate/mean/std`
(configuration_audio.py:206-208)).

They are plain Python (no transformers.PretrainedConfig): the hot path only reads shapes from
them, and the processors (`train_ddp.py:179`) only read attributes.
"""
import copy
import json
import os


class _Config:
    model_type = ""

    def to_dict(self):
        out = {}
        for k, v in self.__dict__.items():
            out[k] = v.to_dict() if isinstance(v, _Config) else copy.deepcopy(v)
        out["model_type"] = self.model_type
        return out

    def __repr__(self):
        return f"{type(self).__name__} {json.dumps(self.to_dict(), indent=2, sort_keys=True, default=str)}"

    # transformers-style accessors some callers use
    @property
    def use_return_dict(self):
        return True


class CLIPTextConfig(_Config):
    model_type = "clip_text_model"

    def __init__(self, vocab_size=49408, hidden_size=512, intermediate_size=2048, projection_dim=512,
                 num_hidden_layers=12, num_attention_heads=8, max_position_embeddings=77,
                 hidden_act="quick_gelu", layer_norm_eps=1e-5, attention_dropout=0.0,
                 initializer_range=0.02, initializer_factor=1.0, pad_token_id=1, bos_token_id=49406,
                 eos_token_id=49407, **kwargs):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.projection_dim = projection_dim
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.max_position_embeddings = max_position_embeddings
        self.layer_norm_eps = layer_norm_eps
        self.hidden_act = hidden_act
        self.initializer_range = initializer_range
        self.initializer_factor = initializer_factor
        self.attention_dropout = attention_dropout
        self.pad_token_id, self.bos_token_id, self.eos_token_id = pad_token_id, bos_token_id, eos_token_id
        self.add_time_attn = False
        self.extra = {k: v for k, v in kwargs.items() if k != "model_type"}


class CLIPVisionConfig(_Config):
    model_type = "clip_vision_model"

    def __init__(self, hidden_size=768, intermediate_size=3072, projection_dim=512, num_hidden_layers=12,
                 num_attention_heads=12, num_channels=3, image_size=224, patch_size=32,
                 hidden_act="quick_gelu", layer_norm_eps=1e-5, attention_dropout=0.0,
                 initializer_range=0.02, initializer_factor=1.0, add_time_attn=False, num_frames=1,
                 force_patch_dropout=0.0, lora_r=2, lora_alpha=16, lora_dropout=0.0, num_mel_bins=0.0,
                 target_length=0.0, video_decode_backend='decord', max_depth=10,
                 audio_sample_rate=16000, audio_mean=0.5, audio_std=0.5, **kwargs):
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.projection_dim = projection_dim
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.num_channels = num_channels
        self.patch_size = patch_size
        self.image_size = image_size
        self.initializer_range = initializer_range
        self.initializer_factor = initializer_factor
        self.attention_dropout = attention_dropout
        self.layer_norm_eps = layer_norm_eps
        self.hidden_act = hidden_act
        self.add_time_attn = add_time_attn
        self.num_frames = num_frames
        self.force_patch_dropout = force_patch_dropout
        self.lora_r = lora_r
        self.lora_alpha = lora_alpha
        self.lora_dropout = lora_dropout
        self.num_mel_bins = num_mel_bins
        self.target_length = target_length
        self.video_decode_backend = video_decode_backend
        self.max_depth = max_depth
        self.audio_sample_rate = audio_sample_rate
        self.audio_mean = audio_mean
        self.audio_std = audio_std
        self.extra = {k: v for k, v in kwargs.items() if k != "model_type"}


class _LanguageBindConfig(_Config):
    is_composition = True

    def __init__(self, text_config=None, vision_config=None, projection_dim=512,
                 logit_scale_init_value=2.6592, **kwargs):
        if isinstance(text_config, _Config):
            text_config = text_config.to_dict()
        if isinstance(vision_config, _Config):
            vision_config = vision_config.to_dict()
        text_config = dict(text_config or {})
        vision_config = dict(vision_config or {})
        text_config.pop("extra", None), vision_config.pop("extra", None)
        self.text_config = CLIPTextConfig(**text_config)
        self.vision_config = CLIPVisionConfig(**vision_config)
        self.projection_dim = projection_dim
        self.logit_scale_init_value = logit_scale_init_value
        self.initializer_factor = 1.0
        self.extra = {k: v for k, v in kwargs.items() if k != "model_type"}

    @classmethod
    def from_json_file(cls, path):
        with open(path) as f:
            d = json.load(f)
        d.pop("model_type", None)
        return cls(**d)

    @classmethod
    def from_pretrained(cls, name_or_path, cache_dir=None, **kwargs):
        p = resolve_checkpoint_dir(name_or_path, cache_dir)
        return cls.from_json_file(os.path.join(p, "config.json"))


class LanguageBindImageConfig(_LanguageBindConfig):
    model_type = "LanguageBindImage"


class LanguageBindVideoConfig(_LanguageBindConfig):
    model_type = "LanguageBindVideo"


class LanguageBindDepthConfig(_LanguageBindConfig):
    model_type = "LanguageBindDepth"


class LanguageBindAudioConfig(_LanguageBindConfig):
    model_type = "LanguageBindAudio"


class LanguageBindThermalConfig(_LanguageBindConfig):
    model_type = "LanguageBindThermal"


def resolve_checkpoint_dir(name_or_path, cache_dir=None):
    """Local resolution of what the reference passes to `from_pretrained`
    (languagebind/__init__.py:63-64: 'LanguageBind/<name>', cache_dir): a directory holding
    config.json (+ weights).  There is no network access; nothing is downloaded."""
    cands = [name_or_path]
    if cache_dir:
        cands += [os.path.join(cache_dir, name_or_path),
                  os.path.join(cache_dir, name_or_path.replace("/", "--")),
                  os.path.join(cache_dir, "models--" + name_or_path.replace("/", "--"))]
    for c in cands:
        if os.path.isfile(os.path.join(c, "config.json")):
            return c
        snap = os.path.join(c, "snapshots")
        if os.path.isdir(snap):
            for s in sorted(os.listdir(snap)):
                if os.path.isfile(os.path.join(snap, s, "config.json")):
                    return os.path.join(snap, s)
    raise FileNotFoundError(
        f"no local checkpoint for {name_or_path!r} (searched {cands}); this environment has no "
        f"network. Set MISSM_SYNTHETIC=1 to build the towers from the synthetic ViT-L/14 configs.")


# ----------------------------------------------------------------------------------------------
# synthetic configs of SURVEY.md section 8(d) (hub config.json files are not available offline)
# ----------------------------------------------------------------------------------------------
VIT_L14 = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
               patch_size=14, image_size=224, lora_r=0, hidden_act="quick_gelu")
CLIP_TEXT = dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                 vocab_size=49408, max_position_embeddings=77, hidden_act="quick_gelu")
SYNTHETIC_PER_MODALITY = {
    'image': {}, 'depth': {}, 'thermal': {},
    'video': dict(add_time_attn=True, num_frames=8),
    'audio': dict(num_mel_bins=112, target_length=1036),
}
