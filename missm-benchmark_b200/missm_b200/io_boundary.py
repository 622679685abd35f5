"""CPU preprocessing boundary (OUT OF SCOPE of the hot path).

The reference's processors decode files with torchvision / torchaudio / decord and its tokenizers
need hub vocabulary files; neither is part of the path being accelerated.  These classes keep the
import surface of `languagebind` intact (`transform_dict[c](config)`, train_ddp.py:179;
`LanguageBindImageTokenizer.from_pretrained`, :170) and delegate to the caller's own
torchvision / transformers installation when it exists."""


class _Processor:
    modality = ""

    def __init__(self, config, tokenizer=None, **kwargs):
        self.config = config
        self.tokenizer = tokenizer

    def __call__(self, *a, **k):
        raise NotImplementedError(
            f"{type(self).__name__}: file decoding / CPU preprocessing is outside the B200 hot path; "
            f"feed tensors in the loader contract of SURVEY.md section 3.4, or use the reference's own "
            f"languagebind/{self.modality}/processing_{self.modality}.py for this step")


class LanguageBindImageProcessor(_Processor):
    modality = "image"


class LanguageBindVideoProcessor(_Processor):
    modality = "video"


class LanguageBindDepthProcessor(_Processor):
    modality = "depth"


class LanguageBindAudioProcessor(_Processor):
    modality = "audio"


class LanguageBindThermalProcessor(_Processor):
    modality = "thermal"


transform_dict = {
    'video': LanguageBindVideoProcessor, 'audio': LanguageBindAudioProcessor,
    'depth': LanguageBindDepthProcessor, 'thermal': LanguageBindThermalProcessor,
    'image': LanguageBindImageProcessor,
}


class _Tokenizer:
    """`CLIPTokenizer` subclass in the reference (tokenization_image.py:29-77, pad token =
    <|endoftext|>); resolved lazily from the caller's transformers install + local vocab files."""

    @classmethod
    def from_pretrained(cls, name_or_path, cache_dir=None, **kwargs):
        from transformers import CLIPTokenizer
        tok = CLIPTokenizer.from_pretrained(name_or_path, cache_dir=cache_dir, local_files_only=True, **kwargs)
        tok.pad_token = "<|endoftext|>"
        return tok


class LanguageBindImageTokenizer(_Tokenizer):
    pass


class LanguageBindVideoTokenizer(_Tokenizer):
    pass


class LanguageBindDepthTokenizer(_Tokenizer):
    pass


class LanguageBindAudioTokenizer(_Tokenizer):
    pass


class LanguageBindThermalTokenizer(_Tokenizer):
    pass
