"""Input boundary: the processors and tokenizers `train_ddp.py:170,179` / `data_loader.py:74-78` use.

Image, depth and thermal: the file is decoded on the host exactly as the
reference does (PIL / OpenCV -- decoding is not part of the path), the decoded pixels are uploaded once as uint8
(float for depth) and ONE CUDA kernel (`missm_image_preprocess`, csrc/preprocess.cu) does what the reference's
torchvision chain does on the host per sample: ToTensor -> Resize(224, BICUBIC) -> CenterCrop(224) -> Normalize
(languagebind/image/processing_image.py:20-29, thermal/processing_thermal.py:15-25; depth adds DepthNorm in front,
depth/processing_depth.py:21-57).  The unchanged loader has 0 workers (data_loader.py:312), so the processor runs
in the training process and returns CUDA tensors; `default_collate` stacks them on the device and `to_device`
(train_ddp.py:224-229) becomes a no-op.  There is no CPU path: without a CUDA device the call raises.

Resize semantics: torchvision's `Resize` on a tensor changed its default from antialias = False (<= 0.16, the
reference's era) to True (>= 0.17, what the reference computes when run in this image).  The default here follows
the reference as run in this image (True); MISSM_RESIZE_ANTIALIAS=0 selects the older behaviour.  Both are pinned
against torchvision in the tests.

Video (decord + pytorchvideo) and audio (torchaudio kaldi fbank) processors are not built: they keep the import
surface and raise with a pointer to the reference's own file.
"""
import os

import torch

OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)      # processing_image.py:10-11
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)


def make_list_of_images(x):
    return x if isinstance(x, list) else [x]


def _antialias_default():
    return os.environ.get("MISSM_RESIZE_ANTIALIAS", "1") != "0"


class _Processor:
    modality = ""

    def __init__(self, config, tokenizer=None, **kwargs):
        self.config = config
        self.tokenizer = tokenizer

    def __call__(self, *a, **k):
        raise NotImplementedError(
            f"{type(self).__name__}: decoding + preprocessing of this modality is not built (needs decord / "
            f"pytorchvideo / torchaudio); feed tensors in the loader contract of SURVEY.md section 3.4, or use the "
            f"reference's own languagebind/{self.modality}/processing_{self.modality}.py for this step")


class _ImageLikeProcessor(_Processor):
    """Shared __call__ of the reference's image / depth / thermal processors (processing_image.py:46-66)."""
    size = 224

    def load(self, path):
        raise NotImplementedError

    def preprocess_params(self):
        return dict(pre_div=255.0)

    def transform(self, pixels, out=None, antialias=None):
        """Decoded pixels (uint8 [H, W, 3] or float32 [H, W]; numpy or torch, host or device) -> CUDA fp32
        [3, 224, 224].  Fails loudly without a CUDA device."""
        from . import ops
        if not torch.cuda.is_available():
            raise RuntimeError(f"missm_b200: {type(self).__name__} preprocesses on the GPU and found no CUDA device "
                               f"(there is no CPU fallback)")
        t = pixels if torch.is_tensor(pixels) else torch.from_numpy(pixels)
        if t.dtype == torch.uint8:
            if t.dim() != 3 or t.shape[2] != 3:
                raise ValueError(f"expected an RGB image [H, W, 3], got {tuple(t.shape)} (the reference's Normalize "
                                 f"with three means fails on other channel counts too)")
        elif t.dim() != 2:
            raise ValueError(f"expected a single-channel float image [H, W], got {tuple(t.shape)}")
        t = t.contiguous().cuda(non_blocking=True)
        if out is None:
            out = torch.empty((3, self.size, self.size), device=t.device, dtype=torch.float32)
        aa = _antialias_default() if antialias is None else antialias
        return ops.image_preprocess(t, out, self.size, OPENAI_DATASET_MEAN, OPENAI_DATASET_STD, antialias=aa,
                                    **self.preprocess_params())

    def __call__(self, images=None, text=None, context_length=77, return_tensors=None, **kwargs):
        if text is None and images is None:
            raise ValueError("You have to specify either text or images. Both cannot be none.")
        encoding = None
        if text is not None:
            encoding = self.tokenizer(text, max_length=context_length, padding='max_length', truncation=True,
                                      return_tensors=return_tensors, **kwargs)
        if images is not None:
            images = make_list_of_images(images)
            batch = torch.empty((len(images), 3, self.size, self.size), device="cuda", dtype=torch.float32) \
                if torch.cuda.is_available() else None
            if batch is None:
                self.transform(None)                                  # raises: no CUDA device
            for i, image in enumerate(images):
                self.transform(self.load(image), out=batch[i])
        if text is not None and images is not None:
            encoding["pixel_values"] = batch
            return encoding
        if text is not None:
            return encoding
        return {"pixel_values": batch}

    def batch_decode(self, skip_special_tokens=True, *args, **kwargs):
        return self.tokenizer.batch_decode(*args, skip_special_tokens=skip_special_tokens, **kwargs)

    def decode(self, skip_special_tokens=True, *args, **kwargs):
        return self.tokenizer.decode(*args, skip_special_tokens=skip_special_tokens, **kwargs)


def _pil_rgb_array(path):
    """`Image.open(path)` + what ToTensor sees of it (processing_image.py:32-35): the image's own mode."""
    import numpy as np
    from PIL import Image, ImageFile
    ImageFile.LOAD_TRUNCATED_IMAGES = True                            # processing_image.py:7-8
    img = path if isinstance(path, Image.Image) else Image.open(path)
    return np.array(img)                # a writable copy: torch.from_numpy refuses to share PIL's read-only buffer quietly


class LanguageBindImageProcessor(_ImageLikeProcessor):
    modality = "image"

    def load(self, path):
        return _pil_rgb_array(path)


class LanguageBindThermalProcessor(_ImageLikeProcessor):
    modality = "thermal"

    def load(self, path):
        return _pil_rgb_array(path)


class LanguageBindDepthProcessor(_ImageLikeProcessor):
    modality = "depth"

    def load(self, path):
        import numpy as np
        if isinstance(path, np.ndarray):
            return path.astype('float32')
        import cv2                                                    # processing_depth.py:17-18
        return cv2.imread(path, cv2.IMREAD_UNCHANGED).astype('float32')

    def preprocess_params(self):
        # DepthNorm (processing_depth.py:21-41): / 1000 (mm -> m), clip to [0.01, max_depth], / max_depth
        max_depth = float(self.config.vision_config.max_depth)
        if max_depth == 0:
            raise NotImplementedError("max_depth = 0 (normalise by the image's own maximum) is not built; the "
                                      "reference config default is 10 (depth/configuration_depth.py:205)")
        return dict(pre_div=1000.0, clip_lo=0.01, clip_hi=max_depth, post_div=max_depth)


class LanguageBindVideoProcessor(_Processor):
    modality = "video"


class LanguageBindAudioProcessor(_Processor):
    modality = "audio"


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
