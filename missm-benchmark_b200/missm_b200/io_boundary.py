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

Video and audio (second half of the same row): decoding stays on the host with the reference's own decoders (OpenCV /
decord frame grab at np.linspace(0, n - 1, num_frames) indices, video/processing_video.py:84-109; torchaudio.load,
audio/processing_audio.py:21-22), everything after the decode runs on the device: the clip's frames go up as uint8
and ONE launch of `missm_video_preprocess` does x / 255 -> Normalize -> ShortSideScale(224) -> CenterCrop(224) -> flip;
the waveform goes up once and `missm_audio_fbank` does kaldi fbank + chunk / repeat + normalise (csrc/preprocess_av.cu).
pytorchvideo's ShortSideScale (absent here, unpinned) is restated from its published algorithm; torchaudio's kaldi
fbank is pinned against torchaudio itself in the tests.
"""
import math
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


class _ClipProcessor(_Processor):
    """Shared __call__ of the reference's video / audio processors (processing_video.py:124-147,
    processing_audio.py:137-160): a list of paths -> stacked `pixel_values`."""

    def one(self, item):
        raise NotImplementedError

    def __call__(self, images=None, text=None, context_length=77, return_tensors=None, **kwargs):
        if text is None and images is None:
            raise ValueError("You have to specify either text or images. Both cannot be none.")
        encoding = None
        if text is not None:
            encoding = self.tokenizer(text, max_length=context_length, padding='max_length', truncation=True,
                                      return_tensors=return_tensors, **kwargs)
        feats = None
        if images is not None:
            feats = torch.stack([self.one(x) for x in make_list_of_images(images)])
        if text is not None and images is not None:
            encoding["pixel_values"] = feats
            return encoding
        return encoding if text is not None else {"pixel_values": feats}

    def batch_decode(self, skip_special_tokens=True, *args, **kwargs):
        return self.tokenizer.batch_decode(*args, skip_special_tokens=skip_special_tokens, **kwargs)

    def decode(self, skip_special_tokens=True, *args, **kwargs):
        return self.tokenizer.decode(*args, skip_special_tokens=skip_special_tokens, **kwargs)


def _need_cuda(who):
    if not torch.cuda.is_available():
        raise RuntimeError(f"missm_b200: {who} preprocesses on the GPU and found no CUDA device (there is no CPU fallback)")


class LanguageBindVideoProcessor(_ClipProcessor):
    modality = "video"
    size = 224

    def decode_frames(self, path):
        """-> uint8 [T, H, W, 3] RGB frames at np.linspace(0, n - 1, num_frames) (processing_video.py:84-109)."""
        import numpy as np
        vc = self.config.vision_config
        if torch.is_tensor(path) or isinstance(path, np.ndarray):       # already decoded frames
            t = path if torch.is_tensor(path) else torch.from_numpy(path)
            if t.dim() != 4 or t.shape[3] != 3 or t.dtype != torch.uint8:
                raise ValueError(f"decoded frames must be uint8 [T, H, W, 3], got {t.dtype} {tuple(t.shape)}")
            return t
        backend = vc.video_decode_backend
        if backend == 'opencv':
            import cv2
            vr = cv2.VideoCapture(path)
            n = int(vr.get(cv2.CAP_PROP_FRAME_COUNT))
            frames = []
            for i in np.linspace(0, n - 1, vc.num_frames, dtype=int):
                vr.set(1, int(i))
                ok, frame = vr.read()
                if not ok:
                    raise IOError(f"cannot read frame {i} of {path}")
                frames.append(torch.from_numpy(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)))
            vr.release()
            return torch.stack(frames)
        if backend == 'decord':
            try:
                import decord
            except ImportError as e:
                raise ImportError("video_decode_backend='decord' needs the decord package; set it to 'opencv' "
                                  "(processing_video.py:96-109 reads the same frame indices)") from e
            decord.bridge.set_bridge('torch')
            vr = decord.VideoReader(path, ctx=decord.cpu(0))
            return vr.get_batch(np.linspace(0, len(vr) - 1, vc.num_frames, dtype=int))
        raise NotImplementedError(f"video_decode_backend={backend!r}: 'opencv' and 'decord' are built (the pytorchvideo "
                                  f"branch re-samples in time inside a third-party transform)")

    def transform(self, frames, hflip=None):
        from . import ops
        _need_cuda(type(self).__name__)
        if hflip is None:                     # RandomHorizontalFlipVideo(p = 0.5) draws from Python's `random` module
            import random
            hflip = random.random() < 0.5
        f = frames.contiguous().cuda(non_blocking=True)
        out = torch.empty((3, f.shape[0], self.size, self.size), device=f.device, dtype=torch.float32)
        return ops.video_preprocess(f, out, self.size, OPENAI_DATASET_MEAN, OPENAI_DATASET_STD, hflip)

    def one(self, item):
        return self.transform(self.decode_frames(item))


def kaldi_mel_banks(num_bins, sample_freq=16000.0, padded=512, low_freq=20.0, high_freq=0.0):
    """The triangular mel filters of torchaudio.compliance.kaldi.get_mel_banks (no VTLN) + the zero Nyquist column fbank
    appends: float32 [num_bins, padded / 2 + 1].  Parameter-space work, done once per processor."""
    nyq = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyq
    mel = lambda f: 1127.0 * math.log(1.0 + f / 700.0)      # noqa: E731
    nfft = padded // 2
    lo, hi = mel(low_freq), mel(high_freq)
    delta = (hi - lo) / (num_bins + 1)
    b = torch.arange(num_bins, dtype=torch.float32).unsqueeze(1)
    left, center, right = lo + b * delta, lo + (b + 1.0) * delta, lo + (b + 2.0) * delta
    m = 1127.0 * (1.0 + (sample_freq / padded) * torch.arange(nfft, dtype=torch.float32) / 700.0).log().unsqueeze(0)
    w = torch.max(torch.zeros(1), torch.min((m - left) / (center - left), (right - m) / (right - center)))
    return torch.nn.functional.pad(w, (0, 1)).contiguous()


class LanguageBindAudioProcessor(_ClipProcessor):
    modality = "audio"

    def load(self, path):
        """-> (float32 [channels, n], sample rate), as torchaudio.load (processing_audio.py:21-22)."""
        if isinstance(path, tuple):
            return path
        import torchaudio
        return torchaudio.load(path)

    def transform(self, wave_and_sr, offsets=None):
        """AudioTransform.__call__ (processing_audio.py:45-94) -> CUDA float32 [3, num_mel_bins, target_length]."""
        from . import ops
        _need_cuda(type(self).__name__)
        vc = self.config.vision_config
        wave, sr = wave_and_sr
        if sr != vc.audio_sample_rate:
            import torchaudio                                          # :49-51 (sinc resampling: third party, kept as is)
            wave = torchaudio.functional.resample(wave, orig_freq=sr, new_freq=vc.audio_sample_rate)
        if vc.audio_sample_rate != 16000:
            raise NotImplementedError("the fbank kernel is built for 16 kHz (400 / 160 / 512-sample frames), the "
                                      "reference's audio_sample_rate default (configuration_audio.py:206)")
        wave = wave.to(torch.float32).contiguous().cuda(non_blocking=True)
        key = (int(vc.num_mel_bins), str(wave.device))
        cache = self.__dict__.setdefault('_mel', {})
        if key not in cache:
            cache[key] = kaldi_mel_banks(int(vc.num_mel_bins), float(vc.audio_sample_rate)).to(wave.device)
        target = int(vc.target_length)
        nf = int(ops.lib().missm_fbank_num_frames(wave.shape[1]))
        if offsets is None:
            offsets = (0, 0, 0)
            if nf > target:                   # three random chunks: front / middle / back thirds (:57-77), numpy's RNG
                import numpy as np
                ranges = np.array_split(list(range(0, nf - target + 1)), 3)
                if len(ranges[1]) == 0:
                    ranges[1] = [0]
                if len(ranges[2]) == 0:
                    ranges[2] = [0]
                offsets = tuple(int(np.random.choice(r)) for r in ranges)
        out, _ = ops.audio_fbank(wave, cache[key], target, offsets, float(vc.audio_mean), float(vc.audio_std))
        return out

    def one(self, item):
        return self.transform(self.load(item))


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
