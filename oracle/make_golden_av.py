"""TEST INFRASTRUCTURE ONLY -- golden vectors for the video / audio input pipeline, produced by the UNMODIFIED
reference processors (/root/reference/languagebind/{video,audio}/processing_*.py through oracle/ref_shim.py) in the
build container:  python oracle/make_golden_av.py  ->  tests/golden/av_preproc.pt

Audio: the reference's own AudioTransform (torchaudio.compliance.kaldi.fbank is the real third-party code, present in
this image) on seeded waveforms -- shorter than target_length (repeat branch), longer (three random chunks; numpy's
RNG seeded, the product draws the same numbers in the same order) and a 2-channel one.
Video: the reference's `decord` / `opencv` transform chain needs pytorchvideo (absent, unpinned); ShortSideScale is
restated here from its published algorithm (bilinear F.interpolate to the floor-scaled size) and composed with
torchvision's own NormalizeVideo / CenterCropVideo exactly as get_video_transform does -- parity unpinned for that one
third-party function, pinned for the chain around it."""
import math
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def short_side_scale(x, size):
    c, t, h, w = x.shape
    if w < h:
        nh, nw = int(math.floor((float(h) / w) * size)), size
    else:
        nh, nw = size, int(math.floor((float(w) / h) * size))
    return torch.nn.functional.interpolate(x, size=(nh, nw), mode="bilinear", align_corners=False)


def main():
    ref_shim.install()
    import pytorchvideo.transforms as pvt
    pvt.ShortSideScale = lambda size: (lambda x: short_side_scale(x, size))       # the one restated third-party piece
    pvt.ApplyTransformToKey = lambda key, transform: transform
    pvt.UniformTemporalSubsample = lambda n: (lambda x: x)
    from languagebind.audio import processing_audio as pa
    from languagebind.video import processing_video as pv
    out = {}
    # ---- audio
    acfg = types.SimpleNamespace(vision_config=types.SimpleNamespace(audio_sample_rate=16000, num_mel_bins=112, target_length=1036,
                                                                   audio_mean=-4.2677393, audio_std=4.5689974))
    tr = pa.get_audio_transform(acfg)
    g = torch.Generator().manual_seed(7)
    for name, ch, n in (("short", 1, 16000 * 3 + 123), ("long", 1, 16000 * 14 + 77), ("stereo", 2, 16000 * 5)):
        wave = (torch.randn(ch, n, generator=g) * 0.1 + 0.01 * torch.sin(torch.arange(n) * 0.05)).float()
        np.random.seed(1234)
        res = tr((wave.clone(), 16000))
        out[f"audio/{name}/wave"] = wave
        out[f"audio/{name}/out_every4th_frame"] = res[:, :, ::4].clone()       # (fixture size: every 4th time step)
    # ---- video
    vcfg = types.SimpleNamespace(vision_config=types.SimpleNamespace(video_decode_backend='opencv', num_frames=2))
    chain = pv.get_video_transform(vcfg)
    for name, (H, W), seed in (("landscape", (240, 426), 1), ("portrait", (400, 250), 2), ("small", (120, 160), 5)):
        frames = torch.randint(0, 256, (2, H, W, 3), generator=g, dtype=torch.uint8)
        clip = frames.permute(3, 0, 1, 2)                                   # (T, H, W, C) -> (C, T, H, W), :100
        random.seed(seed)
        flipped = random.random() < 0.5                                     # what RandomHorizontalFlipVideo will draw
        random.seed(seed)
        res = chain(clip)
        out[f"video/{name}/frames"] = frames
        out[f"video/{name}/out_every2nd_pixel"] = res[:, :, ::2, ::2].clone()   # (fixture size)
        out[f"video/{name}/flipped"] = flipped
        out[f"video/{name}/seed"] = seed
    torch.save(out, os.path.join(GOLD, "av_preproc.pt"))
    print({k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
