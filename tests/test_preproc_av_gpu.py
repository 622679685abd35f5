"""GPU input pipeline for video and audio against golden vectors of the UNMODIFIED
reference processors (tests/golden/av_preproc.pt, oracle/make_golden_av.py) and, for audio, against torchaudio's own
kaldi fbank on fresh waveforms.  Tolerances: video <= 2e-5 absolute (fp32 bilinear blend of values in [-2, 2.2]);
audio <= 5e-4 absolute on the normalised log-mel values (a 400-term fp32 DFT vs torch's FFT, then log)."""
import os
import sys
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "av_preproc.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def _vcfg(**kw):
    return types.SimpleNamespace(vision_config=types.SimpleNamespace(**kw))


@pytest.mark.parametrize("name", ["landscape", "portrait", "small"])
def test_video_transform_matches_reference_chain(gold, name):
    from missm_b200.io_boundary import LanguageBindVideoProcessor
    proc = LanguageBindVideoProcessor(_vcfg(video_decode_backend='opencv', num_frames=2))
    frames = gold[f"video/{name}/frames"]
    import random
    random.seed(gold[f"video/{name}/seed"])                       # the processor draws the flip as torchvision does
    out = proc.one(frames)
    assert out.is_cuda and tuple(out.shape) == (3, 2, 224, 224)
    want = gold[f"video/{name}/out_every2nd_pixel"]
    err = (out.cpu()[:, :, ::2, ::2] - want).abs().max().item()
    print(name, "flipped", gold[f"video/{name}/flipped"], "max abs err", err)
    assert err < 2e-5
    # the other flip state is the mirror image
    a = proc.transform(frames, hflip=False).cpu()
    b = proc.transform(frames, hflip=True).cpu()
    assert torch.equal(a.flip(-1), b)


@pytest.mark.parametrize("name", ["short", "long", "stereo"])
def test_audio_transform_matches_reference(gold, name):
    import numpy as np
    from missm_b200.io_boundary import LanguageBindAudioProcessor
    proc = LanguageBindAudioProcessor(_vcfg(audio_sample_rate=16000, num_mel_bins=112, target_length=1036,
                                            audio_mean=-4.2677393, audio_std=4.5689974))
    wave = gold[f"audio/{name}/wave"]
    np.random.seed(1234)                                           # the three chunk offsets, drawn as the reference does
    out = proc.one((wave.clone(), 16000))
    assert out.is_cuda and tuple(out.shape) == (3, 112, 1036)
    want = gold[f"audio/{name}/out_every4th_frame"]
    err = (out.cpu()[:, :, ::4] - want).abs().max().item()
    print(name, "max abs err", err)
    assert err < 5e-4


def test_audio_fbank_vs_torchaudio_fresh_waveforms():
    torchaudio = pytest.importorskip("torchaudio")
    from missm_b200 import io_boundary as io, ops
    g = torch.Generator().manual_seed(3)
    w = io.kaldi_mel_banks(112).cuda()
    for n in (400, 16000, 16000 * 11 + 5):
        wave = (torch.randn(1, n, generator=g) * 0.3).float()
        ref = torchaudio.compliance.kaldi.fbank(wave - wave.mean(), htk_compat=True, sample_frequency=16000, use_energy=False,
                                                window_type="hanning", num_mel_bins=112, dither=0.0, frame_length=25,
                                                frame_shift=10)
        ref64 = torchaudio.compliance.kaldi.fbank((wave - wave.mean()).double(), htk_compat=True, sample_frequency=16000,
                                                  use_energy=False, window_type="hanning", num_mel_bins=112, dither=0.0,
                                                  frame_length=25, frame_shift=10)
        T = ref.shape[0]
        out, nf = ops.audio_fbank(wave.cuda(), w, T, (0, 0, 0), 0.0, 0.5)      # mean 0, 2 * std = 1: the raw log-mel
        assert nf == T
        mine = out[0].t().cpu()
        err, err64, ta64 = (mine - ref).abs().max().item(), (mine.double() - ref64).abs().max().item(), \
            (ref.double() - ref64).abs().max().item()
        print(f"n {n} frames {T}: vs torchaudio fp32 {err:.2e}; vs its float64 run: this kernel {err64:.2e}, torchaudio fp32 {ta64:.2e}")
        # raw log-mel values: two fp32 spectra (400-term DFT here, torch's FFT there) differ by their own rounding;
        # the kernel must be as close to the float64 result as torchaudio's fp32 run is (x 3), and within 3e-3 of it
        assert err < 3e-3 and err64 < max(5e-4, 3 * ta64)
    with pytest.raises(ValueError):
        ops.audio_fbank(torch.zeros(1, 100).cuda(), w, 10, (0, 0, 0), 0.0, 0.5)


def test_processors_fail_loudly_on_bad_input():
    from missm_b200.io_boundary import LanguageBindVideoProcessor
    proc = LanguageBindVideoProcessor(_vcfg(video_decode_backend='opencv', num_frames=2))
    with pytest.raises(ValueError):
        proc.one(torch.zeros(2, 3, 32, 32))                       # not [T, H, W, 3] uint8
    with pytest.raises(ValueError):
        proc()
