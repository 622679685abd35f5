"""CPU tests of the input-pipeline row (reference: languagebind/image/
processing_image.py:20-35, thermal/processing_thermal.py:15-31, depth/processing_depth.py:17-62): the oracle against
the golden produced by the reference's OWN transforms, and the resampling rule the C ABI documents
(include/missm_b200.h: missm_image_preprocess; stand-in in tests/ops_emulation.py) against torchvision in both
antialias conventions, down- and up-scaling, portrait / landscape / square, odd crop offsets.  The CUDA kernel is
checked against the same references on the B200 by tests/test_preproc_gpu.py."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
sys.path.insert(0, HERE)
import restatement as R  # noqa: E402
import ops_emulation as E  # noqa: E402

GOLD = os.path.join(HERE, "golden")
TOL = 1e-5          # absolute, on O(1) normalised pixels (fp32 summation order differs)


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "preproc.pt"), weights_only=False)


def oracle_case(kind, seed, H, W, max_depth, antialias=True):
    if kind == 'depth':
        return R.depth_transform(R.synth_depth(seed, H, W), max_depth, antialias=antialias)
    return R.image_transform(R.synth_image(seed, H, W), antialias=antialias)


def test_oracle_matches_reference_transforms(gold):
    meta = gold['meta']
    for kind, seed, H, W in meta['cases']:
        y = oracle_case(kind, seed, H, W, meta['max_depth'])
        assert torch.equal(y[:, ::meta['stride'], ::meta['stride']], gold[f'{kind}/{seed}']), (kind, seed)


@pytest.mark.parametrize("antialias", [True, False])
@pytest.mark.parametrize("H,W,S", [(300, 400, 32), (517, 231, 32), (100, 180, 224), (64, 64, 32), (33, 75, 40),
                                   (225, 224, 224), (480, 640, 224), (3000, 4000, 224)])
def test_abi_resampling_rule_matches_torchvision(H, W, S, antialias):
    img = R.synth_image(H + W, H, W)
    ref = R.image_transform(img, size=S, antialias=antialias)
    out = torch.empty((3, S, S))
    E.image_preprocess(torch.from_numpy(img), out, S, R.OPENAI_DATASET_MEAN, R.OPENAI_DATASET_STD, antialias=antialias)
    assert (out - ref).abs().max().item() < TOL
    d = R.synth_depth(H, H, W)
    ref = R.depth_transform(d, 10.0, size=S, antialias=antialias)
    E.image_preprocess(torch.from_numpy(d), out, S, R.OPENAI_DATASET_MEAN, R.OPENAI_DATASET_STD, antialias=antialias,
                       pre_div=1000.0, clip_lo=0.01, clip_hi=10.0, post_div=10.0)
    assert (out - ref).abs().max().item() < TOL


def test_processors_keep_the_reference_surface_and_have_no_cpu_path():
    import languagebind as lb
    from missm_b200 import config as C
    cfg = C.LanguageBindDepthConfig(text_config={}, vision_config={}, projection_dim=64)
    assert cfg.vision_config.max_depth == 10
    for m in ('image', 'depth', 'thermal'):
        proc = lb.transform_dict[m](cfg)
        assert hasattr(proc, 'batch_decode') and hasattr(proc, 'decode') and proc.config is cfg
        with pytest.raises(ValueError):
            proc()                                                     # processing_image.py:47-48
        if not torch.cuda.is_available():
            with pytest.raises(RuntimeError, match="no CPU fallback"):
                proc.transform(np.zeros((8, 8, 3), np.uint8))
    # video / audio: host decode with the reference's own decoders, everything after it on the device -- no CPU path
    vproc, aproc = lb.transform_dict['video'](cfg), lb.transform_dict['audio'](cfg)
    for proc in (vproc, aproc):
        assert hasattr(proc, 'batch_decode') and hasattr(proc, 'decode') and proc.config is cfg
        with pytest.raises(ValueError):
            proc()
    with pytest.raises(ValueError):
        vproc.decode_frames(torch.zeros(2, 3, 8, 8))                   # not uint8 [T, H, W, 3]
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            vproc.transform(torch.zeros(2, 8, 8, 3, dtype=torch.uint8))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            aproc.transform((torch.zeros(1, 16000), 16000))


def test_kaldi_mel_banks_equal_torchaudio():
    """The mel filterbank handed to the fbank kernel is torchaudio's own (get_mel_banks + the zero Nyquist column)."""
    torchaudio = pytest.importorskip("torchaudio")
    from missm_b200.io_boundary import kaldi_mel_banks
    for nb in (112, 128, 23):
        ref, _ = torchaudio.compliance.kaldi.get_mel_banks(nb, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
        assert torch.equal(kaldi_mel_banks(nb), torch.nn.functional.pad(ref, (0, 1)))
