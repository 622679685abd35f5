"""GPU parity of missm_image_preprocess (csrc/preprocess.cu) and the image / depth / thermal processors built on it
against the reference's own transforms (tests/golden/preproc.pt) and the torchvision oracle, both antialias
conventions; plus the processors end to end on files written to a temporary directory (PIL / OpenCV decode on the
host, one upload, one kernel), as data_loader.py:74-78 calls them."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import restatement as R  # noqa: E402
GOLD = os.path.join(HERE, "golden")
TOL = 1e-5


def run_kernel(pixels, S, antialias, **kw):
    from missm_b200 import ops
    out = torch.empty((3, S, S), device="cuda", dtype=torch.float32)
    ops.image_preprocess(torch.from_numpy(pixels).cuda(), out, S, R.OPENAI_DATASET_MEAN, R.OPENAI_DATASET_STD,
                         antialias=antialias, **kw)
    return out.cpu()


def test_kernel_matches_reference_golden():
    gold = torch.load(os.path.join(GOLD, "preproc.pt"), weights_only=False)
    meta = gold['meta']
    worst = 0.0
    for kind, seed, H, W in meta['cases']:
        if kind == 'depth':
            y = run_kernel(R.synth_depth(seed, H, W), 224, True, pre_div=1000.0, clip_lo=0.01,
                           clip_hi=meta['max_depth'], post_div=meta['max_depth'])
        else:
            y = run_kernel(R.synth_image(seed, H, W), 224, True)
        e = (y[:, ::meta['stride'], ::meta['stride']] - gold[f'{kind}/{seed}']).abs().max().item()
        worst = max(worst, e)
        assert e < TOL, (kind, seed, e)
    print('preprocess kernel vs the reference transforms: worst abs error', worst)


@pytest.mark.parametrize("antialias", [True, False])
@pytest.mark.parametrize("H,W,S", [(300, 400, 32), (517, 231, 32), (100, 180, 224), (64, 64, 32), (33, 75, 40),
                                   (225, 224, 224), (1080, 1920, 224), (3000, 4000, 224)])
def test_kernel_matches_torchvision_oracle(H, W, S, antialias):
    img = R.synth_image(H + W, H, W)
    e = (run_kernel(img, S, antialias) - R.image_transform(img, size=S, antialias=antialias)).abs().max().item()
    assert e < TOL, e
    d = R.synth_depth(H, H, W)
    y = run_kernel(d, S, antialias, pre_div=1000.0, clip_lo=0.01, clip_hi=10.0, post_div=10.0)
    e = (y - R.depth_transform(d, 10.0, size=S, antialias=antialias)).abs().max().item()
    assert e < TOL, e


def test_processors_end_to_end_from_files(tmp_path):
    import cv2
    from PIL import Image
    import languagebind as lb
    from missm_b200 import config as C
    cfg = C.LanguageBindDepthConfig(text_config={}, vision_config={}, projection_dim=64)
    imgs = [R.synth_image(21, 240, 320), R.synth_image(22, 400, 250)]
    paths = []
    for i, im in enumerate(imgs):
        paths.append(str(tmp_path / f"im{i}.png"))
        Image.fromarray(im).save(paths[-1])
    for m in ('image', 'thermal'):
        out = lb.transform_dict[m](cfg)(paths)['pixel_values']
        assert out.is_cuda and tuple(out.shape) == (2, 3, 224, 224)
        for i, im in enumerate(imgs):
            assert (out[i].cpu() - R.image_transform(im)).abs().max().item() < TOL
    one = lb.transform_dict['image'](cfg)(paths[0])['pixel_values']       # a single path, as the loader passes it
    assert tuple(one.shape) == (1, 3, 224, 224)
    d16 = R.synth_depth(23, 200, 300).astype(np.uint16)
    dp = str(tmp_path / "depth.png")
    cv2.imwrite(dp, d16)
    out = lb.transform_dict['depth'](cfg)(dp)['pixel_values']
    assert (out[0].cpu() - R.depth_transform(d16.astype(np.float32), 10.0)).abs().max().item() < TOL
    with pytest.raises(ValueError):
        lb.transform_dict['image'](cfg).transform(np.zeros((8, 8), np.uint8))


def test_preprocess_throughput():
    """Not a pass/fail bar: prints what one 1080p -> 224 launch costs next to its algorithmic bytes."""
    from missm_b200 import ops
    img = torch.from_numpy(R.synth_image(1, 1080, 1920)).cuda()
    out = torch.empty((3, 224, 224), device="cuda")
    for aa in (True, False):
        for _ in range(3):
            ops.image_preprocess(img, out, 224, R.OPENAI_DATASET_MEAN, R.OPENAI_DATASET_STD, antialias=aa)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.image_preprocess(img, out, 224, R.OPENAI_DATASET_MEAN, R.OPENAI_DATASET_STD, antialias=aa)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"preprocess 1080x1920 -> 224, antialias={aa}: {us:.1f} us / image "
              f"({(img.numel() + out.numel() * 4) / us / 1e3:.1f} GB/s of 6.8 MB algorithmic)")
