"""CPU: the augmentation oracle (oracle/augment_oracle.py) against the libraries the reference calls -- Pillow's bicubic
resize and torchvision's transforms (data/datasets/make_dataloader.py:245-253) -- on seeded images with given draws, and
the product's coefficient tables (editor_b200/data.py) against the oracle's."""
import numpy as np
import pytest
import torch
from PIL import Image

from oracle import augment_oracle as ao

CASES = [(300, 140, 256, 128), (128, 256, 128, 256), (100, 50, 256, 128), (511, 257, 256, 128), (90, 300, 128, 256),
         (256, 128, 256, 128), (640, 250, 256, 128), (129, 255, 128, 256)]


@pytest.mark.parametrize("hs,ws,h,w", CASES)
def test_resize_equals_pillow(hs, ws, h, w):
    rng = np.random.default_rng(hs * 1000 + ws)
    img = rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((w, h), Image.BICUBIC))
    assert np.array_equal(ao.resize_bicubic_u8(img, h, w), ref)            # byte work: bit-exact


@pytest.mark.parametrize("hs,ws,h,w", CASES[:5])
def test_pipeline_equals_torchvision(hs, ws, h, w):
    import torchvision.transforms as T
    import torchvision.transforms.functional as F
    rng = np.random.default_rng(7 + hs)
    for trial in range(6):
        img = rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
        flip, top, left, erase = ao.sample_params(rng, h, w, 10, 0.5, 0.8)
        noise = None
        pil = T.Resize([h, w], interpolation=3)(Image.fromarray(img))
        if flip:
            pil = F.hflip(pil)
        pil = F.crop(T.Pad(10)(pil), top, left, h, w)
        ref = T.Normalize(mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5])(T.ToTensor()(pil))
        if erase is not None:
            et, el, eh, ew = erase
            noise = torch.randn(3, eh, ew, generator=torch.Generator().manual_seed(trial))
            ref[:, et:et + eh, el:el + ew] = noise
            noise = noise.numpy()
        got = ao.augment(img, h, w, flip, top, left, 10, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5), erase, noise)
        assert np.array_equal(got, ref.numpy())                             # same IEEE operations: bit-exact


def test_product_tables_equal_oracle_tables():
    from editor_b200 import data
    for a, b in ((300, 256), (140, 128), (100, 256), (511, 256), (257, 128), (90, 128), (640, 256), (255, 256), (128, 128)):
        ob, ok = ao.resample_coeffs(a, b)
        pb, pk = data.resample_tables(a, b)
        assert np.array_equal(ob, pb) and np.array_equal(ok, pk)


def test_sampled_draws_are_valid():
    from editor_b200 import data
    from editor_b200.config import cfg
    aug = data.GpuAugment.__new__(data.GpuAugment)
    aug.H, aug.W, aug.pad, aug.flip_p, aug.erase_p = 256, 128, 10, 0.5, 0.5
    aug.min_area, aug.max_area, aug.min_aspect = 0.02, 1 / 3, 0.3
    aug.rng = np.random.default_rng(3)
    p = aug.sample(4000)
    assert 0.45 < p[:, 0].mean() < 0.55 and p[:, 1].min() == 0 and p[:, 1].max() == 20 and p[:, 2].max() == 20
    er = p[p[:, 5] > 0]
    assert 0.42 < len(er) / 4000 < 0.55                                   # RE_PROB 0.5, a few draws fail all 10 attempts
    assert (er[:, 3] + er[:, 5] <= 256).all() and (er[:, 4] + er[:, 6] <= 128).all() and (er[:, 5] < 256).all() and (er[:, 6] < 128).all()
    area = er[:, 5] * er[:, 6] / (256 * 128)
    assert area.min() > 0.015 and area.max() < 0.35
    assert cfg.INPUT.PADDING == 10 and cfg.INPUT.RE_PROB == 0.5
