"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's training-time input pipeline:

    T.Resize(SIZE_TRAIN, interpolation=3) -> T.RandomHorizontalFlip(p) -> T.Pad(PADDING) -> T.RandomCrop(SIZE_TRAIN) ->
    T.ToTensor() -> T.Normalize(mean, std) -> RandomErasing(probability, mode='pixel', max_count=1)
    (data/datasets/make_dataloader.py:245-253; RandomErasing :55-140), applied to each of the three modality images of a
    sample independently (data/datasets/bases.py:100-103).

The arithmetic of the first step lives in a third-party dependency that is not under /root/reference: Pillow
(`Image.resize(..., BICUBIC)`, reached through torchvision's T.Resize on PIL images; requirements.txt:158 pins
torchvision==0.14.1 and leaves Pillow unpinned; this image has Pillow 12.2 / torchvision 0.26).  Its published algorithm (libImaging/Resample.c) is restated here: separable convolution,
support 2 * max(scale, 1) (i.e. antialiased when shrinking), Keys a = -0.5 kernel, coefficients normalised in double,
rounded to 22-bit fixed point, horizontal pass then vertical pass, each rounded and clipped to uint8.  Pinned by
tests/test_augment_oracle.py against Pillow itself and against torchvision's transforms on seeded images.

Random draws are INPUTS here (flip flag, crop offsets, erase rectangle, erase noise): the reference draws them from
Python's `random` / torch's CPU generator per image; parity is defined for given draws.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc: (bounds [out,2] int32 = (xmin, count), kk [out,ksize] int32)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic_u8(img, out_h, out_w):
    """img uint8 [H, W, C] -> uint8 [out_h, out_w, C]; identical to PIL.Image.resize((out_w, out_h), BICUBIC)."""
    h, w, _ = img.shape
    cur = img
    if w != out_w:
        bounds, kk = resample_coeffs(w, out_w)
        out = np.empty((h, out_w, img.shape[2]), np.uint8)
        for xx in range(out_w):
            x0, n = bounds[xx]
            acc = (cur[:, x0:x0 + n, :].astype(np.int64) * kk[xx, :n].astype(np.int64)[None, :, None]).sum(1) + (1 << (PRECISION_BITS - 1))
            out[:, xx, :] = _clip8(acc)
        cur = out
    if h != out_h:
        bounds, kk = resample_coeffs(h, out_h)
        out = np.empty((out_h, cur.shape[1], img.shape[2]), np.uint8)
        for yy in range(out_h):
            y0, n = bounds[yy]
            acc = (cur[y0:y0 + n].astype(np.int64) * kk[yy, :n].astype(np.int64)[:, None, None]).sum(0) + (1 << (PRECISION_BITS - 1))
            out[yy] = _clip8(acc)
        cur = out
    return cur


def augment(img, out_h, out_w, flip, top, left, pad, mean, std, erase=None, noise=None):
    """One modality image, uint8 [H, W, 3] -> float32 [3, out_h, out_w].
    flip: bool; (top, left): RandomCrop offsets in the padded image, 0 .. 2*pad; erase: None or (top, left, h, w);
    noise: float32 [3, h, w] normal draws written into the erase rectangle (RandomErasing mode 'pixel')."""
    r = resize_bicubic_u8(img, out_h, out_w)
    if flip:
        r = r[:, ::-1]
    padded = np.zeros((out_h + 2 * pad, out_w + 2 * pad, 3), np.uint8)          # T.Pad: constant fill 0
    padded[pad:pad + out_h, pad:pad + out_w] = r
    c = padded[top:top + out_h, left:left + out_w]
    t = c.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)              # ToTensor
    t = (t - np.asarray(mean, np.float32)[:, None, None]) / np.asarray(std, np.float32)[:, None, None]
    if erase is not None:
        et, el, eh, ew = erase
        t[:, et:et + eh, el:el + ew] = noise
    return t.astype(np.float32)


def sample_params(rng, out_h, out_w, pad, flip_p, erase_p, min_area=0.02, max_area=1 / 3, min_aspect=0.3):
    """Draws with the reference's distributions (torchvision RandomHorizontalFlip / RandomCrop.get_params,
    make_dataloader.py:104-124 for the erase rectangle).  `rng` is a numpy Generator: the STREAM differs from the
    reference's (python `random` + torch CPU generator), the distributions do not."""
    flip = bool(rng.random() < flip_p)
    top = int(rng.integers(0, 2 * pad + 1))
    left = int(rng.integers(0, 2 * pad + 1))
    erase = None
    if not (rng.random() > erase_p):
        area = out_h * out_w
        la = (math.log(min_aspect), math.log(1 / min_aspect))
        for _ in range(10):
            target = rng.uniform(min_area, max_area) * area
            ar = math.exp(rng.uniform(*la))
            h = int(round(math.sqrt(target * ar)))
            w = int(round(math.sqrt(target / ar)))
            if w < out_w and h < out_h:
                erase = (int(rng.integers(0, out_h - h + 1)), int(rng.integers(0, out_w - w + 1)), h, w)
                break
    return flip, top, left, erase
