"""GPU input pipeline: the reference's training transform

    T.Resize(SIZE_TRAIN, interpolation=3) -> T.RandomHorizontalFlip(PROB) -> T.Pad(PADDING) -> T.RandomCrop(SIZE_TRAIN) ->
    T.ToTensor() -> T.Normalize(PIXEL_MEAN, PIXEL_STD) -> RandomErasing(RE_PROB, mode='pixel', max_count=1)
    (data/datasets/make_dataloader.py:245-253, RandomErasing :55-140; applied per modality image, data/datasets/bases.py:100-103)

as ONE call over uint8 HWC batches on the device (`csrc/augment.cu` behind `edb_augment_u8`).  The host decodes (or holds)
uint8 images and ships 3 bytes per pixel instead of 12: at 3 300 samples/s x 3 modalities the 14-worker PIL pipeline of the
reference (configs/*/EDITOR.yml NUM_WORKERS) and a float32 H2D copy of 151 MB per step are what the hot path would wait for.

Python here is plumbing: the per-image random DRAWS (a flip coin, two crop offsets, the erase rectangle -- a handful of
scalars per image) and the resample coefficient tables.  The tables restate Pillow's `precompute_coeffs` /
`normalize_coeffs_8bpc` (libImaging/Resample.c; Pillow is the third-party dependency behind T.Resize, not part of the
reference tree): Keys bicubic a = -0.5, support 2 * max(scale, 1), coefficients normalised in double and rounded to 22-bit
fixed point.  The kernel output is bit-identical to torchvision's for given draws (tests/test_augment_gpu.py).
"""
import ctypes
import math

import numpy as np
import torch

from . import lib

PRECISION_BITS = 22


class AugImage(ctypes.Structure):
    _fields_ = [("flip", ctypes.c_int), ("top", ctypes.c_int), ("left", ctypes.c_int), ("e_top", ctypes.c_int),
                ("e_left", ctypes.c_int), ("e_h", ctypes.c_int), ("e_w", ctypes.c_int), ("seed_lo", ctypes.c_uint),
                ("seed_hi", ctypes.c_uint)]


PARAM_FIELDS = 9        # int32 words per EdbAugImage


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_tables(in_size, out_size):
    """(bounds int32 [out, 2] = (first tap, taps), coefficients int32 [out, ksize]) of Pillow's bicubic resample."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 2.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / fscale                       # Resample.c multiplies by the reciprocal: keep the same double arithmetic
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(v * (1 << PRECISION_BITS) + (0.5 if v >= 0 else -0.5))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


class GpuAugment:
    """Drop-in for the reference's `train_transforms` on whole batches.

        aug = GpuAugment(cfg, device)
        x = aug({"RGB": u8[B,Hs,Ws,3], "NI": ..., "TI": ...})     # uint8 device tensors -> float32 [B,3,H,W] per modality
    """

    def __init__(self, cfg, device, seed=0):
        self.H, self.W = int(cfg.INPUT.SIZE_TRAIN[0]), int(cfg.INPUT.SIZE_TRAIN[1])
        self.pad = int(cfg.INPUT.PADDING)
        self.flip_p, self.erase_p = float(cfg.INPUT.PROB), float(cfg.INPUT.RE_PROB)
        self.mean = (ctypes.c_float * 3)(*[float(v) for v in cfg.INPUT.PIXEL_MEAN])
        self.std = (ctypes.c_float * 3)(*[float(v) for v in cfg.INPUT.PIXEL_STD])
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.rng = np.random.default_rng(seed)
        self._tables = {}
        self._ws = None
        # RandomErasing defaults of the reference's call (make_dataloader.py:70-79,252)
        self.min_area, self.max_area, self.min_aspect = 0.02, 1.0 / 3.0, 0.3

    def tables(self, in_size, out_size):
        key = (in_size, out_size)
        if key not in self._tables:
            b, k = resample_tables(in_size, out_size)
            self._tables[key] = (torch.from_numpy(b).to(self.device), torch.from_numpy(k).to(self.device), k.shape[1])
        return self._tables[key]

    def sample(self, n):
        """Random draws for n images, int32 [n, 9] (EdbAugImage layout), with the reference's distributions: flip with
        probability PROB; crop offsets uniform on 0 .. 2*PADDING (RandomCrop.get_params on the padded image); with
        probability RE_PROB up to 10 attempts at a rectangle of area U(0.02, 1/3) x H x W and log-uniform aspect in
        [0.3, 1/0.3] that fits strictly inside the image (make_dataloader.py:104-124)."""
        H, W, pad, rng = self.H, self.W, self.pad, self.rng
        out = np.zeros((n, PARAM_FIELDS), np.int32)
        out[:, 0] = rng.random(n) < self.flip_p
        out[:, 1] = rng.integers(0, 2 * pad + 1, n)
        out[:, 2] = rng.integers(0, 2 * pad + 1, n)
        la = (math.log(self.min_aspect), math.log(1.0 / self.min_aspect))
        for i in np.nonzero(~(rng.random(n) > self.erase_p))[0]:
            for _ in range(10):
                target = rng.uniform(self.min_area, self.max_area) * H * W
                ar = math.exp(rng.uniform(*la))
                h, w = int(round(math.sqrt(target * ar))), int(round(math.sqrt(target / ar)))
                if w < W and h < H:
                    out[i, 3:7] = (rng.integers(0, H - h + 1), rng.integers(0, W - w + 1), h, w)
                    break
        out[:, 7:9] = rng.integers(0, 2 ** 31 - 1, (n, 2))
        return out

    def __call__(self, u8, params=None, noise=None, out=None):
        rgb, ni, ti = u8["RGB"], u8["NI"], u8["TI"]
        for t in (rgb, ni, ti):
            if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3 or not t.is_contiguous() or t.shape != rgb.shape \
                    or t.device != self.device:
                raise lib.EdbError("GpuAugment: inputs must be contiguous uint8 [B,Hs,Ws,3] tensors of one shape on %s" % self.device)
        B, Hs, Ws = rgb.shape[0], rgb.shape[1], rgb.shape[2]
        H, W = self.H, self.W
        if params is None:
            params = self.sample(3 * B)
        if not torch.is_tensor(params):
            # torch's caching pinned-memory allocator only recycles a block once the copy that used it has completed
            host = torch.from_numpy(np.ascontiguousarray(params, dtype=np.int32)).pin_memory()
            params = host.to(self.device, non_blocking=True)
        if params.shape != (3 * B, PARAM_FIELDS) or params.dtype != torch.int32:
            raise lib.EdbError("GpuAugment: params must be int32 [3*B, 9]")
        hb = hk = vb = vk = None
        ksh = ksv = 0
        if Ws != W:
            hb, hk, ksh = self.tables(Ws, W)
        if Hs != H:
            vb, vk, ksv = self.tables(Hs, H)
        nbytes = lib.load().edb_augment_workspace_bytes(B, Hs, Ws, W)
        if nbytes and (self._ws is None or self._ws.numel() < nbytes):
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        if out is None:
            out = {k: torch.empty(B, 3, H, W, dtype=torch.float32, device=self.device) for k in ("RGB", "NI", "TI")}
        if noise is not None and (noise.shape != (3 * B, 3, H, W) or noise.dtype != torch.float32 or not noise.is_contiguous()):
            raise lib.EdbError("GpuAugment: noise must be contiguous float32 [3*B, 3, H, W]")
        lib.call("edb_augment_u8", rgb.data_ptr(), ni.data_ptr(), ti.data_ptr(), B, Hs, Ws, H, W, self.pad, lib.ptr(hb),
                 lib.ptr(hk), ksh, lib.ptr(vb), lib.ptr(vk), ksv, ctypes.cast(self.mean, ctypes.c_void_p),
                 ctypes.cast(self.std, ctypes.c_void_p), params.data_ptr(), lib.ptr(noise), out["RGB"].data_ptr(),
                 out["NI"].data_ptr(), out["TI"].data_ptr(), lib.ptr(self._ws) if nbytes else None, nbytes, lib.stream_ptr())
        return out
