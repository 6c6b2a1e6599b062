"""GPU input pipeline (row f-4) against torchvision's transforms on PIL images -- the reference's own pipeline
(data/datasets/make_dataloader.py:245-253) -- and against the oracle, for given random draws: bit-exact."""
import numpy as np
import pytest
import torch
from PIL import Image

import __graft_entry__ as ge
from oracle import augment_oracle as ao

pytestmark = pytest.mark.gpu


def _torchvision(img, h, w, flip, top, left, erase, noise):
    import torchvision.transforms as T
    import torchvision.transforms.functional as F
    pil = T.Resize([h, w], interpolation=3)(Image.fromarray(img))
    if flip:
        pil = F.hflip(pil)
    pil = F.crop(T.Pad(10)(pil), top, left, h, w)
    t = T.Normalize(mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5])(T.ToTensor()(pil))
    if erase is not None:
        et, el, eh, ew = erase
        t[:, et:et + eh, el:el + ew] = noise[:, et:et + eh, el:el + ew]
    return t


@pytest.mark.parametrize("hs,ws,al", [(300, 140, True), (256, 128, True), (128, 256, False), (90, 300, False), (511, 257, True)])
def test_augment_matches_torchvision_bit_exact(hs, ws, al):
    from editor_b200 import data
    model = ge._small_case(al, 2)[0]
    from editor_b200.config import cfg
    c = cfg.clone()
    c.INPUT.SIZE_TRAIN = [256, 128] if al else [128, 256]
    h, w = c.INPUT.SIZE_TRAIN
    B = 5
    rng = np.random.default_rng(hs + ws)
    u8 = {k: rng.integers(0, 256, (B, hs, ws, 3), dtype=np.uint8) for k in ("RGB", "NI", "TI")}
    aug = data.GpuAugment(c, "cuda", seed=5)
    params = aug.sample(3 * B)
    params[0, 3:7] = 0                      # at least one image without erasing ...
    params[1, 3:7] = (3, 5, 40, 30)         # ... and one with a known rectangle
    noise = torch.randn(3 * B, 3, h, w, generator=torch.Generator().manual_seed(1))
    out = aug({k: torch.from_numpy(v).cuda() for k, v in u8.items()}, params=params, noise=noise.cuda())
    torch.cuda.synchronize()
    for m, name in enumerate(("RGB", "NI", "TI")):
        for b in range(B):
            p = params[m * B + b]
            erase = tuple(int(v) for v in p[3:7]) if p[5] > 0 else None
            ref = _torchvision(u8[name][b], h, w, bool(p[0]), int(p[1]), int(p[2]), erase, noise[m * B + b])
            got = out[name][b].cpu()
            assert torch.equal(got, ref), (name, b, (got - ref).abs().max())           # byte / IEEE work: bit-exact
            orc = ao.augment(u8[name][b], h, w, bool(p[0]), int(p[1]), int(p[2]), 10, (0.5,) * 3, (0.5,) * 3, erase,
                             None if erase is None else noise[m * B + b][:, erase[0]:erase[0] + erase[2], erase[1]:erase[1] + erase[3]].numpy())
            assert np.array_equal(got.numpy(), orc)
    del model


def test_augment_device_noise_is_standard_normal_and_confined():
    """Without injected noise the erase rectangles are filled by the kernel's Philox4x32-10 + Box-Muller stream: N(0,1)
    inside the rectangle, untouched pixels outside, different per image / channel, reproducible for equal seeds."""
    from editor_b200 import data
    from editor_b200.config import cfg
    c = cfg.clone()
    aug = data.GpuAugment(c, "cuda", seed=9)
    B, h, w = 64, 256, 128
    u8 = {k: torch.full((B, h, w, 3), 128, dtype=torch.uint8, device="cuda") for k in ("RGB", "NI", "TI")}
    params = aug.sample(3 * B)
    params[:, 0:3] = (0, 10, 10)            # no flip, centred crop: every non-erased pixel is (128/255 - .5)/.5
    params[:, 3:7] = (16, 8, 200, 100)
    a = aug(u8, params=params)
    b = aug(u8, params=params)
    base = (128.0 / 255.0 - 0.5) / 0.5
    for name in ("RGB", "NI", "TI"):
        assert torch.equal(a[name], b[name])
        x = a[name]
        inside = x[:, :, 16:216, 8:108]
        mask = torch.ones_like(x, dtype=torch.bool)
        mask[:, :, 16:216, 8:108] = False
        assert torch.all((x[mask] - base).abs() < 1e-6)
        assert abs(inside.mean().item()) < 5e-3 and abs(inside.std().item() - 1.0) < 5e-3
        assert abs((inside ** 4).mean().item() - 3.0) < 0.05                      # normal kurtosis
        assert not torch.equal(inside[0, 0], inside[0, 1]) and not torch.equal(inside[0, 0], inside[1, 0])
    assert not torch.equal(a["RGB"][:, :, 16:216, 8:108], a["NI"][:, :, 16:216, 8:108])
