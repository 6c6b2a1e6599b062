"""Deterministic synthetic weights and 3-modal batches (no datasets / checkpoints offline).

* ``state_dict_schema`` enumerates the reference's 222 ``state_dict`` keys and shapes
  (reference: modeling/make_model.py:86-141, vit_pytorch.py:261-296,
  420-520, OCFR.py:14-16, pytorch_wavelets/dwt/transform2d.py:36-40,101-105).
* ``synthetic_state_dict`` fills that schema from a seeded CPU generator so that the
  build container (where the unmodified reference produces the golden vectors) and the
  GPU box (where the CUDA path and the oracle are compared) hold bit-identical weights
  without shipping 480 MB.
* ``synthetic_batch`` builds uint8-quantised RGB/NI/TI images as real data would be
  (SURVEY.md 8(d) d-1): the mean of 9 quantised values can never be 0, so the
  frequency-branch threshold (Frequency.py:52) is well defined in every precision.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

DIM = 768
HEADS = 12
DEPTH = 12
HIDDEN = 3072


def state_dict_schema(num_class=171, camera_num=4, num_tokens=129, al=True):
    """Ordered {key: (shape, dtype, kind)} identical to the reference EDITOR.state_dict()."""
    s = OrderedDict()

    def add(k, shape, kind, dtype=torch.float32):
        s[k] = (tuple(shape), dtype, kind)

    b = "BACKBONE.base."
    add(b + "cls_token", (1, 1, DIM), "embed")
    add(b + "pos_embed", (1, num_tokens, DIM), "embed")
    if camera_num > 1:
        add(b + "sie_embed", (camera_num, 1, DIM), "embed")
    add(b + "patch_embed.proj.weight", (DIM, 3, 16, 16), "conv")
    add(b + "patch_embed.proj.bias", (DIM,), "bias")
    for i in range(DEPTH):
        p = b + "blocks.%d." % i
        add(p + "norm1.weight", (DIM,), "ln_w")
        add(p + "norm1.bias", (DIM,), "ln_b")
        add(p + "attn.qkv.weight", (3 * DIM, DIM), "linear")
        add(p + "attn.qkv.bias", (3 * DIM,), "bias")
        add(p + "attn.proj.weight", (DIM, DIM), "linear")
        add(p + "attn.proj.bias", (DIM,), "bias")
        add(p + "norm2.weight", (DIM,), "ln_w")
        add(p + "norm2.bias", (DIM,), "ln_b")
        add(p + "mlp.fc1.weight", (HIDDEN, DIM), "linear")
        add(p + "mlp.fc1.bias", (HIDDEN,), "bias")
        add(p + "mlp.fc2.weight", (DIM, HIDDEN), "linear")
        add(p + "mlp.fc2.bias", (DIM,), "bias")
    add(b + "norm.weight", (DIM,), "ln_w")
    add(b + "norm.bias", (DIM,), "ln_b")
    add(b + "fc.weight", (1000, DIM), "linear")
    add(b + "fc.bias", (1000,), "bias")
    for w in ("DWT", "IDWT"):
        pre = "h" if w == "DWT" else "g"
        add("FREQ_INDEX.%s.%s0_col" % (w, pre), (1, 1, 2, 1), "haar")
        add("FREQ_INDEX.%s.%s1_col" % (w, pre), (1, 1, 2, 1), "haar")
        add("FREQ_INDEX.%s.%s0_row" % (w, pre), (1, 1, 1, 2), "haar")
        add("FREQ_INDEX.%s.%s1_row" % (w, pre), (1, 1, 1, 2), "haar")
    f = "FUSE_block."
    # registration order follows BlockMask.__init__ (vit_pytorch.py:266-297)
    for m, (n1, at, n2, ml) in (("R", ("normR", "attnR", "normR_", "mlpR")),
                                 ("N", ("normN", "attnN", "normN_", "mlpN")),
                                 ("T", ("normT", "attnT", "normT_", "mlpT")),
                                 ("J", ("norm1", "attn1", "norm2", "mlp"))):
        add(f + n1 + ".weight", (DIM,), "ln_w")
        add(f + n1 + ".bias", (DIM,), "ln_b")
        add(f + at + ".qkv.weight", (3 * DIM, DIM), "linear")
        add(f + at + ".proj.weight", (DIM, DIM), "linear")
        add(f + n2 + ".weight", (DIM,), "ln_w")
        add(f + n2 + ".bias", (DIM,), "ln_b")
        add(f + ml + ".fc1.weight", (HIDDEN, DIM), "linear")
        add(f + ml + ".fc2.weight", (DIM, HIDDEN), "linear")
    add(f + "out_norm.weight", (DIM,), "ln_w")
    add(f + "out_norm.bias", (DIM,), "ln_b")
    for m in ("RGB", "NIR", "TIR"):
        add(f + "memory_cls.%s_centers" % m, (num_class, DIM), "center")
    for m in ("RGB", "NIR", "TIR"):
        add("%s_REDUCE.weight" % m, (DIM, 2 * DIM), "reduce")
        add("%s_REDUCE.bias" % m, (DIM,), "bias")

    def bn(name, n):
        add(name + ".weight", (n,), "ln_w")
        add(name + ".bias", (n,), "ln_b")
        add(name + ".running_mean", (n,), "bn_mean")
        add(name + ".running_var", (n,), "bn_var")
        add(name + ".num_batches_tracked", (), "count", torch.int64)

    add("FUSE_HEAD.weight", (num_class, 3 * DIM), "head")
    bn("FUSE_BN", 3 * DIM)
    add("BACKBONE_HEAD.weight", (num_class, DIM), "head")
    bn("BACKBONE_BN", DIM)
    if al:
        add("AL_HEAD.weight", (num_class, 3 * DIM), "head")
        bn("AL_BN", 3 * DIM)
    return s


_HAAR = {
    # pytorch_wavelets/dwt/lowlevel.py prep_filt_afb2d reverses the pywt taps, prep_filt_sfb2d does not
    "h0": [1.0, 1.0], "h1": [1.0, -1.0], "g0": [1.0, 1.0], "g1": [1.0, -1.0],
}


def synthetic_state_dict(seed=1111, num_class=171, camera_num=4, num_tokens=129, al=True,
                         zero_centers=False):
    """Seeded, well-conditioned weights for every key of the schema (CPU float32).

    Scales: linears N(0, 0.02) like the reference's trunc_normal_(std=.02)
    (vit_pytorch.py:524-531) -- except ``*_REDUCE`` at 0.02 and heads at 0.02 so that
    ``||cls4t||`` stays O(1) (at raw kaiming init the soft-margin triplet overflows);
    LayerNorm/BN affine parameters are perturbed away from (1, 0) so that they matter.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    s = 1.0 / math.sqrt(2.0)
    out = OrderedDict()
    for k, (shape, dtype, kind) in state_dict_schema(num_class, camera_num, num_tokens, al).items():
        if kind == "haar":
            taps = _HAAR[k.split(".")[-1][:2]]
            out[k] = (torch.tensor(taps, dtype=torch.float32) * s).reshape(shape)
        elif kind == "count":
            out[k] = torch.zeros((), dtype=torch.int64)
        elif kind in ("linear", "embed", "bias", "head", "reduce"):
            out[k] = torch.randn(shape, generator=g) * 0.02
        elif kind == "conv":
            out[k] = torch.randn(shape, generator=g) * math.sqrt(2.0 / (16 * 16 * DIM)) * 2.0
        elif kind == "ln_w":
            out[k] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "ln_b":
            out[k] = 0.05 * torch.randn(shape, generator=g)
        elif kind == "bn_mean":
            out[k] = 0.05 * torch.randn(shape, generator=g)
        elif kind == "bn_var":
            out[k] = 1.0 + 0.2 * torch.rand(shape, generator=g)
        elif kind == "center":
            c = torch.randn(shape, generator=g) * 0.02
            out[k] = torch.zeros(shape) if zero_centers else c
        else:  # pragma: no cover
            raise KeyError(kind)
    return out


def synthetic_batch(batch, height=256, width=128, seed=1, num_cams=4, ids=None, instances=None):
    """uint8-quantised 3-modal batch + P x K labels + camera ids (CPU tensors).

    img = bilinear_up16(randn[B,3,H/16,W/16]) + 0.5 randn[B,3,H,W]; p = round(clamp(.25 img+.5,0,1)*255);
    x = (p/255 - .5)/.5  (normalisation of config/defaults.py:70-72).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x = {}
    for name in ("RGB", "NI", "TI"):
        low = torch.randn(batch, 3, height // 16, width // 16, generator=g)
        img = F.interpolate(low, scale_factor=16, mode="bilinear", align_corners=False)
        img = img + 0.5 * torch.randn(batch, 3, height, width, generator=g)
        p = torch.round(torch.clamp(0.25 * img + 0.5, 0.0, 1.0) * 255.0)
        x[name] = (p / 255.0 - 0.5) / 0.5
    if instances is None:
        instances = 16 if batch % 16 == 0 else (2 if batch % 2 == 0 else 1)
    n_ids = batch // instances
    label = torch.arange(n_ids, dtype=torch.int64).repeat_interleave(instances)
    if ids is not None:
        label = torch.as_tensor(ids, dtype=torch.int64)[label]
    g2 = torch.Generator(device="cpu")
    g2.manual_seed(seed + 1)
    cam = torch.randint(0, max(num_cams, 1), (batch,), generator=g2, dtype=torch.int64)
    return x, label, cam
