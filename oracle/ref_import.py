"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference (/root/reference) on CPU.

Used only in the build container (the GPU box has no /root/reference) by
``tests/golden/make_golden.py`` to generate golden vectors and by the optional
``tests/test_oracle_vs_reference.py`` cross-check.  Nothing in ``editor_b200/``
imports this file.

The reference needs four third-party modules that are not in this image and
hard-codes ``.cuda()`` in six places:

* ``yacs``       -> tiny dict-backed ``CfgNode`` (config/defaults.py:1 only constructs + assigns)
* ``pywt``       -> ``Wavelet('haar')`` taps + ``dwt_coeff_len`` (pytorch_wavelets/dwt/transform2d.py:2,23;
                    lowlevel.py:6,153) -- the only third-party arithmetic on the path is 1/sqrt(2)
* ``matplotlib`` / ``seaborn`` -> empty modules (vit_pytorch.py:26,34,35; Frequency.py:3 -- dead viz code)
* ``Tensor.cuda`` / ``Module.cuda`` -> identity on CPU
"""
import math
import os
import sys
import types

import torch
import yaml

REF_ROOT = os.environ.get("EDITOR_REFERENCE_ROOT", "/root/reference")


from baseline.stubs import _CfgNode, install_stubs as _install_stubs  # noqa: E402,F401


_PATCHED = {}


def patch_cuda_identity():
    """CPU only: make ``.cuda()`` a no-op (the reference calls it unconditionally)."""
    if not _PATCHED:
        _PATCHED["t"] = torch.Tensor.cuda
        _PATCHED["m"] = torch.nn.Module.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self


def unpatch_cuda_identity():
    if _PATCHED:
        torch.Tensor.cuda = _PATCHED.pop("t")
        torch.nn.Module.cuda = _PATCHED.pop("m")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "modeling"))


def load_reference(dataset="RGBNT201", num_class=171, camera_num=4, opts=(), cpu=True):
    """Return (model, cfg) of the unmodified reference built from its own yml."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    if cpu:
        patch_cuda_identity()
    # our repo root also has packages named `modeling`/`config` (drop-in shims): the
    # reference must win inside this helper, so purge and put it first on sys.path.
    for name in list(sys.modules):
        if name.split(".")[0] in ("modeling", "config", "pytorch_wavelets", "layers", "solver"):
            del sys.modules[name]
    sys.path.insert(0, REF_ROOT)
    try:
        from config import cfg as _cfg  # noqa
        cfg = _cfg.clone()
        cfg.merge_from_file(os.path.join(REF_ROOT, "configs", dataset, "EDITOR.yml"))
        cfg.MODEL.PRETRAIN_CHOICE = "none"
        cfg.merge_from_list(list(opts))
        from modeling import make_model  # noqa
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            model = make_model(cfg, num_class=num_class, camera_num=camera_num)
    finally:
        sys.path.remove(REF_ROOT)
        for name in list(sys.modules):
            if name.split(".")[0] in ("modeling", "config"):
                # keep module objects alive through `model`, but free the names for our shims
                del sys.modules[name]
    return model, cfg


class NullWriter:
    """Stands in for the SummaryWriter the training forward expects (make_model.py:200)."""

    def __init__(self):
        self.scalars = []

    def add_scalar(self, tag, value, step=None):
        self.scalars.append((tag, float(value), step))
