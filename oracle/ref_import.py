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


class _CfgNode(dict):
    """Minimal yacs.config.CfgNode stand-in (attribute access, yaml merge, list merge)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        out = _CfgNode()
        for k, v in self.items():
            out[k] = v.clone() if isinstance(v, _CfgNode) else (list(v) if isinstance(v, list) else v)
        return out

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self:
                    self[k] = _CfgNode()
                self[k]._merge(v)
            else:
                self[k] = v

    def merge_from_file(self, path):
        with open(path) as f:
            self._merge(yaml.safe_load(f))

    def merge_from_list(self, opts):
        assert len(opts) % 2 == 0
        for k, v in zip(opts[0::2], opts[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            if isinstance(v, str):
                try:
                    v = yaml.safe_load(v)
                except Exception:
                    pass
            node[parts[-1]] = v

    def freeze(self):
        pass

    def defrost(self):
        pass


def _install_stubs():
    if "yacs" not in sys.modules:
        yacs = types.ModuleType("yacs")
        yc = types.ModuleType("yacs.config")
        yc.CfgNode = _CfgNode
        yacs.config = yc
        sys.modules["yacs"] = yacs
        sys.modules["yacs.config"] = yc
    if "pywt" not in sys.modules:
        pywt = types.ModuleType("pywt")
        s = 1.0 / math.sqrt(2.0)

        class Wavelet:  # noqa: D401 - haar only
            def __init__(self, name):
                assert name in ("haar", "db1"), name
                self.dec_lo = [s, s]
                self.dec_hi = [-s, s]
                self.rec_lo = [s, s]
                self.rec_hi = [s, -s]

        pywt.Wavelet = Wavelet
        pywt.dwt_coeff_len = lambda N, L, mode="zero": (N + L - 1) // 2
        sys.modules["pywt"] = pywt
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


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
