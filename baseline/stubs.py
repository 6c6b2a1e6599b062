"""Stand-ins for the four third-party modules the UNMODIFIED reference imports and this image lacks:

* ``yacs``       -> tiny dict-backed ``CfgNode`` (config/defaults.py:1 only constructs + assigns)
* ``pywt``       -> ``Wavelet('haar')`` taps + ``dwt_coeff_len`` (pytorch_wavelets/dwt/transform2d.py:2,23;
                    lowlevel.py:6,153) -- the only third-party arithmetic on the path is 1/sqrt(2)
* ``matplotlib`` / ``seaborn`` -> empty modules (vit_pytorch.py:26,34,35; Frequency.py:3; utils/metrics.py:7 -- dead viz code)

Measurement / test infrastructure only: used by ``baseline/run_ref.py`` (GPU baseline of the reference) and by
``oracle/ref_import.py`` (golden generation).  Nothing in ``editor_b200/`` imports this file.
"""
import math
import sys
import types

import yaml


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


def install_stubs():
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
