"""Small yacs-compatible CfgNode (yacs is not installed here; the reference's config/defaults.py:1 only needs
construction, attribute access, merge_from_file / merge_from_list, clone and freeze)."""
import copy

import yaml


class CfgNode(dict):
    def __init__(self, init=None):
        super().__init__()
        self.__dict__["_frozen"] = False
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        if self.__dict__.get("_frozen", False):
            raise AttributeError("attempted to set %s on a frozen CfgNode" % k)
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            out[k] = copy.deepcopy(v, memo)
        return out

    def _merge(self, other, path=""):
        for k, v in other.items():
            if k not in self:
                raise KeyError("Non-existent config key: %s%s" % (path, k))
            if isinstance(v, dict):
                self[k]._merge(v, path + k + ".")
            else:
                if isinstance(self[k], tuple) and isinstance(v, list):
                    v = tuple(v)
                self[k] = v

    def merge_from_file(self, path):
        with open(path) as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, opts):
        opts = list(opts or [])
        assert len(opts) % 2 == 0, "override list must have an even number of entries"
        for k, v in zip(opts[0::2], opts[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            if parts[-1] not in node:
                raise KeyError("Non-existent config key: %s" % k)
            if isinstance(v, str):
                try:
                    v = yaml.safe_load(v)
                except yaml.YAMLError:
                    pass
            node[parts[-1]] = v

    def freeze(self):
        self.__dict__["_frozen"] = True
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def defrost(self):
        self.__dict__["_frozen"] = False
        for v in self.values():
            if isinstance(v, CfgNode):
                v.defrost()
