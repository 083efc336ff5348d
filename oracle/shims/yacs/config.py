"""Minimal stand-in for yacs==0.1.8 `CfgNode` (test infrastructure): attribute dict with
defrost()/freeze()/clone()/merge_from_list()."""
import copy


class CfgNode(dict):
    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        object.__setattr__(self, "_frozen", False)
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v, new_allowed=True) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        if k in self:
            return self[k]
        raise AttributeError(k)

    def __setattr__(self, k, v):
        if object.__getattribute__(self, "_frozen"):
            raise AttributeError("Attempted to set {} to {}, but CfgNode is immutable".format(k, v))
        self[k] = v

    def _set_frozen(self, f):
        object.__setattr__(self, "_frozen", f)
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(f)

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def is_frozen(self):
        return object.__getattribute__(self, "_frozen")

    def clone(self):
        return copy.deepcopy(self)
