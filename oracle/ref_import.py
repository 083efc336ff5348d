"""Import the UNMODIFIED reference (`/root/reference/genrobo3d`) on CPU with offline shims.

TEST INFRASTRUCTURE — only usable in the build container (the GPU box has no /root/reference).
Used by `oracle/make_golden.py` to mint `tests/golden/*` and by the `not gpu` test that pins the
restatement `oracle/lotus_oracle.py` against the live reference when it is present.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("LOTUS_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "genrobo3d"))


def import_reference():
    """Returns the reference `genrobo3d.models.simple_policy_ptv3` / `motion_planner_ptv3` modules.

    The reference's flash branch casts q/k/v with `.half()` (model.py:544, model_ca.py:63); the fp32
    oracle keeps fp32 by making `Tensor.half` an identity for this process (SURVEY finding 4 / D.1).
    """
    import torch
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    for p in (REFERENCE_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    # drop a real flash_attn / spconv if something imported them earlier
    for name in list(sys.modules):
        if name.split(".")[0] in ("flash_attn", "spconv", "torch_scatter", "timm", "addict", "easydict", "yacs"):
            f = getattr(sys.modules[name], "__file__", "") or ""
            if not f.startswith(_SHIMS):
                del sys.modules[name]
    torch.Tensor.half = lambda self, *a, **k: self
    import genrobo3d.models.simple_policy_ptv3 as sp
    import genrobo3d.models.motion_planner_ptv3 as mp
    return sp, mp
