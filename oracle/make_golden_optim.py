"""Mint optimizer golden vectors from the UNMODIFIED reference optimizer (test infrastructure; build container only).

    python oracle/make_golden_optim.py      # writes tests/golden/adamw.pt

VERDICT r1 weak #3: `lotus_oracle.adamw_step` restated train/optim/adamw.py:53-112 but was checked against nothing from the
reference.  Here the reference's own `build_optimizer` (train/optim/misc.py:13-55: names containing 'bias' are not decayed) and
`AdamW` class run four steps with `torch.nn.utils.clip_grad_norm_` (train_simple_policy.py:237-243) and a changing learning rate
on a small named-parameter set; parameters, both moments and the clip norm after every step are stored.
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SHAPES = [("stem.conv.weight", (8, 27, 3)), ("enc.block.norm1.0.weight", (16,)), ("enc.block.norm1.0.bias", (16,)),
          ("enc.block.attn.qkv.weight", (48, 16)), ("enc.block.attn.qkv.bias", (48,)), ("head.mlp.3.weight", (5, 16)),
          ("head.mlp.3.bias", (5,))]


class Holder(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(77)
        self.names = []
        for n, shp in SHAPES:
            self.register_parameter(n.replace(".", "__"), torch.nn.Parameter(torch.randn(shp, generator=g)))

    def named_parameters(self, *a, **k):      # the reference groups by NAME: give it the dotted names
        for (n, _), (_, p) in zip(SHAPES, super().named_parameters(*a, **k)):
            yield n, p


def main():
    import_reference()
    from genrobo3d.train.optim.misc import build_optimizer
    model = Holder()
    opts = types.SimpleNamespace(learning_rate=2e-3, weight_decay=0.05, optim="adamw", betas=(0.9, 0.98))
    opt, _ = build_optimizer(model, opts)
    params = [p for _, p in model.named_parameters()]
    out = {"shapes": SHAPES, "init": [p.detach().clone() for p in params], "lr": [], "grads": [], "norm": [], "params": [], "exp_avg": [],
           "exp_avg_sq": [], "max_grad_norm": 0.5, "weight_decay": 0.05, "betas": (0.9, 0.98),
           "groups": [[id(p) for p in g["params"]] for g in opt.param_groups]}
    ids = {id(p): i for i, p in enumerate(params)}
    out["groups"] = [[ids[i] for i in g] for g in out["groups"]]
    out["group_wd"] = [g["weight_decay"] for g in opt.param_groups]
    gen = torch.Generator().manual_seed(5)
    for step in range(1, 5):
        amp = [0.02, 5.0, 0.3, 1.0][step - 1]                     # below, far above, near the clip threshold
        grads = [torch.randn(p.shape, generator=gen) * amp / p.numel() ** 0.5 for p in params]
        for p, g in zip(params, grads):
            p.grad = g.clone()
        lr = 2e-3 * step / 4
        for g in opt.param_groups:
            g["lr"] = lr
        norm = torch.nn.utils.clip_grad_norm_(params, 0.5)
        opt.step()
        out["lr"].append(lr)
        out["grads"].append(grads)
        out["norm"].append(float(norm))
        out["params"].append([p.detach().clone() for p in params])
        out["exp_avg"].append([opt.state[p]["exp_avg"].clone() for p in params])
        out["exp_avg_sq"].append([opt.state[p]["exp_avg_sq"].clone() for p in params])
    torch.save(out, os.path.join(GOLD, "adamw.pt"))
    print("adamw.pt: groups", out["groups"], out["group_wd"], "norms", out["norm"])


if __name__ == "__main__":
    main()
