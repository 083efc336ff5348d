"""Mint golden vectors for the non-production variants from the UNMODIFIED reference (test infrastructure; run in the
build container):      python oracle/make_golden_variants.py      # writes tests/golden/variants.pt

Cases (reduced widths, `config.small_variant_config`): SimplePolicyPTV3AdaNorm (PDNorm),
SimplePolicyPTV3Concat, SimplePolicyPTV3CA with pose / step-id context tokens (row 15), MotionPlannerPTV3CA and
MotionPlannerPTV3AdaNorm (row 19).  Weights are re-derived on both sides by `testing.deterministic_reinit`.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402
from oracle.make_golden import GOLD, RandpermLog, clone_batch, deterministic_reinit, pkg  # noqa: E402


def run_case(cls, cfg, batch, reinit_seed):
    torch.manual_seed(0)
    model = cls(cfg.clone())
    deterministic_reinit(model, reinit_seed)
    model.train()
    feats = {}

    def hook(mod, inp, out):
        feats["dec_feat"] = out[-1].feat.detach().clone()
    h = model.ptv3_model.register_forward_hook(hook)
    with RandpermLog() as log:
        actions, losses = model(clone_batch(batch), compute_loss=True, compute_final_action=False)
    h.remove()
    losses["total"].backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    bn_after = {k: v.detach().clone() for k, v in model.state_dict().items() if "running_" in k}
    model.eval()
    with RandpermLog() as log_eval, torch.no_grad():
        final = model(clone_batch(batch), compute_loss=False)
    f = feats["dec_feat"]
    return {
        "perms": torch.stack(log.perms), "perms_eval": torch.stack(log_eval.perms),
        "losses": {k: v.detach().clone() for k, v in losses.items()},
        "train_actions": actions.detach().clone(), "final_actions_eval": final.detach().clone(),
        "dec_feat_head": f[:16].clone(), "dec_feat_abs_sum": f.double().abs().sum().item(), "dec_feat_shape": tuple(f.shape),
        "grad_norms": {k: g.norm().item() for k, g in grads.items()},
        "grads": {k: g for k, g in grads.items() if g.numel() <= 1200},
        "params_without_grad": [k for k, p in model.named_parameters() if p.grad is None],
        "bn_after": bn_after,
    }


def main():
    sp, mp = import_reference()
    cfgm, syn = pkg("config"), pkg("synthetic")
    out = {}
    cases = [
        ("adanorm", sp.SimplePolicyPTV3AdaNorm, lambda: syn.randomize_poses(syn.make_batch(seed=21, sizes=[260, 150, 200]), 5)),
        ("concat", sp.SimplePolicyPTV3Concat,
         lambda: syn.randomize_poses(syn.make_batch(seed=22, sizes=[240, 170], txt_len_range=(1, 1)), 6)),
        ("ca_tokens", sp.SimplePolicyPTV3CA, lambda: syn.randomize_poses(syn.make_batch(seed=23, sizes=[300, 140]), 7)),
        ("planner_ca", mp.MotionPlannerPTV3CA, lambda: syn.make_planner_batch(seed=24, sizes=[250, 180, 140])),
        ("planner_adanorm", mp.MotionPlannerPTV3AdaNorm, lambda: syn.make_planner_batch(seed=25, sizes=[220, 160])),
    ]
    for i, (kind, cls, mk) in enumerate(cases):
        cfg = cfgm.small_variant_config(kind)
        batch = mk()
        g = run_case(cls, cfg, batch, 9000 + i)
        g.update(batch=batch, config=dict(cfg.clone()), reinit_seed=9000 + i, model_class=cls.__name__, kind=kind)
        out[kind] = g
        print(kind, {k: round(float(v), 6) for k, v in g["losses"].items()}, "no-grad params:", g["params_without_grad"])
    path = os.path.join(GOLD, "variants.pt")
    torch.save(out, path)
    print("variants: file MB", os.path.getsize(path) / 1e6)


if __name__ == "__main__":
    main()
