"""Mint golden vectors from the UNMODIFIED reference (test infrastructure; run in the build container).

    python oracle/make_golden.py            # writes tests/golden/*.pt

The reference ships no tests / fixtures, so parity is pinned on outputs of the reference's
own Python code imported through `oracle/ref_import.py` (fp32, CPU, flash-path semantics).  The files
written here are the fixtures `tests/` checks both the restatement (`oracle/lotus_oracle.py`) and the CUDA
path against.  Nothing on the GPU box reads /root/reference.
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def pkg(name):
    return importlib.import_module("robot-3dlotus_b200." + name)


def deterministic_reinit(model, seed):
    """Both sides (reference here, CUDA module on the GPU box) re-draw every tensor of the state dict in
    name order from one seeded CPU generator, so big models need no weight fixture."""
    return pkg("testing").deterministic_reinit(model, seed)


class RandpermLog:
    """records the `torch.randperm(4)` draws of `shuffle_orders` (model.py:130-134, 750-754)."""

    def __init__(self):
        self.perms = []
        self._orig = torch.randperm

    def __enter__(self):
        def rp(n, *a, **k):
            p = self._orig(n, *a, **k)
            self.perms.append(p.clone())
            return p
        torch.randperm = rp
        return self

    def __exit__(self, *a):
        torch.randperm = self._orig


def capture_hooks(model, store):
    hs = []

    def grab_point(name, full=False):
        def hook(mod, inp, out):
            p = out
            d = {"feat": p.feat.detach().clone()}
            if full:
                for k in ("coord", "grid_coord", "serialized_code", "serialized_order", "serialized_inverse",
                          "batch", "offset", "pooling_inverse"):
                    if k in p.keys():
                        d[k] = p[k].detach().clone()
                d["serialized_depth"] = int(p.serialized_depth)
            for k in ("pad", "unpad", "cu_seqlens_key"):
                if k in p.keys():
                    d[k] = p[k].detach().clone()
            store[name] = d
        return hook

    pt = model.ptv3_model
    hs.append(pt.embedding.register_forward_hook(grab_point("embedding", full=True)))
    for sname, stage in list(pt.enc._modules.items()) + list(pt.dec._modules.items()):
        for mname, mod in stage._modules.items():
            hs.append(mod.register_forward_hook(grab_point(f"{sname}.{mname}", full=(mname in ("down", "up")))))
    return hs


def clone_batch(b):
    out = {}
    for k, v in b.items():
        if isinstance(v, torch.Tensor):
            out[k] = v.clone()
        elif isinstance(v, list) and v and isinstance(v[0], torch.Tensor):
            out[k] = [t.clone() for t in v]
        else:
            out[k] = list(v) if isinstance(v, list) else v
    return out


def run_policy_case(sp, cfg, batch, seed, store_weights, reinit_seed=None, grads_to_keep=()):
    torch.manual_seed(seed)
    model = sp.SimplePolicyPTV3CA(cfg)
    if reinit_seed is not None:
        deterministic_reinit(model, reinit_seed)
    model.train()
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    inter = {}
    hooks = capture_hooks(model, inter)
    with RandpermLog() as log:
        actions, losses = model(clone_batch(batch), compute_loss=True, compute_final_action=False)
    for h in hooks:
        h.remove()
    losses["total"].backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    bn_after = {k: v.detach().clone() for k, v in model.state_dict().items()
                if "running_" in k or "num_batches" in k}
    # inference pass: eval mode (BN running stats), final action decode
    model.eval()
    with RandpermLog() as log_eval, torch.no_grad():
        final = model(clone_batch(batch), compute_loss=False)
    out = {
        "perms": torch.stack(log.perms), "perms_eval": torch.stack(log_eval.perms),
        "losses": {k: v.detach().clone() for k, v in losses.items()},
        "train_actions": actions.detach().clone(),
        "final_actions_eval": final.detach().clone(),
        "inter": inter,
        "grad_norms": {k: g.norm().item() for k, g in grads.items()},
        "grads": {k: grads[k] for k in grads_to_keep} if grads_to_keep else grads,
        "bn_after": bn_after,
    }
    if store_weights:
        out["state_dict"] = state0
    return out


def main():
    sp, mp = import_reference()
    cfgm, syn = pkg("config"), pkg("synthetic")
    os.makedirs(GOLD, exist_ok=True)

    # ---------------------------------------------------------------- policy, small width, weights shipped
    cfg = cfgm.small_model_config(attn_drop=0.0, proj_drop=0.0)
    cfg.action_config.dropout = 0.0
    batch = syn.make_batch(seed=11, sizes=[700, 130, 300, 90])
    g = run_policy_case(sp, cfg, batch, seed=0, store_weights=False, reinit_seed=4321)
    g["batch"] = batch
    g["config"] = dict(cfg.clone())
    g["reinit_seed"] = 4321
    # keep the file small: drop the bulkiest grads but keep all norms
    keep = [k for k in g["grads"] if g["grads"][k].numel() <= 5000]
    g["grads"] = {k: g["grads"][k] for k in keep}
    torch.save(g, os.path.join(GOLD, "policy_small.pt"))
    print("policy_small: losses", {k: float(v) for k, v in g["losses"].items()},
          "file MB", os.path.getsize(os.path.join(GOLD, "policy_small.pt")) / 1e6)

    # ---------------------------------------------------------------- policy, production width, weights re-derived
    cfg = cfgm.production_model_config(attn_drop=0.0, proj_drop=0.0)
    cfg.action_config.dropout = 0.0
    batch = syn.make_batch(seed=12, sizes=[600, 257])
    g = run_policy_case(sp, cfg, batch, seed=0, store_weights=False, reinit_seed=1234,
                        grads_to_keep=("txt_fc.bias", "act_proj_head.action_mlp.3.bias",
                                       "ptv3_model.embedding.stem.conv.weight",
                                       "ptv3_model.enc.enc4.block0.attn.q_norm.weight",
                                       "ptv3_model.dec.dec0.ca_block0.attn.kv.bias"))
    g["batch"] = batch
    g["config"] = dict(cfg.clone())
    g["reinit_seed"] = 1234
    slim = {}
    for k, v in g["inter"].items():
        f = v["feat"]
        slim[k] = {"feat_head": f[:8].clone(), "feat_sum": f.double().sum().item(),
                   "feat_abs_sum": f.double().abs().sum().item(), "shape": tuple(f.shape)}
    g["inter"] = slim
    torch.save(g, os.path.join(GOLD, "policy_production.pt"))
    print("policy_production: losses", {k: float(v) for k, v in g["losses"].items()},
          "file MB", os.path.getsize(os.path.join(GOLD, "policy_production.pt")) / 1e6)

    # ---------------------------------------------------------------- unit-level: codes, patch plan, RPE, FPS, decode
    from genrobo3d.models.PointTransformerV3.serialization import encode
    from genrobo3d.models.PointTransformerV3 import model as M
    units = {}
    rng = np.random.default_rng(5)
    codes = {}
    for depth in (1, 3, 7, 9, 12, 16):
        gc = torch.from_numpy(rng.integers(0, 2 ** depth, size=(257, 3))).int()
        gc[0] = 0
        gc[1] = 2 ** depth - 1
        bt = torch.from_numpy(rng.integers(0, 5, size=(257,))).long()
        codes[depth] = {"grid": gc, "batch": bt,
                        **{o: encode(gc, bt, depth, order=o) for o in ("z", "z-trans", "hilbert", "hilbert-trans")}}
    units["codes"] = codes
    plans = {}
    attn = M.SerializedAttention(channels=32, num_heads=2, patch_size=128, enable_flash=True,
                                 upcast_attention=False, upcast_softmax=False)
    for name, counts in {"mixed": [700, 130, 300, 90, 128, 129, 256, 1], "single_small": [17],
                         "exact": [128, 256, 384], "big": [4096, 3000]}.items():
        p = M.Point(offset=torch.tensor(np.cumsum(counts)).long())
        pad, unpad, cu = attn.get_padding_and_inverse(p)
        plans[name] = {"counts": counts, "pad": pad, "unpad": unpad, "cu_seqlens": cu}
    units["patch_plan"] = plans
    torch.manual_seed(3)
    rpe = M.RPE(128, 4)
    rel = torch.from_numpy(rng.integers(-40, 41, size=(2, 128, 128, 3))).long()
    units["rpe"] = {"table": rpe.rpe_table.detach().clone(), "pos_bnd": rpe.pos_bnd, "rel": rel[:, :16, :16].clone(),
                    "bias": rpe(rel[:, :16, :16]).detach().clone()}
    from genrobo3d.train.datasets.common import farthest_point_sample
    pts = rng.standard_normal((500, 6)).astype(np.float32)
    np.random.seed(7)
    start = np.random.randint(0, 500)
    np.random.seed(7)
    units["fps"] = {"points": torch.from_numpy(pts), "npoint": 64, "start": start,
                    "sampled": torch.from_numpy(farthest_point_sample(pts, 64))}
    from genrobo3d.utils.action_position_utils import get_disc_gt_pos_prob, get_best_pos_from_disc_pos
    from genrobo3d.utils.rotation_transform import discrete_euler_to_quaternion
    xyz = rng.uniform(-0.3, 0.3, size=(200, 3)).astype(np.float32)
    gt = xyz[17] + np.array([0.003, -0.002, 0.004], dtype=np.float32)
    prob = get_disc_gt_pos_prob(xyz, gt, pos_bin_size=0.01, pos_bins=15, heatmap_type="dist")
    units["pos_codec"] = {"xyz": torch.from_numpy(xyz), "gt": torch.from_numpy(gt), "prob": torch.from_numpy(prob),
                          "best": torch.from_numpy(get_best_pos_from_disc_pos(prob, xyz, pos_bin_size=0.01, pos_bins=15))}
    bins = rng.integers(0, 72, size=(16, 3))
    units["rot_codec"] = {"bins": torch.from_numpy(bins),
                          "quat": torch.from_numpy(np.stack([discrete_euler_to_quaternion(b, 5) for b in bins]))}
    torch.save(units, os.path.join(GOLD, "units.pt"))
    print("units: file MB", os.path.getsize(os.path.join(GOLD, "units.pt")) / 1e6)


if __name__ == "__main__":
    main()
