"""Mint the BASELINE config-2-sized float fixture from the UNMODIFIED reference (test infrastructure; build container only).

    python oracle/make_golden_big.py        # writes tests/golden/policy_production_big.pt

VERDICT r1 weak #2: the production-width golden held two clouds of 600 and 257 points, so the deep stages
(~130 / ~42 points per cloud, many short sequences) were never compared in float at real size.  This fixture runs the
reference SimplePolicyPTV3CA (production width, fp32, CPU, dropout 0) on two keysteps of 4096 and 3000 points -- the
per-keystep sizes of BASELINE.json configs[1] -- forward + backward.  The batch is NOT stored (it is regenerated from
(seed, sizes) by robot-3dlotus_b200/synthetic.py and verified by checksum); stored are the losses, per-module output
digests (first rows, sums), every gradient norm, a few full gradients and the BatchNorm statistics.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402
from oracle.make_golden import GOLD, pkg, run_policy_case  # noqa: E402

SEED, SIZES = 21, [4096, 3000]


def main():
    sp, _ = import_reference()
    cfgm, syn = pkg("config"), pkg("synthetic")
    cfg = cfgm.production_model_config(attn_drop=0.0, proj_drop=0.0)
    cfg.action_config.dropout = 0.0
    batch = syn.make_batch(seed=SEED, sizes=SIZES)
    g = run_policy_case(sp, cfg, batch, seed=0, store_weights=False, reinit_seed=4242,
                        grads_to_keep=("txt_fc.bias", "act_proj_head.action_mlp.3.bias", "act_proj_head.heatmap_mlp.3.bias",
                                       "ptv3_model.enc.enc4.block0.attn.q_norm.weight", "ptv3_model.enc.enc3.block0.norm1.0.weight",
                                       "ptv3_model.dec.dec0.ca_block0.attn.kv.bias", "ptv3_model.enc.enc0.block0.cpe.0.bias"))
    g["batch_seed"], g["batch_sizes"] = SEED, SIZES
    g["batch_digest"] = {"pc_fts_sum": batch["pc_fts"].double().sum().item(), "pc_fts_abs_sum": batch["pc_fts"].double().abs().sum().item(),
                         "txt_sum": batch["txt_embeds"].double().sum().item(), "gt_sum": batch["gt_actions"].double().sum().item(),
                         "prob_sum": float(sum(p.double().sum() for p in batch["disc_pos_probs"])), "txt_lens": list(batch["txt_lens"])}
    g["config"] = dict(cfg.clone())
    g["reinit_seed"] = 4242
    slim = {}
    for k, v in g["inter"].items():
        f = v["feat"]
        slim[k] = {"feat_head": f[:8].clone(), "feat_tail": f[-8:].clone(), "feat_sum": f.double().sum().item(),
                   "feat_abs_sum": f.double().abs().sum().item(), "shape": tuple(f.shape)}
    g["inter"] = slim
    path = os.path.join(GOLD, "policy_production_big.pt")
    torch.save(g, path)
    print("policy_production_big: losses", {k: float(v) for k, v in g["losses"].items()}, "stage rows",
          [slim[k]["shape"][0] for k in slim if k.endswith("block0") and k.startswith("enc")], "file MB", os.path.getsize(path) / 1e6)


if __name__ == "__main__":
    main()
