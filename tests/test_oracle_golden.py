"""Pins the CPU oracle (`oracle/lotus_oracle.py`) against golden vectors minted from the unmodified
reference (`oracle/make_golden.py`).  Integer outputs bit-exact; fp32 within 2e-5 relative."""
import numpy as np
import pytest
import torch

from conftest import pkg
from oracle import lotus_oracle as O


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("depth", [1, 3, 7, 9, 12, 16])
def test_codes_match_reference(units, depth):
    c = units["codes"][depth]
    for order in ("z", "z-trans", "hilbert", "hilbert-trans"):
        got = O.encode(c["grid"], c["batch"], depth, order)
        assert torch.equal(got, c[order]), order


@pytest.mark.parametrize("name", ["mixed", "single_small", "exact", "big"])
def test_patch_plan_matches_reference(units, name):
    p = units["patch_plan"][name]
    pad, unpad, cu = O.patch_plan(torch.tensor(np.cumsum(p["counts"])), 128)
    assert torch.equal(pad, p["pad"]) and torch.equal(unpad, p["unpad"])
    assert torch.equal(cu, p["cu_seqlens"]) and cu.dtype == torch.int32


def test_rpe_matches_reference(units):
    r = units["rpe"]
    assert r["pos_bnd"] == 15 and r["table"].shape[0] == 93
    assert torch.allclose(O.rpe_bias(r["table"], r["rel"], r["pos_bnd"]), r["bias"], atol=1e-7)


def test_fps_matches_reference(units):
    f = units["fps"]
    idx = O.farthest_point_sample_indices(f["points"].numpy(), f["npoint"], f["start"])
    assert torch.equal(f["points"][idx], f["sampled"])
    assert len(set(idx.tolist())) == f["npoint"]


def test_position_codec_matches_reference(units):
    syn = pkg("synthetic")
    p = units["pos_codec"]
    prob = syn.disc_gt_pos_prob(p["xyz"].numpy(), p["gt"].numpy(), 0.01, 15)
    assert np.allclose(prob, p["prob"].numpy(), rtol=1e-6, atol=0)
    # decode(encode(gt)) within one bin of gt; oracle decode == reference decode
    logits = torch.log(p["prob"].float().clamp_min(1e-30)).view(3, -1, 30)
    best = O.best_pos_from_logits(logits, p["xyz"], [p["xyz"].shape[0]], 15, 0.01)[0]
    assert torch.allclose(best.double(), p["best"].double(), atol=1e-6)
    assert (best - p["gt"]).abs().max() <= 0.01 + 1e-6


def test_rotation_codec_matches_reference(units):
    r = units["rot_codec"]
    q = O.euler_bins_to_quat(r["bins"].numpy(), 5)
    ref = r["quat"].numpy()
    # scipy may flip the overall sign of a quaternion; compare up to sign
    sign = np.sign((q * ref).sum(-1, keepdims=True))
    assert np.allclose(q * sign, ref, atol=1e-12)


def _integer_checks(inter, res_inter):
    pass


@pytest.mark.parametrize("which", ["small", "prod", "big"])
def test_policy_forward_backward_matches_reference(which, gold_small, gold_prod, request):
    # "big": two keysteps of 4096 / 3000 points at production width (BASELINE config-2 per-keystep sizes)
    g = gold_small if which == "small" else (gold_prod if which == "prod" else request.getfixturevalue("gold_big"))
    cfgm, testing = pkg("config"), pkg("testing")
    cfg = cfgm.ConfigNode(g["config"])
    # weights: re-derived exactly as the generator did (same seeded procedure over the same names/shapes)
    model = pkg("policy").build_parameter_shell(cfg)
    testing.deterministic_reinit(model, g["reinit_seed"])
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running_" not in k}
    sd.update(params)
    inter, bn_out = {}, {}
    res = O.policy_forward(sd, cfg, g["batch"], perms=g["perms"], train=True, bn_out=bn_out, inter=inter)
    for k in ("pos", "rot", "open", "total"):
        assert abs(res["losses"][k].item() - g["losses"][k].item()) <= 2e-5 * abs(g["losses"][k].item()), k
    for name, ref in g["inter"].items():
        got = inter[name]
        if which == "small":
            assert rel_err(got, ref["feat"]) < 2e-5, name
        else:
            assert got.shape == ref["shape"]
            assert rel_err(got[:8], ref["feat_head"]) < 2e-5, name
            assert abs(got.double().abs().sum().item() - ref["feat_abs_sum"]) < 2e-5 * ref["feat_abs_sum"], name
    res["losses"]["total"].backward()
    # some gradients are mathematically zero (k_norm.bias: softmax shift invariance; biases feeding a
    # BatchNorm): their fp32 values are rounding noise, so they are only required to stay below a floor
    floor = 1e-6 * max(g["grad_norms"].values())
    errs = []
    for k, n in g["grad_norms"].items():
        gn = params[k].grad.norm().item()
        if n < floor:                      # mathematically zero (bias in front of a BatchNorm / softmax)
            assert gn < floor, (k, gn, n)
        else:
            errs.append(abs(gn - n) / n)
    # Grid-pool max is non-smooth: one near-tie argmax flip (two candidates within fp32 rounding of each other;
    # observed once at production width: enc4.down, channel 141) reroutes one gradient element and moves
    # upstream gradient norms by ~1e-3 while every activation still agrees to 1e-6.  The reference itself
    # (fp32 vs fp64) is exposed to the same flips.  So: the typical gradient must agree to 1e-4 and no
    # gradient may be off by more than 2e-2; the small-width fixture (no flip) must agree to 1e-4 throughout.
    errs = sorted(errs)
    assert errs[len(errs) // 2] < 1e-5 and errs[-1] < (1e-4 if which == "small" else 2e-2), errs[-5:]
    for k, ref in g["grads"].items():
        if ref.norm().item() > floor:
            assert rel_err(params[k].grad, ref) < (1e-4 if which == "small" else 2e-2), k
    for k, ref in g["bn_after"].items():
        if "num_batches" in k:
            continue
        assert rel_err(bn_out[k], ref) < 1e-5, k
    # parameters the reference leaves without grad must be exactly the ones we leave without grad
    assert {k for k, p in params.items() if p.grad is not None} == set(g["grad_norms"].keys())


def test_policy_integer_structures_match_reference(gold_small):
    g = gold_small
    emb = g["inter"]["embedding"]
    b = g["batch"]
    batch_ids = O.offset2batch(b["offset"])
    grid = O.grid_coords(b["pc_fts"][:, :3])
    assert torch.equal(grid, emb["grid_coord"]) and torch.equal(batch_ids, emb["batch"])
    ser = O.serialize(grid, batch_ids, perm=g["perms"][0])
    assert ser["depth"] == emb["serialized_depth"]
    assert torch.equal(ser["code"], emb["serialized_code"])
    assert torch.equal(ser["order"], emb["serialized_order"])
    assert torch.equal(ser["inverse"], emb["serialized_inverse"])
    pt = dict(grid=grid, batch=batch_ids, offset=b["offset"], **ser)
    blk = g["inter"]["enc0.block0"]
    pad, unpad, cu = O.patch_plan(pt["offset"], 128)
    assert torch.equal(pad, blk["pad"]) and torch.equal(unpad, blk["unpad"]) and torch.equal(cu, blk["cu_seqlens_key"])
    for s in range(1, 5):
        st = O.pool_structure(pt, g["perms"][s])
        ref = g["inter"][f"enc{s}.down"]
        assert torch.equal(st["cluster"], ref["pooling_inverse"]), s
        assert torch.equal(st["code"], ref["serialized_code"]), s
        assert torch.equal(st["order"], ref["serialized_order"]), s
        assert torch.equal(st["inverse"], ref["serialized_inverse"]), s
        assert torch.equal(st["grid"], ref["grid_coord"]) and torch.equal(st["batch"], ref["batch"]), s
        assert torch.equal(st["offset"], ref["offset"]) and st["depth"] == ref["serialized_depth"], s
        pt = dict(grid=st["grid"], batch=st["batch"], offset=st["offset"], code=st["code"], order=st["order"],
                  inverse=st["inverse"], depth=st["depth"])


def test_final_action_eval_matches_reference(gold_small):
    g = gold_small
    cfgm, testing = pkg("config"), pkg("testing")
    cfg = cfgm.ConfigNode(g["config"])
    model = pkg("policy").build_parameter_shell(cfg)
    testing.deterministic_reinit(model, g["reinit_seed"])
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd.update({k: v for k, v in g["bn_after"].items()})          # eval pass ran after the train-mode forward
    with torch.no_grad():
        res = O.policy_forward(sd, cfg, g["batch"], perms=g["perms_eval"], train=False, compute_loss=False,
                               compute_final_action=True)
    ref = g["final_actions_eval"]
    assert ref.dtype == torch.float64 and ref.shape == (4, 8)
    assert torch.allclose(res["final"][:, :3].double(), ref[:, :3], atol=1e-6)
    q, qr = res["final"][:, 3:7].double(), ref[:, 3:7]
    sign = torch.sign((q * qr).sum(-1, keepdim=True))
    assert torch.allclose(q * sign, qr, atol=1e-9)
    assert rel_err(res["final"][:, 7], ref[:, 7]) < 2e-5


VARIANTS = ["adanorm", "concat", "ca_tokens", "planner_ca", "planner_adanorm"]


def _variant_of(kind):
    return {"adanorm": "adanorm", "concat": "concat", "ca_tokens": "ca", "planner_ca": "ca", "planner_adanorm": "adanorm"}[kind]


@pytest.mark.parametrize("kind", VARIANTS)
def test_variant_models_match_reference(kind, gold_variants):
    """§8a rows 15 (pose / step tokens), 19 (motion planner), 20 (PDNorm): oracle vs the reference-minted fixtures"""
    g = gold_variants[kind]
    cfg = pkg("config").ConfigNode(g["config"])
    model = pkg("policy").build_parameter_shell(cfg.clone())
    assert type(model).__name__ == g["model_class"]
    pkg("testing").deterministic_reinit(model, g["reinit_seed"])
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running_" not in k}
    sd.update(params)
    bn_out = {}
    fwd = O.planner_forward if kind.startswith("planner") else O.policy_forward
    cfg_run = pkg("policy").build_parameter_shell(cfg.clone()).config      # planner ctor widens in_channels; harmless here
    res = fwd(sd, cfg_run, g["batch"], perms=g["perms"], train=True, bn_out=bn_out, variant=_variant_of(kind))
    for k, ref in g["losses"].items():
        assert abs(res["losses"][k].item() - ref.item()) <= 2e-5 * abs(ref.item()), k
    f = res["dec_feat"]
    assert tuple(f.shape) == g["dec_feat_shape"]
    assert rel_err(f[:16], g["dec_feat_head"]) < 2e-5
    assert abs(f.double().abs().sum().item() - g["dec_feat_abs_sum"]) < 2e-5 * g["dec_feat_abs_sum"]
    assert rel_err(res["final"], g["train_actions"]) < 1e-6
    res["losses"]["total"].backward()
    floor = 1e-6 * max(g["grad_norms"].values())
    errs = []
    for k, n in g["grad_norms"].items():
        gn = params[k].grad.norm().item()
        if n < floor:
            assert gn < floor, (k, gn, n)
        else:
            errs.append((abs(gn - n) / n, k))
    errs.sort()
    assert errs[len(errs) // 2][0] < 1e-5 and errs[-1][0] < 1e-4, errs[-5:]
    for k, ref in g["grads"].items():
        if ref.norm().item() > floor:
            assert rel_err(params[k].grad, ref) < 1e-4, k
    for k, ref in g["bn_after"].items():
        assert rel_err(bn_out[k], ref) < 1e-5, k
    assert {k for k, p in params.items() if p.grad is None} == set(g["params_without_grad"])
    # eval pass: running statistics + final-action decode
    sd_eval = {k: v.detach() for k, v in sd.items()}
    sd_eval.update(g["bn_after"])                          # the reference's eval pass ran after its training step
    with torch.no_grad():
        ev = fwd(sd_eval, cfg_run, g["batch"], perms=g["perms_eval"], train=False,
                 compute_loss=False, compute_final_action=True, variant=_variant_of(kind))
    assert ev["final"].dtype == g["final_actions_eval"].dtype == torch.float64
    assert torch.allclose(ev["final"], g["final_actions_eval"], rtol=1e-5, atol=1e-6)


def test_soft_label_restatement_matches_reference_all_branches():
    """`synthetic.disc_gt_pos_prob` (the restatement the GPU producer test compares `lotus_disc_pos_prob` with) against vectors
    minted from the reference's own `get_disc_gt_pos_prob` (oracle/make_golden_labels.py): both heatmap types, the argmin one-hot
    fallback of an axis with no candidate within 1 cm, an exact hit (distance 0), a one-point cloud."""
    from conftest import load_golden
    syn = pkg("synthetic")
    cases = load_golden("labels.pt")
    assert len(cases) == 4
    for c in cases:
        xyz, gt = c["xyz"].numpy(), c["gt"].numpy()
        for kind in ("dist", "plain"):
            ref = c[kind].numpy()
            got = syn.disc_gt_pos_prob(xyz, gt, 0.01, 15, kind)
            assert got.shape == ref.shape == (3, xyz.shape[0] * 30)
            assert np.allclose(got, ref, rtol=1e-12, atol=0), kind
            assert np.allclose(ref.sum(-1), 1.0, atol=1e-6)
    # the degenerate case really exercises the fallback: axis 1 is a one-hot
    deg = cases[3]["dist"].numpy()
    assert (deg[1] > 0).sum() == 1 and deg[1].max() == 1.0


def test_adamw_restatement_matches_reference_optimizer():
    """`O.adamw_step` (the checker of the fused CUDA optimizer and of the gloo data-parallel test) against the reference's own
    AdamW + build_optimizer + clip_grad_norm_ (fixture: oracle/make_golden_optim.py, minted from train/optim/adamw.py:53-112,
    train/optim/misc.py:13-55, train_simple_policy.py:237-243).  Four steps, changing lr, gradients below / above the clip norm."""
    from conftest import load_golden
    g = load_golden("adamw.pt")
    names = [n for n, _ in g["shapes"]]
    # the reference's grouping: names containing 'bias' escape weight decay
    nodecay = set(g["groups"][1])
    assert nodecay == {i for i, n in enumerate(names) if "bias" in n} and g["group_wd"] == [0.05, 0.0]
    wds = [0.0 if i in nodecay else g["weight_decay"] for i in range(len(names))]
    params = [p.clone() for p in g["init"]]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    for step in range(1, 5):
        norm = O.adamw_step(params, [x.clone() for x in g["grads"][step - 1]], m, v, step, g["lr"][step - 1], wds, betas=g["betas"], eps=1e-6,
                            max_grad_norm=g["max_grad_norm"])
        assert abs(norm - g["norm"][step - 1]) <= 1e-6 * g["norm"][step - 1]
        for i, n in enumerate(names):
            # fp32 arithmetic on both sides; the clip coefficient is formed in a different order -> a few ulp
            assert rel_err(params[i], g["params"][step - 1][i]) < 2e-6, (step, n)
            assert rel_err(m[i], g["exp_avg"][step - 1][i]) < 2e-6, (step, n)
            assert rel_err(v[i], g["exp_avg_sq"][step - 1][i]) < 2e-6, (step, n)


def test_producer_restatement_matches_reference_stream():
    """batch-producer restatement (`O.augment_keystep`) against the fixture replayed around the reference's own `random_rotate_z`
    with numpy's global generator (oracle/make_golden_producer.py; simple_policy_dataset.py:158-181, 281-346): identical float32 features"""
    from conftest import load_golden
    for c in load_golden("producer.pt"):
        ft, cen, gt = O.augment_keystep(c["raw_xyz"].numpy(), c["raw_rgb"].numpy() if c["use_color"] else None, c["point_idxs"].numpy(), c["angle"],
                                        c["noise"].numpy(), c["gt_pos_in"].numpy(), c["table_height"])
        assert torch.equal(ft, c["pc_ft"]) and torch.equal(gt, c["gt_pos_out"])
        assert np.array_equal(cen, c["centroid"].numpy())
