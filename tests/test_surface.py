"""Host-side mirror of the reference interface: constructor/config/checkpoint surface."""
import numpy as np
import torch

from conftest import pkg


def test_production_parameter_census():
    cfg = pkg("config").production_model_config()
    m = pkg("policy").build_parameter_shell(cfg)
    assert m.num_parameters == (68177587, 421)
    sd = m.state_dict()
    assert len(sd) == 460
    assert tuple(sd["ptv3_model.embedding.stem.conv.weight"].shape) == (64, 5, 5, 5, 7)
    assert tuple(sd["ptv3_model.enc.enc4.block0.cpe.0.weight"].shape) == (768, 3, 3, 3, 768)
    assert tuple(sd["ptv3_model.enc.enc4.block0.attn.q_norm.weight"].shape) == (24,)
    assert tuple(sd["ptv3_model.dec.dec0.ca_block0.attn.kv.weight"].shape) == (256, 256)
    assert tuple(sd["act_proj_head.action_mlp.3.weight"].shape) == (217, 128)
    assert tuple(sd["act_proj_head.heatmap_mlp.3.weight"].shape) == (90, 128)
    # weight-decay grouping is name based ('bias' escapes decay): optim/misc.py:15,33-39
    no_decay = [n for n, _ in m.named_parameters() if "bias" in n]
    assert "ptv3_model.enc.enc0.block0.norm1.0.weight" not in no_decay and len(no_decay) > 150


def test_yaml_default_adanorm_model_constructs():
    cfg = pkg("config").yaml_default_model_config()
    m = pkg("policy").build_parameter_shell(cfg)
    assert len(m.state_dict()) == 661
    assert cfg.ptv3_config.pdnorm_only_decoder is False and cfg.is_frozen()


def test_config_node_contract():
    cfg = pkg("config").production_model_config()
    cfg.freeze()
    try:
        cfg.action_config.pos_bins = 3
        raise AssertionError("frozen node accepted a write")
    except AttributeError:
        pass
    cfg.defrost()
    cfg.action_config.pos_bins = 15
    assert cfg.action_config.get("best_disc_pos", "max") == "max" and dict(**cfg.ptv3_config)["in_channels"] == 7


def test_sync_batchnorm_conversion_and_modes():
    cfg = pkg("config").small_model_config()
    m = pkg("policy").build_parameter_shell(cfg)
    m2 = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m)
    assert len(m2.state_dict()) == 460
    m.eval(); assert not m.ptv3_model.embedding.stem.norm.training
    m.train(); assert m.ptv3_model.embedding.stem.norm.training
    bns = [k for k in m.state_dict() if k.endswith("running_mean")]
    assert len(bns) == 13                                            # 1 stem + 4 pool + 8 unpool


def test_synthetic_batch_schema():
    syn = pkg("synthetic")
    b = syn.make_batch(seed=1, sizes=[300, 150, 40])
    n = b["npoints_in_batch"]
    assert b["pc_fts"].shape == (sum(n), 7) and b["pc_fts"].dtype == torch.float32
    assert torch.equal(b["offset"], torch.tensor(np.cumsum(n))) and b["offset"].dtype == torch.int64
    assert b["txt_embeds"].shape == (sum(b["txt_lens"]), 512)
    assert b["gt_actions"].shape == (3, 7)
    for i, p in enumerate(b["disc_pos_probs"]):
        assert p.shape == (3, n[i] * 30) and torch.allclose(p.sum(-1), torch.ones(3), atol=1e-5)
    # unique voxels under the model's quantisation
    c = b["pc_fts"][:, :3]
    g = torch.div(c - c.min(0)[0], 0.01, rounding_mode="trunc").int()
    key = torch.cat([torch.repeat_interleave(torch.arange(3), torch.tensor(n))[:, None], g], 1)
    assert torch.unique(key, dim=0).shape[0] == key.shape[0]


def test_native_executor_refuses_stochastic_depth():
    """the C++ executor has no DropPath; the bare yaml default (drop_path 0.1 -> per-block rates linspace(0, .1, 5))
    must not be routed to it silently.  The shipped job scripts set drop_path 0.0 and stay supported."""
    cfgm, policy, native = pkg("config"), pkg("policy"), pkg("native")
    assert native.supported(policy.build_parameter_shell(cfgm.production_model_config()))
    assert not native.supported(policy.build_parameter_shell(cfgm.production_model_config(drop_path=0.1)))
