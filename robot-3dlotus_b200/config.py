"""Config surface the policy constructors read.

Mirrors what `genrobo3d.models.*` expects of `config.MODEL`
(reference: genrobo3d/configs/default.py:6-10 yacs CfgNode(new_allowed=True);
models/simple_policy_ptv3.py:166-168,384-397): attribute access, `.get`, `**`-expansion of
`ptv3_config`, `defrost()/freeze()`.  yacs is not installable offline, so this is a small
self-contained node with the same behaviour for those calls.
"""
import copy


class ConfigNode(dict):
    def __init__(self, init=None, **kw):
        super().__init__()
        object.__setattr__(self, "_frozen", False)
        d = dict(init or {})
        d.update(kw)
        for k, v in d.items():
            self[k] = ConfigNode(v) if isinstance(v, dict) and not isinstance(v, ConfigNode) else v

    def __getattr__(self, k):
        if k in self:
            return self[k]
        raise AttributeError(k)

    def __setattr__(self, k, v):
        if object.__getattribute__(self, "_frozen"):
            raise AttributeError(f"Attempted to set {k} to {v}, but ConfigNode is immutable")
        self[k] = ConfigNode(v) if isinstance(v, dict) and not isinstance(v, ConfigNode) else v

    def _set_frozen(self, f):
        object.__setattr__(self, "_frozen", f)
        for v in self.values():
            if isinstance(v, ConfigNode):
                v._set_frozen(f)

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def is_frozen(self):
        return object.__getattribute__(self, "_frozen")

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = ConfigNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        return out


def _ptv3_defaults():
    # genrobo3d/configs/rlbench/simple_policy_ptv3.yaml:95-128
    return dict(
        in_channels=6, order=["z", "z-trans", "hilbert", "hilbert-trans"], stride=[2, 2, 2, 2],
        enc_depths=[2, 2, 2, 6, 2], enc_channels=[32, 64, 128, 256, 512], enc_num_head=[2, 4, 8, 16, 32],
        enc_patch_size=[128] * 5, dec_depths=[2, 2, 2, 2], dec_channels=[64, 64, 128, 256],
        dec_num_head=[4, 4, 8, 16], dec_patch_size=[128] * 4, mlp_ratio=4, qkv_bias=True, qk_scale=None,
        qk_norm=False, scaled_cosine_attn=False, attn_drop=0.1, proj_drop=0.1, drop_path=0.1, pre_norm=True,
        shuffle_orders=True, enable_rpe=False, enable_flash=True, upcast_attention=False, upcast_softmax=False,
        cls_mode=False, pdnorm_bn=True, pdnorm_ln=True, pdnorm_decouple=False, pdnorm_adaptive=True,
        pdnorm_affine=True, pdnorm_conditions=None, pdnorm_only_decoder=False, add_coords_in_attn="none",
    )


def _action_defaults():
    # genrobo3d/configs/rlbench/simple_policy_ptv3.yaml:130-154
    return dict(
        voxel_size=0.01, context_channels=256, txt_ft_size=512, max_txt_len=77, txt_reduce="mean",
        use_ee_pose=True, use_step_id=False, max_steps=30, reduce="max", max_traj_len=1, dim_actions=8,
        pos_pred_type="heatmap_mlp", pos_heatmap_temp=0.1, rot_pred_type="euler", dropout=0.1,
        pos_bins=40, pos_bin_size=0.01, best_disc_pos="max",
    )


def yaml_default_model_config():
    """What `get_config(simple_policy_ptv3.yaml).MODEL` builds (SimplePolicyPTV3AdaNorm)."""
    return ConfigNode(model_class="SimplePolicyPTV3AdaNorm", ptv3_config=_ptv3_defaults(),
                      action_config=_action_defaults(), loss_config=dict(pos_weight=1, rot_weight=1))


def production_model_config(**ptv3_overrides):
    """`config.MODEL` of the published 3D-LOTUS checkpoints:
    job_scripts/train_3dlotus_policy.sh:39-87 applied over simple_policy_ptv3.yaml."""
    p = _ptv3_defaults()
    p.update(
        drop_path=0.0, attn_drop=0.1, proj_drop=0.1, in_channels=7, pdnorm_only_decoder=False, qk_norm=True,
        scaled_cosine_attn=False, enable_flash=True, enc_depths=[1, 1, 1, 1, 1], dec_depths=[1, 1, 1, 1],
        enc_channels=[64, 128, 256, 512, 768], dec_channels=[128, 128, 256, 512],
        pdnorm_bn=False, pdnorm_ln=False, pdnorm_adaptive=False,
    )
    p.update(ptv3_overrides)
    a = _action_defaults()
    a.update(dropout=0.2, voxel_size=0.01, reduce="max", dim_actions=7, rot_pred_type="euler_disc",
             pos_heatmap_temp=0.1, max_steps=30, use_step_id=False, use_ee_pose=False,
             pos_pred_type="heatmap_disc", pos_bins=15)
    return ConfigNode(model_class="SimplePolicyPTV3CA", ptv3_config=p, action_config=a,
                      loss_config=dict(pos_weight=1, rot_weight=1))


def production_planner_config(**ptv3_overrides):
    """`config.MODEL` of the 3D-LOTUS++ motion planner: job_scripts/train_3dlotusplus_motion_planner.sh:76-110 applied over
    configs/rlbench/motion_planner_ptv3.yaml (pc_label_channels 64, traj_embed_size 64, max_traj_len 5; xyz + height inputs)."""
    cfg = production_model_config(in_channels=4, **ptv3_overrides)
    cfg.model_class = "MotionPlannerPTV3CA"
    cfg.action_config.update(max_traj_len=5, pc_label_channels=64, traj_embed_size=64, txt_reduce="attn", use_ee_pose=False)
    return cfg


def small_model_config(**ptv3_overrides):
    """Production topology at reduced width — used for committed golden fixtures (weights fit in git)."""
    return production_model_config(
        **{**dict(enc_channels=[32, 32, 48, 64, 96], enc_num_head=[2, 1, 2, 4, 4],
                  dec_channels=[32, 32, 48, 64], dec_num_head=[2, 1, 2, 4]), **ptv3_overrides})


def small_variant_config(kind):
    """Reduced-width configurations of the non-production variants (golden fixtures / parity tests):
    'adanorm'  SimplePolicyPTV3AdaNorm — PDNorm(BN)+PDNorm(LN) adaptive, txt_reduce 'attn', pose + step-id added, depth-2 stages (the reference indexes decoder outputs assuming dec_depths all 1 or all 2: simple_policy_ptv3.py:243)
    'concat'   SimplePolicyPTV3Concat  — per-cloud context concatenated to every point feature (plain norms)
    'ca_tokens' SimplePolicyPTV3CA     — production topology with the pose and step-id context tokens enabled
    'planner_ca' / 'planner_adanorm'   — MotionPlannerPTV3{CA,AdaNorm} (motion_planner_ptv3.yaml + train_3dlotusplus_motion_planner.sh)"""
    cfg = small_model_config(attn_drop=0.0, proj_drop=0.0)
    cfg.action_config.dropout = 0.0
    p, a = cfg.ptv3_config, cfg.action_config
    if kind in ("adanorm", "planner_adanorm"):
        p.update(pdnorm_bn=True, pdnorm_ln=True, pdnorm_decouple=False, pdnorm_adaptive=True, pdnorm_affine=True,
                 qk_norm=False, enc_depths=[2, 1, 1, 2, 1], dec_depths=[2, 2, 2, 2])
        a.update(txt_reduce="attn", context_channels=64)
        p.update(pdnorm_context_channels=64)
    if kind == "adanorm":
        cfg.model_class = "SimplePolicyPTV3AdaNorm"
        a.update(use_ee_pose=True, use_step_id=True)
    elif kind == "concat":
        cfg.model_class = "SimplePolicyPTV3Concat"
        a.update(context_channels=9, use_ee_pose=True, use_step_id=True, txt_reduce="mean")
        p.update(in_channels=7 + 9)
    elif kind == "ca_tokens":
        a.update(use_ee_pose=True, use_step_id=True)
    elif kind in ("planner_ca", "planner_adanorm"):
        cfg.model_class = "MotionPlannerPTV3CA" if kind == "planner_ca" else "MotionPlannerPTV3AdaNorm"
        p.update(in_channels=4)
        a.update(max_traj_len=5, pc_label_channels=12, traj_embed_size=16, txt_reduce="attn",
                 use_ee_pose=(kind == "planner_adanorm"))
    else:
        raise KeyError(kind)
    return cfg
