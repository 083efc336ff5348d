"""Drop-in policy classes: same constructor / config / checkpoint / forward surface as
`genrobo3d.models.simple_policy_ptv3` (reference: models/simple_policy_ptv3.py:19-68,160-192,225-306,376-431;
models/base.py:10-49), with the arithmetic executed by the sm_100a CUDA engine behind the C ABI
(`include/lotus.h`).  There is NO CPU fallback: `forward` raises if the library or a CUDA device is missing.
"""
import numpy as np
import torch
import torch.nn as nn

from .modules import PointTransformerV3, PointTransformerV3CA


class BaseModel(nn.Module):
    """models/base.py:10-49"""

    @property
    def num_parameters(self):
        nweights, nparams = 0, 0
        for _, v in self.named_parameters():
            nweights += int(np.prod(v.size()))
            nparams += 1
        return nweights, nparams

    @property
    def num_trainable_parameters(self):
        nweights, nparams = 0, 0
        for _, v in self.named_parameters():
            if v.requires_grad:
                nweights += int(np.prod(v.size()))
                nparams += 1
        return nweights, nparams

    def prepare_batch(self, batch):
        device = next(self.parameters()).device
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                batch[k] = v.to(device, non_blocking=True)
        return batch

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)
        elif isinstance(m, nn.Embedding):
            nn.init.trunc_normal_(m.weight, std=0.02)


class RobotPoseEmbedding(nn.Module):
    """models/base.py:52-78 (parameter holder)."""

    def __init__(self, hidden_size):
        super().__init__()
        self.open_embedding = nn.Embedding(2, hidden_size)
        self.pos_embedding = nn.Linear(3, hidden_size)
        self.rot_embedding = nn.Linear(6, hidden_size)
        self.layer_norm = nn.LayerNorm(hidden_size, eps=1e-12)


class ActionHead(nn.Module):
    """models/simple_policy_ptv3.py:19-68 (parameter holder; same Sequential indices 0 and 3)."""

    def __init__(self, reduce, pos_pred_type, rot_pred_type, hidden_size, dim_actions, dropout=0, voxel_size=0.01,
                 euler_resolution=5, ptv3_config=None, pos_bins=50):
        super().__init__()
        assert reduce in ["max", "mean", "attn", "multiscale_max", "multiscale_max_large"]
        assert pos_pred_type in ["heatmap_mlp", "heatmap_mlp3", "heatmap_mlp_topk", "heatmap_mlp_clf",
                                 "heatmap_normmax", "heatmap_disc"]
        assert rot_pred_type in ["quat", "rot6d", "euler", "euler_delta", "euler_disc"]
        self.reduce, self.pos_pred_type, self.rot_pred_type = reduce, pos_pred_type, rot_pred_type
        self.hidden_size, self.dim_actions, self.voxel_size = hidden_size, dim_actions, voxel_size
        self.euler_resolution, self.euler_bins, self.pos_bins = euler_resolution, 360 // euler_resolution, pos_bins
        self.dropout = dropout
        out_hm = 3 * pos_bins * 2 if pos_pred_type == "heatmap_disc" else 1 + 3
        self.heatmap_mlp = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.LeakyReLU(0.02), nn.Dropout(dropout),
                                         nn.Linear(hidden_size, out_hm))
        out_act = self.euler_bins * 3 + 1 if rot_pred_type == "euler_disc" else dim_actions - 3
        if reduce == "attn":
            out_act += 1
        self.action_mlp = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.LeakyReLU(0.02), nn.Dropout(dropout),
                                        nn.Linear(hidden_size, out_act))


class SimplePolicyPTV3AdaNorm(BaseModel):
    """models/simple_policy_ptv3.py:160-192 (adaptive-norm conditioning; the yaml-default model)."""

    backbone_cls = PointTransformerV3
    variant = "adanorm"
    planner = False

    def __init__(self, config):
        super().__init__()
        if self.variant == "adanorm" or self.variant == "concat":
            config.defrost()
            config.ptv3_config.pdnorm_only_decoder = config.ptv3_config.get("pdnorm_only_decoder", False)
            config.freeze()
        self.config = config
        self.ptv3_model = self.backbone_cls(**config.ptv3_config)
        act_cfg = config.action_config
        self.txt_fc = nn.Linear(act_cfg.txt_ft_size, act_cfg.context_channels)
        if self.variant != "ca" and act_cfg.txt_reduce == "attn":
            self.txt_attn_fc = nn.Linear(act_cfg.txt_ft_size, 1)
        if act_cfg.use_ee_pose:
            self.pose_embedding = RobotPoseEmbedding(act_cfg.context_channels)
        if act_cfg.use_step_id:
            self.stepid_embedding = nn.Embedding(act_cfg.max_steps, act_cfg.context_channels)
        self.act_proj_head = ActionHead(
            act_cfg.reduce, act_cfg.pos_pred_type, act_cfg.rot_pred_type, config.ptv3_config.dec_channels[0],
            act_cfg.dim_actions, dropout=act_cfg.dropout, voxel_size=act_cfg.voxel_size,
            ptv3_config=config.ptv3_config, pos_bins=config.action_config.pos_bins)
        self.apply(self._init_weights)
        self._engine = None

    # ------------------------------------------------------------------ engine plumbing
    def engine(self):
        if self._engine is None:
            from .engine import PolicyEngine
            self._engine = PolicyEngine(self)
        return self._engine

    def forward(self, batch, compute_loss=False, **kwargs):
        """`forward(batch, compute_loss=False, compute_final_action=True)` -> final_pred_actions (B,8) float64
        [, {'pos','rot','open','total'}] exactly as simple_policy_ptv3.py:225-306."""
        batch = self.prepare_batch(batch)
        return self.engine().forward(batch, compute_loss=compute_loss, **kwargs)


class SimplePolicyPTV3CA(SimplePolicyPTV3AdaNorm):
    """models/simple_policy_ptv3.py:376-431 (cross-attention conditioning; the production model)."""
    backbone_cls = PointTransformerV3CA
    variant = "ca"


class SimplePolicyPTV3Concat(SimplePolicyPTV3AdaNorm):
    """models/simple_policy_ptv3.py:434-463"""
    variant = "concat"


class TrajActionHead(nn.Module):
    """models/motion_planner_ptv3.py:20-76 (parameter holder; heads read [point feature | trajectory-step embedding])."""

    def __init__(self, reduce, pos_pred_type, rot_pred_type, hidden_size, dim_actions, max_traj_len, dropout=0,
                 voxel_size=0.01, euler_resolution=5, ptv3_config=None, pos_bins=50, traj_embed_size=64):
        super().__init__()
        assert reduce in ["max", "mean", "attn"]
        assert pos_pred_type in ["heatmap_mlp", "heatmap_disc"]
        assert rot_pred_type in ["quat", "rot6d", "euler", "euler_disc"]
        self.reduce, self.pos_pred_type, self.rot_pred_type = reduce, pos_pred_type, rot_pred_type
        self.hidden_size, self.dim_actions, self.voxel_size = hidden_size, dim_actions, voxel_size
        self.euler_resolution, self.euler_bins, self.pos_bins = euler_resolution, 360 // euler_resolution, pos_bins
        self.max_traj_len, self.dropout = max_traj_len, dropout
        if traj_embed_size > 0:
            self.traj_embedding = nn.Embedding(max_traj_len, traj_embed_size)
        else:
            assert max_traj_len == 1
            self.traj_embedding = None
        out_hm = 3 * pos_bins * 2 if pos_pred_type == "heatmap_disc" else 1 + 3
        self.heatmap_mlp = nn.Sequential(nn.Linear(hidden_size + traj_embed_size, hidden_size), nn.LeakyReLU(0.02),
                                         nn.Dropout(dropout), nn.Linear(hidden_size, out_hm))
        out_act = self.euler_bins * 3 + 1 if rot_pred_type == "euler_disc" else dim_actions - 3
        if reduce == "attn":
            out_act += 1
        self.action_mlp = nn.Sequential(nn.Linear(hidden_size + traj_embed_size, hidden_size), nn.LeakyReLU(0.02),
                                        nn.Dropout(dropout), nn.Linear(hidden_size, out_act + 1))   # + stop logit


class MotionPlannerPTV3AdaNorm(BaseModel):
    """models/motion_planner_ptv3.py:151-199 (3D-LOTUS++ motion planner, adaptive-norm conditioning)."""

    backbone_cls = PointTransformerV3
    variant = "adanorm"
    planner = True

    def __init__(self, config):
        super().__init__()
        config.defrost()
        config.ptv3_config.pdnorm_only_decoder = config.ptv3_config.get("pdnorm_only_decoder", False)
        config.ptv3_config.in_channels += config.action_config.pc_label_channels
        config.freeze()
        self.config = config
        self.ptv3_model = self.backbone_cls(**config.ptv3_config)
        act_cfg = config.action_config
        self.pc_label_embedding = nn.Embedding(4, act_cfg.pc_label_channels)    # obstacle, robot, object, target
        self.txt_fc = nn.Linear(act_cfg.txt_ft_size, act_cfg.context_channels)
        if act_cfg.txt_reduce == "attn":
            self.txt_attn_fc = nn.Linear(act_cfg.txt_ft_size, 1)       # dead in the CA variant, kept for checkpoints
        if act_cfg.use_ee_pose:
            self.pose_embedding = RobotPoseEmbedding(act_cfg.context_channels)
        self.act_proj_head = TrajActionHead(
            act_cfg.reduce, act_cfg.pos_pred_type, act_cfg.rot_pred_type, config.ptv3_config.dec_channels[0],
            act_cfg.dim_actions, act_cfg.max_traj_len, dropout=act_cfg.dropout, voxel_size=act_cfg.voxel_size,
            ptv3_config=config.ptv3_config, pos_bins=act_cfg.pos_bins, traj_embed_size=act_cfg.traj_embed_size)
        self.apply(self._init_weights)
        self._engine = None

    def engine(self):
        if self._engine is None:
            from .engine import PolicyEngine
            self._engine = PolicyEngine(self)
        return self._engine

    def forward(self, batch, compute_loss=False, **kwargs):
        """-> final_pred_actions (B, T, 9) float64 [, {'pos','rot','open','stop','total'}] (motion_planner_ptv3.py:230-305)"""
        batch = self.prepare_batch(batch)
        return self.engine().forward(batch, compute_loss=compute_loss, **kwargs)


class MotionPlannerPTV3CA(MotionPlannerPTV3AdaNorm):
    """models/motion_planner_ptv3.py:400-463 (cross-attention conditioning; the shipped 3D-LOTUS++ planner)."""
    backbone_cls = PointTransformerV3CA
    variant = "ca"


MODEL_FACTORY = {
    "SimplePolicyPTV3AdaNorm": SimplePolicyPTV3AdaNorm,
    "SimplePolicyPTV3CA": SimplePolicyPTV3CA,
    "SimplePolicyPTV3Concat": SimplePolicyPTV3Concat,
    # train_motion_planner.py:46-49
    "MotionPlannerPTV3AdaNorm": MotionPlannerPTV3AdaNorm,
    "MotionPlannerPTV3CA": MotionPlannerPTV3CA,
}


def build_parameter_shell(config):
    """Instantiate `MODEL_FACTORY[config.model_class]` (train_simple_policy.py:46-50,113-114)."""
    return MODEL_FACTORY[config.get("model_class", "SimplePolicyPTV3CA")](config)
