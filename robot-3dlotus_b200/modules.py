"""Parameter tree of the PTv3 / PTv3-CA backbone — identical names, shapes, init and registration order to
the reference so `model_step_N.pt` round-trips with `strict=True`
(reference: genrobo3d/models/PointTransformerV3/model.py:329-398,586-657,683-711,793-815,831-857,864-1063;
model_ca.py:19-44,104-133,155-381).

These modules only HOLD parameters/buffers (as real nn.Linear / nn.LayerNorm / nn.BatchNorm1d, so
`_init_weights`, `convert_sync_batchnorm`, DDP, `clip_grad_norm_` and the optimizer's name-based grouping
behave as in the reference).  The arithmetic is done by the CUDA engine (`engine.py` -> C-ABI library).
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn


class SubMConv3d(nn.Module):
    """Parameter holder for a submanifold sparse conv; weight [Cout, k, k, k, Cin] (spconv>=2.2 KRSC),
    init = spconv's kaiming_uniform_(a=sqrt(5)) -> U(+-1/sqrt(Cin*k^3)) (not touched by BaseModel._init_weights)."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=True, indice_key=None):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.indice_key = indice_key
        k = kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, k, k, k, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        bound = 1.0 / math.sqrt(in_channels * k ** 3)
        nn.init.uniform_(self.weight, -bound, bound)
        if bias:
            nn.init.uniform_(self.bias, -bound, bound)


class Seq(nn.Module):
    """Named container (stands where the reference has PointSequential, model.py:188-227)."""

    def __init__(self, *mods, **named):
        super().__init__()
        if len(mods) == 1 and isinstance(mods[0], OrderedDict):
            for k, m in mods[0].items():
                self.add_module(k, m)
        else:
            for i, m in enumerate(mods):
                self.add_module(str(i), m)
        for k, m in named.items():
            self.add_module(k, m)

    def add(self, module, name=None):
        self.add_module(str(len(self._modules)) if name is None else name, module)

    def __getitem__(self, i):
        return list(self._modules.values())[i]

    def __len__(self):
        return len(self._modules)


class RPE(nn.Module):
    """model.py:307-326 (pos_bnd = int((4K)**(1/3)*2) -> 15 for K=128)."""

    def __init__(self, patch_size, num_heads):
        super().__init__()
        self.pos_bnd = int((4 * patch_size) ** (1 / 3) * 2)
        self.rpe_num = 2 * self.pos_bnd + 1
        self.rpe_table = nn.Parameter(torch.zeros(3 * self.rpe_num, num_heads))
        nn.init.trunc_normal_(self.rpe_table, std=0.02)


class PDNorm(nn.Module):
    """model.py:257-304 (decouple=False path used by the yaml default: one norm + SiLU->Linear modulation)."""

    def __init__(self, num_features, norm_layer, context_channels=256, conditions=None, decouple=True, adaptive=False):
        super().__init__()
        self.conditions, self.decouple, self.adaptive = conditions, decouple, adaptive
        if decouple:
            self.norm = nn.ModuleList([norm_layer(num_features) for _ in conditions])
        else:
            self.norm = norm_layer(num_features)
        if adaptive:
            self.modulation = nn.Sequential(nn.SiLU(), nn.Linear(context_channels, 2 * num_features, bias=True))


class SerializedAttention(nn.Module):
    def __init__(self, channels, num_heads, patch_size, qkv_bias=True, qk_scale=None, qk_norm=False,
                 attn_drop=0.0, proj_drop=0.0, order_index=0, enable_rpe=False, enable_flash=True):
        super().__init__()
        assert channels % num_heads == 0
        self.channels, self.num_heads, self.head_dim = channels, num_heads, channels // num_heads
        self.scale = qk_scale or self.head_dim ** -0.5
        self.patch_size, self.order_index = patch_size, order_index
        self.attn_drop, self.proj_drop_p = attn_drop, proj_drop
        self.enable_rpe, self.enable_flash, self.qk_norm = enable_rpe, enable_flash, qk_norm
        if enable_flash:
            assert enable_rpe is False, "Set enable_rpe to False when enable Flash Attention"
        self.qkv = nn.Linear(channels, channels * 3, bias=qkv_bias)
        self.proj = nn.Linear(channels, channels)
        self.rpe = RPE(patch_size, num_heads) if enable_rpe else None
        self.q_norm = nn.LayerNorm(self.head_dim, eps=1e-6) if qk_norm else nn.Identity()
        self.k_norm = nn.LayerNorm(self.head_dim, eps=1e-6) if qk_norm else nn.Identity()


class MLP(nn.Module):
    def __init__(self, in_channels, hidden_channels, out_channels):
        super().__init__()
        self.fc1 = nn.Linear(in_channels, hidden_channels)
        self.fc2 = nn.Linear(hidden_channels, out_channels)


class Block(nn.Module):
    def __init__(self, channels, num_heads, patch_size, mlp_ratio, qkv_bias, qk_scale, qk_norm, attn_drop, proj_drop,
                 drop_path, norm_layer, order_index, cpe_indice_key, enable_rpe, enable_flash, pre_norm=True):
        super().__init__()
        assert pre_norm, "post-norm blocks are not used by any shipped config"
        self.channels, self.drop_path_p, self.proj_drop = channels, drop_path, proj_drop
        self.cpe = Seq(SubMConv3d(channels, channels, 3, bias=True, indice_key=cpe_indice_key),
                       nn.Linear(channels, channels), norm_layer(channels))
        self.norm1 = Seq(norm_layer(channels))
        self.attn = SerializedAttention(channels, num_heads, patch_size, qkv_bias, qk_scale, qk_norm, attn_drop,
                                        proj_drop, order_index, enable_rpe, enable_flash)
        self.norm2 = Seq(norm_layer(channels))
        self.mlp = Seq(MLP(channels, int(channels * mlp_ratio), channels))


class CrossAttention(nn.Module):
    def __init__(self, channels, num_heads, kv_channels, attn_drop, proj_drop, qk_norm, enable_flash):
        super().__init__()
        self.channels, self.num_heads, self.head_dim = channels, num_heads, channels // num_heads
        self.scale = self.head_dim ** -0.5
        self.attn_drop, self.proj_drop_p, self.qk_norm, self.enable_flash = attn_drop, proj_drop, qk_norm, enable_flash
        self.q = nn.Linear(channels, channels, bias=True)
        self.kv = nn.Linear(kv_channels or channels, channels * 2, bias=True)
        self.proj = nn.Linear(channels, channels)
        self.q_norm = nn.LayerNorm(self.head_dim, eps=1e-6) if qk_norm else nn.Identity()
        self.k_norm = nn.LayerNorm(self.head_dim, eps=1e-6) if qk_norm else nn.Identity()


class CABlock(nn.Module):
    def __init__(self, channels, num_heads, kv_channels, attn_drop, proj_drop, mlp_ratio, norm_layer, qk_norm,
                 enable_flash, pre_norm=True):
        super().__init__()
        assert pre_norm
        self.channels, self.proj_drop = channels, proj_drop
        self.norm1 = Seq(norm_layer(channels))
        self.attn = CrossAttention(channels, num_heads, kv_channels, attn_drop, proj_drop, qk_norm, enable_flash)
        self.norm2 = Seq(norm_layer(channels))
        self.mlp = Seq(MLP(channels, int(channels * mlp_ratio), channels))


class SerializedPooling(nn.Module):
    def __init__(self, in_channels, out_channels, stride, norm_layer, shuffle_orders):
        super().__init__()
        assert stride == 2 ** (math.ceil(stride) - 1).bit_length()
        self.stride, self.shuffle_orders = stride, shuffle_orders
        self.proj = nn.Linear(in_channels, out_channels)
        self.norm = Seq(norm_layer(out_channels))


class SerializedUnpooling(nn.Module):
    def __init__(self, in_channels, skip_channels, out_channels, norm_layer):
        super().__init__()
        self.proj = Seq(nn.Linear(in_channels, out_channels), norm_layer(out_channels))
        self.proj_skip = Seq(nn.Linear(skip_channels, out_channels), norm_layer(out_channels))


class Embedding(nn.Module):
    def __init__(self, in_channels, embed_channels, norm_layer):
        super().__init__()
        self.stem = Seq(conv=SubMConv3d(in_channels, embed_channels, 5, bias=False, indice_key="stem"))
        self.stem.add(norm_layer(embed_channels), name="norm")


def _norm_factories(pdnorm_bn, pdnorm_ln, pdnorm_decouple, pdnorm_adaptive, pdnorm_context_channels,
                    pdnorm_affine, pdnorm_conditions):
    from functools import partial
    vanilla_bn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
    bn = vanilla_bn
    if pdnorm_bn:
        bn = partial(PDNorm, norm_layer=partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01, affine=pdnorm_affine),
                     conditions=pdnorm_conditions, decouple=pdnorm_decouple, adaptive=pdnorm_adaptive,
                     context_channels=pdnorm_context_channels)
    ln = nn.LayerNorm
    if pdnorm_ln:
        ln = partial(PDNorm, norm_layer=partial(nn.LayerNorm, elementwise_affine=pdnorm_affine),
                     conditions=pdnorm_conditions, decouple=pdnorm_decouple, adaptive=pdnorm_adaptive,
                     context_channels=pdnorm_context_channels)
    return bn, vanilla_bn, ln, nn.LayerNorm


class PointTransformerV3(nn.Module):
    """Constructor signature == reference `PointTransformerV3.__init__` (model.py:865-902) so
    `PointTransformerV3(**config.ptv3_config)` accepts every yaml key.  `with_ca` adds the CABlocks of
    `PointTransformerV3CA` (model_ca.py:155-381)."""

    with_ca = False

    def __init__(self, in_channels=6, order=("z", "z-trans", "hilbert", "hilbert-trans"), stride=(2, 2, 2, 2),
                 enc_depths=(2, 2, 2, 6, 2), enc_channels=(32, 64, 128, 256, 512), enc_num_head=(2, 4, 8, 16, 32),
                 enc_patch_size=(1024,) * 5, dec_depths=(2, 2, 2, 2), dec_channels=(64, 64, 128, 256),
                 dec_num_head=(4, 4, 8, 16), dec_patch_size=(1024,) * 4, mlp_ratio=4, ctx_channels=256,
                 qkv_bias=True, qk_scale=None, qk_norm=False, scaled_cosine_attn=False, attn_drop=0.0, proj_drop=0.0,
                 drop_path=0.3, pre_norm=True, shuffle_orders=True, enable_rpe=False, enable_flash=True,
                 upcast_attention=False, upcast_softmax=False, cls_mode=False, pdnorm_bn=False, pdnorm_ln=False,
                 pdnorm_decouple=True, pdnorm_adaptive=False, pdnorm_context_channels=256, pdnorm_affine=True,
                 pdnorm_conditions=("ScanNet", "S3DIS", "Structured3D"), pdnorm_only_decoder=False,
                 add_coords_in_attn="none"):
        super().__init__()
        assert not scaled_cosine_attn, "scaled_cosine_attn is marked TODO in the reference and unused"
        assert add_coords_in_attn in ("none", False), "add_coords_in_attn is unused by every shipped config"
        assert not cls_mode, "cls_mode is unused by the policies"
        self.num_stages = len(enc_depths)
        self.order = [order] if isinstance(order, str) else list(order)
        self.shuffle_orders = shuffle_orders
        self.stride = list(stride)
        self.enc_depths, self.dec_depths = list(enc_depths), list(dec_depths)
        self.enc_channels, self.dec_channels_cfg = list(enc_channels), list(dec_channels)
        self.enc_num_head, self.dec_num_head = list(enc_num_head), list(dec_num_head)
        self.enc_patch_size, self.dec_patch_size = list(enc_patch_size), list(dec_patch_size)
        self.qk_norm, self.enable_rpe, self.enable_flash = qk_norm, enable_rpe, enable_flash
        self.pdnorm_bn, self.pdnorm_ln, self.pdnorm_adaptive = pdnorm_bn, pdnorm_ln, pdnorm_adaptive
        self.pdnorm_only_decoder = pdnorm_only_decoder
        assert self.num_stages == len(stride) + 1 == len(enc_channels) == len(enc_num_head) == len(enc_patch_size)
        assert self.num_stages == len(dec_depths) + 1 == len(dec_channels) + 1 == len(dec_num_head) + 1
        assert not (pdnorm_bn or pdnorm_ln) or not pdnorm_decouple, "decoupled PDNorm is unused by the policies"

        bn, vanilla_bn, ln, vanilla_ln = _norm_factories(pdnorm_bn, pdnorm_ln, pdnorm_decouple, pdnorm_adaptive,
                                                         pdnorm_context_channels, pdnorm_affine, pdnorm_conditions)
        ca = self.with_ca
        only_dec = pdnorm_only_decoder
        self.embedding = Embedding(in_channels, enc_channels[0], bn if ca else (vanilla_bn if only_dec else bn))
        enc_dp = [x.item() for x in torch.linspace(0, drop_path, sum(enc_depths))]
        self.enc = Seq()
        for s in range(self.num_stages):
            dp = enc_dp[sum(enc_depths[:s]): sum(enc_depths[: s + 1])]
            enc = Seq()
            if s > 0:
                enc.add(SerializedPooling(enc_channels[s - 1], enc_channels[s], stride[s - 1],
                                          bn if ca else (vanilla_bn if only_dec else bn), shuffle_orders), name="down")
            for i in range(enc_depths[s]):
                blk_ln = ln if ca else (vanilla_ln if (only_dec and s < self.num_stages - 1) else ln)
                enc.add(Block(enc_channels[s], enc_num_head[s], enc_patch_size[s], mlp_ratio, qkv_bias, qk_scale,
                              qk_norm, attn_drop, proj_drop, dp[i], blk_ln, i % len(self.order), f"stage{s}",
                              enable_rpe, enable_flash, pre_norm), name=f"block{i}")
                if ca and ((not only_dec) or s == self.num_stages - 1):
                    enc.add(CABlock(enc_channels[s], enc_num_head[s], ctx_channels, attn_drop, proj_drop, mlp_ratio,
                                    ln, qk_norm, enable_flash, pre_norm), name=f"ca_block{i}")
            self.enc.add(enc, name=f"enc{s}")
        dec_dp = [x.item() for x in torch.linspace(0, drop_path, sum(dec_depths))]
        self.dec = Seq()
        dch = list(dec_channels) + [enc_channels[-1]]
        for s in reversed(range(self.num_stages - 1)):
            dp = dec_dp[sum(dec_depths[:s]): sum(dec_depths[: s + 1])]
            dp.reverse()
            dec = Seq()
            dec.add(SerializedUnpooling(dch[s + 1], enc_channels[s], dch[s], bn), name="up")
            for i in range(dec_depths[s]):
                dec.add(Block(dch[s], dec_num_head[s], dec_patch_size[s], mlp_ratio, qkv_bias, qk_scale, qk_norm,
                              attn_drop, proj_drop, dp[i], ln, i % len(self.order), f"stage{s}", enable_rpe,
                              enable_flash, pre_norm), name=f"block{i}")
                if ca:
                    dec.add(CABlock(dch[s], dec_num_head[s], ctx_channels, attn_drop, proj_drop, mlp_ratio, ln,
                                    qk_norm, enable_flash, pre_norm), name=f"ca_block{i}")
            self.dec.add(dec, name=f"dec{s}")
        self.dec_channels = dch


class PointTransformerV3CA(PointTransformerV3):
    with_ca = True
