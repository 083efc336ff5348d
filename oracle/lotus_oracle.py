"""CPU oracle: a restatement of the 3D-LOTUS policy hot path.

TEST INFRASTRUCTURE.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this file; the product (`robot-3dlotus_b200/`) never does and fails
loudly when its CUDA library is missing.

Parity status: PINNED.  The reference ships no golden vectors, so this restatement is pinned
against outputs of the reference's own, unmodified Python code imported in the build container
(`oracle/ref_import.py` + `oracle/make_golden.py` -> `tests/golden/*.pt`); `tests/test_oracle_golden.py`
checks every function below against those fixtures (bit-exact for integer outputs, 1e-5 for fp32).

Integer work is numpy/torch int64; float work is torch fp32 on CPU (autograd supplies the backward).
Attention follows the reference's *flash* branch semantics in fp32 (SURVEY finding 5 / Appendix D.1).
All paths in the citations are relative to /root/reference/genrobo3d/.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# §8a rows 1-2: batch ids, grid coordinates, space-filling-curve codes, orders
# --------------------------------------------------------------------------------------------------


def offset2bincount(offset):
    """models/PointTransformerV3/model.py:30-34"""
    return torch.diff(offset, prepend=offset.new_zeros(1))


def offset2batch(offset):
    """models/PointTransformerV3/model.py:37-42"""
    cnt = offset2bincount(offset)
    return torch.repeat_interleave(torch.arange(len(cnt), dtype=torch.long), cnt)


def grid_coords(coord, grid_size=0.01):
    """model.py:96-98 — min over the WHOLE packed batch, fp32 divide, truncate."""
    return torch.div(coord - coord.min(0)[0], grid_size, rounding_mode="trunc").int()


def serialized_depth(grid):
    """model.py:102"""
    return int(grid.max()).bit_length()


def z_order_code(x, y, z, depth):
    """serialization/z_order.py:40-50,66-101 — bit i of x -> 3i+2, y -> 3i+1, z -> 3i.
    (The reference uses 8-bit LUTs; the result is the plain bit interleave.)"""
    x, y, z = (np.asarray(v).astype(np.int64) for v in (x, y, z))
    key = np.zeros_like(x)
    for i in range(depth):
        key |= ((x >> i) & 1) << (3 * i + 2)
        key |= ((y >> i) & 1) << (3 * i + 1)
        key |= ((z >> i) & 1) << (3 * i)
    return key


def hilbert_code(x, y, z, depth):
    """serialization/hilbert.py:91-198 — Skilling's transform, MSB-first plane loop, then the bits of
    (X0,X1,X2) interleaved MSB-first and Gray->binary over the whole 3*depth-bit string."""
    X = [np.asarray(v).astype(np.int64).copy() for v in (x, y, z)]
    for bit in range(depth):                      # bit 0 == most significant plane
        q = np.int64(1) << (depth - 1 - bit)
        low = q - 1                               # planes below this one
        for d in range(3):
            on = (X[d] & q) != 0
            # plane set: invert the low planes of axis 0; plane clear: swap low planes of axis 0 and d
            t = np.where(on, 0, (X[0] ^ X[d]) & low)
            X[0] = np.where(on, X[0] ^ low, X[0] ^ t)
            X[d] = X[d] ^ t if d != 0 else X[0]
    g = np.zeros_like(X[0])
    for i in range(depth):                        # interleave: plane i of X0,X1,X2 -> bits 3i+2,3i+1,3i
        g |= ((X[0] >> i) & 1) << (3 * i + 2)
        g |= ((X[1] >> i) & 1) << (3 * i + 1)
        g |= ((X[2] >> i) & 1) << (3 * i)
    h = g.copy()                                  # Gray -> binary (prefix xor from the MSB)
    s = 1
    while s < 3 * depth:
        h ^= h >> s
        s <<= 1
    return h


def encode(grid, batch, depth, order):
    """serialization/default.py:10-27"""
    g = grid.numpy().astype(np.int64)
    if order in ("z-trans", "hilbert-trans"):
        g = g[:, [1, 0, 2]]
    fn = z_order_code if order.startswith("z") else hilbert_code
    code = fn(g[:, 0], g[:, 1], g[:, 2], depth)
    code = (batch.numpy().astype(np.int64) << (3 * depth)) | code
    return torch.from_numpy(code)


def order_and_inverse(code):
    """model.py:121-128"""
    order = torch.argsort(code, dim=1, stable=True)
    inverse = torch.zeros_like(order)
    ar = torch.arange(code.shape[1]).repeat(code.shape[0], 1)
    inverse.scatter_(1, order, ar)
    return order, inverse


def serialize(grid, batch, orders=("z", "z-trans", "hilbert", "hilbert-trans"), perm=None):
    """model.py:83-138; `perm` is the recorded `torch.randperm(len(orders))` of shuffle_orders."""
    depth = serialized_depth(grid)
    assert depth * 3 + int(batch.max() + 1).bit_length() <= 63 and depth <= 16
    code = torch.stack([encode(grid, batch, depth, o) for o in orders])
    order, inverse = order_and_inverse(code)
    if perm is not None:
        code, order, inverse = code[perm], order[perm], inverse[perm]
    return dict(code=code, order=order, inverse=inverse, depth=depth)


# --------------------------------------------------------------------------------------------------
# §8a row 6: patch plan
# --------------------------------------------------------------------------------------------------


def patch_plan(offset, K=128):
    """model.py:410-466.  pad: padded slot -> sorted position; unpad: sorted position ->
    padded slot; cu_seqlens: int32 sequence starts + total."""
    counts = offset2bincount(offset).tolist()
    pad, unpad, cu = [], [], []
    start, start_pad = 0, 0
    for c in counts:
        if c > K:
            c_pad = (c + K - 1) // K * K
        else:
            c_pad = c
        p = np.arange(c_pad, dtype=np.int64)
        if c_pad != c:
            r = c % K
            # the tail of the last patch re-uses the same slots of the previous patch
            p[c_pad - K + r:] = p[c_pad - 2 * K + r: c_pad - K]
        pad.append(p + start)
        unpad.append(np.arange(c, dtype=np.int64) + start_pad)
        cu.append(np.arange(start_pad, start_pad + c_pad, K, dtype=np.int32) if c_pad > 0 else np.zeros(0, np.int32))
        start += c
        start_pad += c_pad
    cu.append(np.array([start_pad], dtype=np.int32))
    return (torch.from_numpy(np.concatenate(pad)), torch.from_numpy(np.concatenate(unpad)),
            torch.from_numpy(np.concatenate(cu)))


# --------------------------------------------------------------------------------------------------
# §8a rows 3-4: sparsify + submanifold conv (third-party spconv 2.3.6 semantics, restated)
# --------------------------------------------------------------------------------------------------


def neighbor_table(batch, grid, k):
    """nbr[p, d] = index of the occupied voxel at grid[p] + (d0,d1,d2) - k//2 of the same cloud, else -1;
    d = (d0*k + d1)*k + d2 with d0 along x (column order of the spconv indices, model.py:169-171)."""
    r = k // 2
    b = batch.long()
    c = grid.long() + r
    dims = c.max(0).values + r + 1

    def lin(cc):
        return ((b * dims[0] + cc[:, 0]) * dims[1] + cc[:, 1]) * dims[2] + cc[:, 2]

    skey, sidx = torch.sort(lin(c))
    n = c.shape[0]
    nbr = torch.full((n, k ** 3), -1, dtype=torch.long)
    d = 0
    for d0 in range(k):
        for d1 in range(k):
            for d2 in range(k):
                qk = lin(c + torch.tensor([d0 - r, d1 - r, d2 - r]))
                pos = torch.searchsorted(skey, qk).clamp(max=n - 1)
                nbr[:, d] = torch.where(skey[pos] == qk, sidx[pos], torch.full_like(pos, -1))
                d += 1
    return nbr


def subm_conv(feat, nbr, weight, bias=None):
    """out[p] = sum_d W[:, d, :] @ feat[nbr[p, d]]  (+ bias); weight [Cout, k, k, k, Cin] (model.py:616-622, 845-852)."""
    cout, cin = weight.shape[0], weight.shape[-1]
    w = weight.reshape(cout, -1, cin)
    out = feat.new_zeros(feat.shape[0], cout)
    for d in range(w.shape[1]):
        q = nbr[:, d]
        m = (q >= 0).nonzero()[:, 0]
        if m.numel():
            out = out.index_add(0, m, feat[q[m]] @ w[:, d].t())
    return out if bias is None else out + bias


# --------------------------------------------------------------------------------------------------
# norms / activations
# --------------------------------------------------------------------------------------------------


def batch_norm(x, w, b, rm, rv, train, eps=1e-3, momentum=0.01):
    """nn.BatchNorm1d(eps=1e-3, momentum=0.01) (model.py:932).  Returns y and the updated (rm, rv)."""
    if train:
        mean = x.mean(0)
        var = x.var(0, unbiased=False)
        n = x.shape[0]
        new_rm = (1 - momentum) * rm + momentum * mean.detach()
        new_rv = (1 - momentum) * rv + momentum * var.detach() * n / max(n - 1, 1)
    else:
        mean, var, new_rm, new_rv = rm, rv, rm, rv
    return (x - mean) / torch.sqrt(var + eps) * w + b, new_rm, new_rv


def layer_norm(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


# --------------------------------------------------------------------------------------------------
# §8a rows 7-9: patch self-attention, RPE, cross-attention
# --------------------------------------------------------------------------------------------------


def _varlen_attention(q, k, v, cu_q, cu_k, scale, bias=None):
    """softmax(q k^T * scale [+ bias]) v per (cu_q[i]:cu_q[i+1]) x (cu_k[i]:cu_k[i+1]) segment; q (Nq,H,D)."""
    out = torch.empty_like(q)
    lens = {}
    for i in range(len(cu_q) - 1):
        lens.setdefault((cu_q[i + 1] - cu_q[i], cu_k[i + 1] - cu_k[i]), []).append(i)
    pieces, index = [], []
    for (lq, lk), ids in lens.items():
        if lq == 0:
            continue
        iq = torch.tensor([cu_q[i] for i in ids])[:, None] + torch.arange(lq)[None]
        ik = torch.tensor([cu_k[i] for i in ids])[:, None] + torch.arange(lk)[None]
        s = torch.einsum("bqhd,bkhd->bhqk", q[iq], k[ik]) * scale
        if bias is not None:
            s = s + bias[torch.tensor(ids)]
        p = torch.softmax(s, dim=-1)
        pieces.append(torch.einsum("bhqk,bkhd->bqhd", p, v[ik]).reshape(-1, q.shape[1], q.shape[2]))
        index.append(iq.reshape(-1))
    index = torch.cat(index)
    inv = torch.empty_like(index)
    inv[index] = torch.arange(index.numel())
    return torch.cat(pieces)[inv]


def rpe_bias(table, rel, pos_bnd):
    """model.py:307-326: bias[n,h,i,j] = sum_axis table[clamp(rel)+bnd+axis*(2bnd+1), h]."""
    num = 2 * pos_bnd + 1
    idx = rel.clamp(-pos_bnd, pos_bnd) + pos_bnd + torch.arange(3) * num
    return table[idx.reshape(-1)].view(idx.shape + (-1,)).sum(3).permute(0, 3, 1, 2)


def patch_attention(x, P, pre, H, order, inverse, plan, K=128, qk_norm=True, rpe_table=None, grid=None):
    """SerializedAttention.forward, flash branch semantics in fp32 (model.py:468-557).
    x is the *normed* input; returns proj(attn) (dropout 0)."""
    pad, unpad, cu = plan
    o = order[pad]
    inv = unpad[inverse]
    if P.get(pre + "qkv.weight") is None:      # kernel-level tests: `x` already holds the (N, 3C) projections
        C = x.shape[1] // 3
        D = C // H
        qkv = x[o].reshape(-1, 3, H, D)
    else:
        C = x.shape[1]
        D = C // H
        qkv = F.linear(x, P[pre + "qkv.weight"], P.get(pre + "qkv.bias"))[o].reshape(-1, 3, H, D)
    q, k, v = qkv.unbind(1)
    if qk_norm:
        q = layer_norm(q, P[pre + "q_norm.weight"], P[pre + "q_norm.bias"], 1e-6)
        k = layer_norm(k, P[pre + "k_norm.weight"], P[pre + "k_norm.bias"], 1e-6)
    cul = cu.tolist()
    bias = None
    if rpe_table is not None:
        # only defined when every sequence is a full K patch (model.py:400-408 reshape(-1, K, 3))
        g = grid[o].reshape(-1, K, 3).long()
        pos_bnd = int((4 * K) ** (1 / 3) * 2)
        bias = rpe_bias(rpe_table, g.unsqueeze(2) - g.unsqueeze(1), pos_bnd)
    a = _varlen_attention(q, k, v, cul, cul, D ** -0.5, bias)
    a = a.reshape(-1, C)[inv]
    return F.linear(a, P[pre + "proj.weight"], P[pre + "proj.bias"])


def cross_attention(x, ctx, P, pre, H, offset, ctx_offset, qk_norm=True):
    """CrossAttention.forward (model_ca.py:46-101), flash branch semantics in fp32."""
    C = x.shape[1]
    D = C // H
    if P.get(pre + "q.weight") is None:        # kernel-level tests: `x` / `ctx` already hold the q / kv projections
        q, kv = x.view(-1, H, D), ctx.view(-1, 2, H, D)
    else:
        q = F.linear(x, P[pre + "q.weight"], P[pre + "q.bias"]).view(-1, H, D)
        kv = F.linear(ctx, P[pre + "kv.weight"], P[pre + "kv.bias"]).view(-1, 2, H, D)
    k, v = kv[:, 0], kv[:, 1]
    if qk_norm:
        q = layer_norm(q, P[pre + "q_norm.weight"], P[pre + "q_norm.bias"], 1e-6)
        k = layer_norm(k, P[pre + "k_norm.weight"], P[pre + "k_norm.bias"], 1e-6)
    cq = [0] + offset.tolist()
    ck = [0] + ctx_offset.tolist()
    a = _varlen_attention(q, k, v, cq, ck, D ** -0.5).reshape(-1, C)
    return F.linear(a, P[pre + "proj.weight"], P[pre + "proj.bias"])


def mlp(x, P, pre):
    """model.py:560-583 (dropout 0)"""
    h = F.gelu(F.linear(x, P[pre + "fc1.weight"], P[pre + "fc1.bias"]))
    return F.linear(h, P[pre + "fc2.weight"], P[pre + "fc2.bias"])


def _modulate(pt, P, name, y):
    """adaptive half of PDNorm.forward (model.py:293-303): per-cloud [shift|scale] = Linear(SiLU(context))"""
    if name + ".modulation.1.weight" not in P:
        return y
    ss = F.linear(F.silu(pt["context"]), P[name + ".modulation.1.weight"], P[name + ".modulation.1.bias"])
    ss = torch.repeat_interleave(ss, offset2bincount(pt["offset"]), dim=0)
    shift, scale = ss.chunk(2, dim=1)
    return y * (1.0 + scale) + shift


def _ln(pt, P, name, x):
    """plain LayerNorm or PDNorm(LayerNorm) (decouple=False), chosen by the parameter names present"""
    if name + ".norm.weight" in P:
        return _modulate(pt, P, name, layer_norm(x, P[name + ".norm.weight"], P[name + ".norm.bias"]))
    return layer_norm(x, P[name + ".weight"], P[name + ".bias"])


def block(pt, P, pre, H, order_index, K=128, qk_norm=True):
    """Block.forward (model.py:659-680), pre-norm, drop_path 0.

    The CPE conv reads `point.sparse_conv_feat.features` (model.py:235-238), which after a
    SerializedUnpooling is the *skip branch only*: `parent.feat = parent.feat + point.feat[inverse]`
    (model.py:824) updates `feat` but not `sparse_conv_feat`.  `pt["conv_feat"]` carries that tensor."""
    x = pt["feat"]
    c = subm_conv(pt.pop("conv_feat", x), pt["nbr3"], P[pre + "cpe.0.weight"], P[pre + "cpe.0.bias"])
    c = F.linear(c, P[pre + "cpe.1.weight"], P[pre + "cpe.1.bias"])
    x = x + _ln(pt, P, pre + "cpe.2", c)
    h = _ln(pt, P, pre + "norm1.0", x)
    x = x + patch_attention(h, P, pre + "attn.", H, pt["order"][order_index], pt["inverse"][order_index],
                            pt["plan"], K, qk_norm)
    h = _ln(pt, P, pre + "norm2.0", x)
    x = x + mlp(h, P, pre + "mlp.0.")
    pt["feat"] = x
    return pt


def ca_block(pt, P, pre, H, qk_norm=True):
    """CABlock.forward (model_ca.py:135-152)."""
    x = pt["feat"]
    h = layer_norm(x, P[pre + "norm1.0.weight"], P[pre + "norm1.0.bias"])
    x = x + cross_attention(h, pt["context"], P, pre + "attn.", H, pt["offset"], pt["context_offset"], qk_norm)
    h = layer_norm(x, P[pre + "norm2.0.weight"], P[pre + "norm2.0.bias"])
    x = x + mlp(h, P, pre + "mlp.0.")
    pt["feat"] = x
    return pt


# --------------------------------------------------------------------------------------------------
# §8a rows 10-11: grid pooling / unpooling
# --------------------------------------------------------------------------------------------------


def pool_structure(pt, perm=None, stride=2):
    """The integer half of SerializedPooling.forward (model.py:713-755, 766-771)."""
    pooling_depth = (math.ceil(stride) - 1).bit_length()
    if pooling_depth > pt["depth"]:
        pooling_depth = 0
    code = pt["code"] >> (pooling_depth * 3)
    _, cluster, counts = torch.unique(code[0], sorted=True, return_inverse=True, return_counts=True)
    indices = torch.sort(cluster, stable=True)[1]
    idx_ptr = torch.cat([counts.new_zeros(1), torch.cumsum(counts, 0)])
    head = indices[idx_ptr[:-1]]
    code = code[:, head]
    order, inverse = order_and_inverse(code)
    if perm is not None:
        code, order, inverse = code[perm], order[perm], inverse[perm]
    batch = pt["batch"][head]
    return dict(cluster=cluster, indices=indices, idx_ptr=idx_ptr, head=head, code=code, order=order,
                inverse=inverse, depth=pt["depth"] - pooling_depth, grid=pt["grid"][head] >> pooling_depth,
                batch=batch, offset=torch.cumsum(torch.bincount(batch), 0))


def segment_max(x, idx_ptr):
    return torch.segment_reduce(x, "max", offsets=idx_ptr, axis=0)


def segment_mean(x, idx_ptr):
    return torch.segment_reduce(x, "mean", offsets=idx_ptr, axis=0)


# --------------------------------------------------------------------------------------------------
# §8a rows 14-18: backbone, context, head, loss, decode
# --------------------------------------------------------------------------------------------------


def _stage_aux(pt, K):
    pt["nbr3"] = neighbor_table(pt["batch"], pt["grid"], 3)
    pt["plan"] = patch_plan(pt["offset"], K)
    return pt


def ptv3ca_forward(P, cfg, feat, coord, offset, context, context_offset, perms, train, bn_out=None, inter=None):
    """PointTransformerV3CA.forward(return_dec_layers=True) (model_ca.py:383-412).
    P: state dict with the `ptv3_model.` prefix stripped.  perms: list of 5 recorded randperm(4) (or None)."""
    pre = ""
    S = len(cfg["enc_channels"])
    qk_norm = cfg["qk_norm"]
    assert not (cfg.get("pdnorm_bn") or cfg.get("pdnorm_ln")) or not cfg.get("pdnorm_decouple", True), \
        "decoupled PDNorm is unused by the policies"
    batch = offset2batch(offset)
    grid = grid_coords(coord, 0.01)
    ser = serialize(grid, batch, cfg["order"], perms[0] if perms is not None else None)
    pt = dict(feat=feat, coord=coord, grid=grid, batch=batch, offset=offset, context=context,
              context_offset=context_offset, **ser)
    K = cfg["enc_patch_size"][0]
    _stage_aux(pt, K)

    def bn(x, name, p=None):
        mod = name
        if name + ".norm.weight" in P:                      # PDNorm(BatchNorm1d) of the AdaNorm / Concat variants
            name = name + ".norm"
        y, rm, rv = batch_norm(x, P[name + ".weight"], P[name + ".bias"], P[name + ".running_mean"],
                               P[name + ".running_var"], train)
        if bn_out is not None and train:
            bn_out[name + ".running_mean"], bn_out[name + ".running_var"] = rm, rv
        return _modulate(p if p is not None else pt, P, mod, y)

    def rec(name, p):
        if inter is not None:
            inter[name] = p["feat"]

    # stem (model.py:831-861)
    nbr5 = neighbor_table(batch, grid, 5)
    x = subm_conv(feat, nbr5, P["embedding.stem.conv.weight"], None)
    pt["feat"] = F.gelu(bn(x, "embedding.stem.norm"))
    rec("embedding", pt)
    stack = []
    for s in range(S):
        if s > 0:
            parent = pt
            st = pool_structure(parent, perms[s] if perms is not None else None, cfg["stride"][s - 1])
            pn = f"enc.enc{s}.down."
            proj = F.linear(parent["feat"], P[pn + "proj.weight"], P[pn + "proj.bias"])
            f = segment_max(proj[st["indices"]], st["idx_ptr"])
            c = segment_mean(parent["coord"][st["indices"]], st["idx_ptr"])
            stack.append((parent, st["cluster"]))
            pt = dict(feat=f, coord=c, grid=st["grid"], batch=st["batch"], offset=st["offset"],
                      context=context, context_offset=context_offset, code=st["code"], order=st["order"],
                      inverse=st["inverse"], depth=st["depth"])
            pt["feat"] = F.gelu(bn(f, pn + "norm.0", pt))
            _stage_aux(pt, cfg["enc_patch_size"][s])
            rec(f"enc{s}.down", pt)
        for i in range(cfg["enc_depths"][s]):
            pt = block(pt, P, f"enc.enc{s}.block{i}.", cfg["enc_num_head"][s], i % len(cfg["order"]),
                       cfg["enc_patch_size"][s], qk_norm)
            rec(f"enc{s}.block{i}", pt)
            if f"enc.enc{s}.ca_block{i}.norm1.0.weight" in P:
                pt = ca_block(pt, P, f"enc.enc{s}.ca_block{i}.", cfg["enc_num_head"][s], qk_norm)
                rec(f"enc{s}.ca_block{i}", pt)
    outs = [pt]
    for s in reversed(range(S - 1)):
        parent, cluster = stack.pop()
        un = f"dec.dec{s}.up."
        a = F.gelu(bn(F.linear(pt["feat"], P[un + "proj.0.weight"], P[un + "proj.0.bias"]), un + "proj.1", pt))
        b = F.gelu(bn(F.linear(parent["feat"], P[un + "proj_skip.0.weight"], P[un + "proj_skip.0.bias"]),
                      un + "proj_skip.1", parent))
        parent = dict(parent)
        parent["feat"] = b + a[cluster]
        parent["conv_feat"] = b            # sparse_conv_feat is NOT refreshed by the add (model.py:824)
        pt = parent
        rec(f"dec{s}.up", pt)
        for i in range(cfg["dec_depths"][s]):
            pt = block(pt, P, f"dec.dec{s}.block{i}.", cfg["dec_num_head"][s], i % len(cfg["order"]),
                       cfg["dec_patch_size"][s], qk_norm)
            rec(f"dec{s}.block{i}", pt)
            if f"dec.dec{s}.ca_block{i}.norm1.0.weight" in P:
                pt = ca_block(pt, P, f"dec.dec{s}.ca_block{i}.", cfg["dec_num_head"][s], qk_norm)
                rec(f"dec{s}.ca_block{i}", pt)
            outs.append(pt)
    return outs


def action_head(x, P, npoints, pos_bins=15, euler_bins=72):
    """ActionHead.forward, heatmap_disc + max + euler_disc (simple_policy_ptv3.py:70-157), dropout 0."""
    pre = "act_proj_head."
    h = F.leaky_relu(F.linear(x, P[pre + "heatmap_mlp.0.weight"], P[pre + "heatmap_mlp.0.bias"]), 0.02)
    xt = F.linear(h, P[pre + "heatmap_mlp.3.weight"], P[pre + "heatmap_mlp.3.bias"])
    xt = xt.view(x.shape[0], 3, 2 * pos_bins).permute(1, 0, 2)               # 'n (c b) -> c n b'
    pc = torch.stack([t.max(0)[0] for t in torch.split(x, npoints)], 0)
    h = F.leaky_relu(F.linear(pc, P[pre + "action_mlp.0.weight"], P[pre + "action_mlp.0.bias"]), 0.02)
    a = F.linear(h, P[pre + "action_mlp.3.weight"], P[pre + "action_mlp.3.bias"])
    xr = a[..., : euler_bins * 3].reshape(-1, euler_bins, 3)
    xo = a[..., -1]
    return xt, xr, xo


def policy_loss(xt, xr, xo, gt_actions, disc_pos_probs, npoints, pos_weight=1.0, rot_weight=1.0):
    """compute_loss (simple_policy_ptv3.py:308-373), heatmap_disc + euler_disc."""
    pos = 0
    for i, t in enumerate(torch.split(xt, npoints, dim=1)):
        pos = pos + F.cross_entropy(t.reshape(3, -1), disc_pos_probs[i], reduction="mean")
    pos = pos / len(npoints)
    rot = F.cross_entropy(xr, gt_actions[..., 3:-1].long(), reduction="mean")
    opn = F.binary_cross_entropy_with_logits(xo, gt_actions[..., -1], reduction="mean")
    return {"pos": pos, "rot": rot, "open": opn, "total": pos_weight * pos + rot_weight * rot + opn}


def euler_bins_to_quat(bins, resolution=5):
    """utils/rotation_transform.py:192-193: extrinsic 'xyz' euler (deg) -> quaternion xyzw, closed form
    (scipy: q = qz * qy * qx), canonicalised like scipy >= 1.? does NOT canonicalise sign; we match scipy."""
    e = np.deg2rad(np.asarray(bins, dtype=np.float64) * resolution - 180.0)
    cx, sx = np.cos(e[..., 0] / 2), np.sin(e[..., 0] / 2)
    cy, sy = np.cos(e[..., 1] / 2), np.sin(e[..., 1] / 2)
    cz, sz = np.cos(e[..., 2] / 2), np.sin(e[..., 2] / 2)
    x = sx * cy * cz - cx * sy * sz
    y = cx * sy * cz + sx * cy * sz
    z = cx * cy * sz - sx * sy * cz
    w = cx * cy * cz + sx * sy * sz
    return np.stack([x, y, z, w], -1)


def best_pos_from_logits(xt, coord, npoints, pos_bins=15, pos_bin_size=0.01):
    """final-action position decode (simple_policy_ptv3.py:250-275; utils/action_position_utils.py:48-60,
    best='max'): argmax over (n_i * 2*pos_bins) per axis of the per-cloud softmax -> xyz[n,c] + (b-pos_bins)*size."""
    outs = []
    shift = np.arange(-pos_bins, pos_bins) * pos_bin_size
    for t, c in zip(torch.split(xt, npoints, dim=1), torch.split(coord, npoints)):
        prob = torch.softmax(t.reshape(3, -1), dim=-1).numpy()
        idx = prob.argmax(-1)
        xyz = c.numpy()
        cands = xyz[:, :, None] + shift[None, None, :]                         # float64 via numpy promotion
        cands = np.transpose(cands, (1, 0, 2)).reshape(3, -1)
        outs.append(cands[np.arange(3), idx])
    return torch.from_numpy(np.array(outs)).float()


def robot_pose_embedding(sd, pre, poses):
    """RobotPoseEmbedding.forward (models/base.py:61-78): Linear(pos) + Embedding(open) + Linear(sin/cos euler 'xyz'),
    LayerNorm eps 1e-12; the euler angles come from scipy on the host exactly as in the reference."""
    from scipy.spatial.transform import Rotation as R
    e = torch.from_numpy(R.from_quat(poses[..., 3:7].detach().cpu().numpy()).as_euler("xyz")).float()
    rot_in = torch.cat([torch.sin(e), torch.cos(e)], -1)
    x = F.linear(poses[..., :3], sd[pre + "pos_embedding.weight"], sd[pre + "pos_embedding.bias"]) \
        + F.linear(rot_in, sd[pre + "rot_embedding.weight"], sd[pre + "rot_embedding.bias"]) \
        + sd[pre + "open_embedding.weight"][poses[..., -1].long()]
    return F.layer_norm(x, (x.shape[-1],), sd[pre + "layer_norm.weight"], sd[pre + "layer_norm.bias"], 1e-12)


def policy_context(sd, cfg, batch, variant):
    """prepare_ptv3_batch context: CA (simple_policy_ptv3.py:403-431; motion_planner_ptv3.py:437-463) -> packed tokens +
    offsets; AdaNorm / Concat (:194-223, :438-463; planner :201-228) -> one vector per cloud."""
    ac = cfg["action_config"]
    lens = [int(v) for v in batch["txt_lens"]]
    ctx = F.linear(batch["txt_embeds"], sd["txt_fc.weight"], sd["txt_fc.bias"])
    pose = robot_pose_embedding(sd, "pose_embedding.", batch["ee_poses"]) if ac.get("use_ee_pose") else None
    step = sd["stepid_embedding.weight"][batch["step_ids"]] if (ac.get("use_step_id") and "stepid_embedding.weight" in sd) else None
    if variant == "ca":
        toks = list(torch.split(ctx, lens))
        n = torch.tensor(lens, dtype=torch.long)
        for extra in (pose, step):
            if extra is not None:
                toks = [torch.cat([c, e.unsqueeze(0)], 0) for c, e in zip(toks, extra)]
                n = n + 1
        return torch.cat(toks, 0), torch.cumsum(n, 0)
    if ac.get("txt_reduce") == "attn" and variant != "concat":
        w = torch.split(F.linear(batch["txt_embeds"], sd["txt_attn_fc.weight"], sd["txt_attn_fc.bias"]), lens)
        ctx = torch.stack([(torch.softmax(wi, 0) * ci).sum(0) for wi, ci in zip(w, torch.split(ctx, lens))], 0)
    if pose is not None:
        ctx = ctx + pose
    if step is not None:
        ctx = ctx + step
    return ctx, None


def policy_forward(sd, cfg, batch, perms=None, train=True, compute_loss=True, compute_final_action=False,
                   bn_out=None, inter=None, variant="ca"):
    """SimplePolicyPTV3{CA,AdaNorm,Concat}.forward (simple_policy_ptv3.py:225-306, 194-223, 403-463).
    sd: state dict (reference names); cfg: MODEL config mapping; perms: (5,4) recorded shuffles or None."""
    pc, ac = cfg["ptv3_config"], cfg["action_config"]
    fts = batch["pc_fts"]
    npoints = list(batch["npoints_in_batch"])
    ctx, ctx_off = policy_context(sd, cfg, batch, variant)
    feat = fts
    if variant == "concat":
        feat = torch.cat([fts, torch.repeat_interleave(ctx, torch.tensor(npoints), 0)], -1)
        ctx = None
    P = {k[len("ptv3_model."):]: v for k, v in sd.items() if k.startswith("ptv3_model.")}
    bn_local = {} if bn_out is not None else None
    outs = ptv3ca_forward(P, pc, feat, fts[:, :3], batch["offset"], ctx, ctx_off, perms, train, bn_local, inter)
    if bn_out is not None:
        bn_out.update({"ptv3_model." + k: v for k, v in bn_local.items()})
    last = outs[-1]
    xt, xr, xo = action_head(last["feat"], sd, npoints, ac["pos_bins"], 72)
    if compute_final_action:
        pos = best_pos_from_logits(xt.detach(), last["coord"], npoints, ac["pos_bins"], ac["pos_bin_size"])
    else:
        pos = batch["gt_actions"][..., :3]
    rot = torch.from_numpy(euler_bins_to_quat(torch.argmax(xr, 1).numpy(), 5))
    final = torch.cat([pos, rot, xo.detach().unsqueeze(-1)], -1)
    res = {"final": final, "xt": xt, "xr": xr, "xo": xo, "dec_feat": last["feat"]}
    if compute_loss:
        lc = cfg["loss_config"]
        res["losses"] = policy_loss(xt, xr, xo, batch["gt_actions"], batch["disc_pos_probs"], npoints,
                                    lc["pos_weight"], lc["rot_weight"])
    return res


# --------------------------------------------------------------------------------------------------
# §8a row 19: motion planner (models/motion_planner_ptv3.py)
# --------------------------------------------------------------------------------------------------


def traj_action_head(x, P, npoints, T, pos_bins=15, euler_bins=72):
    """trajectory ActionHead.forward, heatmap_disc + max + euler_disc (motion_planner_ptv3.py:77-148), dropout 0"""
    pre = "act_proj_head."
    pe = x.unsqueeze(1).expand(-1, T, -1)
    te = P[pre + "traj_embedding.weight"][:T]
    pe = torch.cat([pe, te.expand(pe.size(0), -1, -1)], -1)                         # (N, T, C + E)
    h = F.leaky_relu(F.linear(pe, P[pre + "heatmap_mlp.0.weight"], P[pre + "heatmap_mlp.0.bias"]), 0.02)
    xt = F.linear(h, P[pre + "heatmap_mlp.3.weight"], P[pre + "heatmap_mlp.3.bias"])   # (N, T, 3*2*pos_bins)
    xt = xt.view(x.shape[0], T, 3, 2 * pos_bins).permute(1, 2, 0, 3)                  # 'n t (c b) -> t c n b'
    pc = torch.stack([t.max(0)[0] for t in torch.split(pe, npoints)], 0)              # (B, T, C + E)
    h = F.leaky_relu(F.linear(pc, P[pre + "action_mlp.0.weight"], P[pre + "action_mlp.0.bias"]), 0.02)
    a = F.linear(h, P[pre + "action_mlp.3.weight"], P[pre + "action_mlp.3.bias"])     # (B, T, 72*3 + 2)
    xr = a[..., : euler_bins * 3].reshape(a.shape[0], T, euler_bins, 3)
    return xt, xr, a[..., -2], a[..., -1]


def traj_loss(xt, xr, xo, xstop, gt_trajs, gt_stop, masks, disc_pos_probs, npoints, pos_weight=1.0, rot_weight=1.0):
    """MotionPlannerPTV3*.compute_loss (motion_planner_ptv3.py:307-397), heatmap_disc + euler_disc"""
    B, T = gt_trajs.shape[:2]
    m = masks.float()
    pos = 0
    for i, t in enumerate(torch.split(xt, npoints, dim=2)):
        mk = m[i].unsqueeze(1).expand(-1, 3).reshape(-1)
        ce = F.cross_entropy(t.reshape(T * 3, -1), disc_pos_probs[i].reshape(T * 3, -1), reduction="none")
        pos = pos + torch.sum(ce * mk) / mk.sum()
    pos = pos / len(npoints)
    tgt = gt_trajs[..., 3:-1].long()
    ce = F.cross_entropy(xr.permute(0, 1, 3, 2).reshape(B * T * 3, -1), tgt.reshape(-1), reduction="none").reshape(B, T, 3)
    rot = torch.sum(ce * m.unsqueeze(-1)) / m.sum() / 3
    opn = (F.binary_cross_entropy_with_logits(xo, gt_trajs[..., -1], reduction="none") * m).sum() / m.sum()
    stop = (F.binary_cross_entropy_with_logits(xstop, gt_stop.float(), reduction="none") * m).sum() / m.sum()
    return {"pos": pos, "rot": rot, "open": opn, "stop": stop,
            "total": pos_weight * pos + rot_weight * rot + opn + stop}


def planner_forward(sd, cfg, batch, perms=None, train=True, compute_loss=True, compute_final_action=False,
                    bn_out=None, inter=None, variant="ca"):
    """MotionPlannerPTV3{CA,AdaNorm}.forward (motion_planner_ptv3.py:201-305, 437-463)"""
    pc, ac = cfg["ptv3_config"], cfg["action_config"]
    T = ac["max_traj_len"]
    fts = batch["pc_fts"]
    npoints = list(batch["npoints_in_batch"])
    feat = torch.cat([fts, sd["pc_label_embedding.weight"][batch["pc_labels"]]], -1)
    ctx, ctx_off = policy_context(sd, cfg, batch, variant)
    P = {k[len("ptv3_model."):]: v for k, v in sd.items() if k.startswith("ptv3_model.")}
    bn_local = {} if bn_out is not None else None
    outs = ptv3ca_forward(P, pc, feat, fts[:, :3], batch["offset"], ctx, ctx_off, perms, train, bn_local, inter)
    if bn_out is not None:
        bn_out.update({"ptv3_model." + k: v for k, v in bn_local.items()})
    last = outs[-1]
    xt, xr, xo, xs = traj_action_head(last["feat"], sd, npoints, T, ac["pos_bins"], 72)
    B = len(npoints)
    if compute_final_action:
        pos = torch.stack([best_pos_from_logits(xt[t].detach(), last["coord"], npoints, ac["pos_bins"], ac["pos_bin_size"])
                           for t in range(T)], 1)
    else:
        pos = batch["gt_trajs"][..., :3]
    rot = torch.from_numpy(euler_bins_to_quat(torch.argmax(xr.reshape(B * T, 72, 3), 1).numpy(), 5)).reshape(B, T, 4)
    final = torch.cat([pos, rot, xo.detach().unsqueeze(-1), xs.detach().unsqueeze(-1)], -1)
    res = {"final": final, "xt": xt, "xr": xr, "xo": xo, "xstop": xs, "dec_feat": last["feat"]}
    if compute_loss:
        lc = cfg["loss_config"]
        res["losses"] = traj_loss(xt, xr, xo, xs, batch["gt_trajs"], batch["gt_trajs_stop"], batch["traj_masks"],
                                  batch["gt_trajs_disc_pos_probs"], npoints, lc["pos_weight"], lc["rot_weight"])
    return res


# --------------------------------------------------------------------------------------------------
# §8a rows 12-13: FPS (reference helper) and kNN / ball query (no reference function; brute force)
# --------------------------------------------------------------------------------------------------


def farthest_point_sample_indices(xyz, npoint, start):
    """train/datasets/common.py:108-129 with the random start made explicit.  xyz (N,3) float32; the
    reference computes in the array's dtype: dist = sum((xyz - c)**2, -1) in x,y,z order; argmax -> lowest index."""
    xyz = np.asarray(xyz)[:, :3]
    n = xyz.shape[0]
    idx = np.zeros(npoint, dtype=np.int64)
    distance = np.full(n, 1e10, dtype=np.float64)
    far = int(start)
    for i in range(npoint):
        idx[i] = far
        d = np.sum((xyz - xyz[far]) ** 2, -1)
        m = d < distance
        distance[m] = d[m]
        far = int(np.argmax(distance))
    return idx


def knn_indices(query, ref, k, q_offset=None, r_offset=None):
    """brute-force kNN (squared L2 in fp32, x,y,z summation order), ties -> lowest index; per cloud."""
    def one(q, r):
        d = ((q[:, None, 0] - r[None, :, 0]) ** 2 + (q[:, None, 1] - r[None, :, 1]) ** 2) + (q[:, None, 2] - r[None, :, 2]) ** 2
        order = np.argsort(d, axis=1, kind="stable")[:, :k]
        return order, np.take_along_axis(d, order, 1)
    q, r = np.asarray(query, np.float32), np.asarray(ref, np.float32)
    if q_offset is None:
        return one(q, r)
    idx, dist = [], []
    qs = rs = 0
    for qe, re in zip(q_offset, r_offset):
        i, d = one(q[qs:qe], r[rs:re])
        idx.append(i + rs)
        dist.append(d)
        qs, rs = qe, re
    return np.concatenate(idx), np.concatenate(dist)


def ball_query_indices(query, ref, radius, nsample):
    """first `nsample` reference indices (ascending index) within `radius`; missing slots repeat the first hit
    (pointnet2 convention), -1 if none."""
    q, r = np.asarray(query, np.float32), np.asarray(ref, np.float32)
    d = ((q[:, None, 0] - r[None, :, 0]) ** 2 + (q[:, None, 1] - r[None, :, 1]) ** 2) + (q[:, None, 2] - r[None, :, 2]) ** 2
    out = np.full((q.shape[0], nsample), -1, dtype=np.int64)
    r2 = np.float32(radius) * np.float32(radius)
    for i in range(q.shape[0]):
        hit = np.nonzero(d[i] < r2)[0][:nsample]
        if hit.size:
            out[i, :] = hit[0]
            out[i, : hit.size] = hit
    return out


# ---------------------------------------------------------------------------------------------------------------------
# optimizer step (train/train_simple_policy.py:237-243 -> torch clip_grad_norm_; train/optim/adamw.py:53-112;
# parameter grouping train/optim/misc.py:13-55: names containing 'bias' get weight_decay 0)
def clip_coefficient(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_: total L2 norm over all tensors, coefficient clamped to 1"""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    return float(torch.clamp(max_norm / (total + 1e-6), max=1.0)), float(total)


def adamw_step(params, grads, exp_avg, exp_avg_sq, step, lr, weight_decays, betas=(0.9, 0.98), eps=1e-6, max_grad_norm=10.0):
    """in-place reference AdamW over lists of tensors (`step` is the 1-based step being taken); returns the grad norm"""
    norm = None
    if max_grad_norm and max_grad_norm > 0:
        coef, norm = clip_coefficient(grads, max_grad_norm)
        grads = [g * coef for g in grads]
    b1, b2 = betas
    step_size = lr * math.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    for p, g, m, v, wd in zip(params, grads, exp_avg, exp_avg_sq, weight_decays):
        m.mul_(b1).add_(g, alpha=1.0 - b1)
        v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
        p.addcdiv_(m, v.sqrt().add_(eps), value=-step_size)
        if wd > 0.0:
            p.add_(p, alpha=-lr * wd)
    return norm


# ---------------------------------------------------------------------------------------------------------------------
# batch producer: geometric augmentation + feature assembly of one keystep
# (train/datasets/simple_policy_dataset.py:158-181 `_augment_pc`, :281-296 subsample + height, :321-346 centring / features;
#  train/datasets/common.py:81-87 `random_rotate_z`).  The random draws (point indices, angle, noise) are INPUTS.
def augment_keystep(xyz, rgb, point_idxs, angle, noise, gt_pos, table_height, xyz_shift="center", use_height=True):
    """xyz (m,3) float32 raw, rgb (m,3) uint8 or None -> pc_ft (n, 7|6|4|3) float32, centroid (3,) float64, gt position float32"""
    xyz = np.asarray(xyz)[point_idxs]
    height = xyz[:, -1] - table_height
    cosval, sinval = np.cos(angle), np.sin(angle)
    Rm = np.array([[cosval, -sinval, 0], [sinval, cosval, 0], [0, 0, 1]])
    xyz = np.dot(xyz, np.transpose(Rm))
    gt = np.dot(np.asarray(gt_pos, dtype=np.float64), np.transpose(Rm))
    xyz = noise + xyz
    centroid = np.mean(xyz, 0) if xyz_shift == "center" else np.zeros(3)
    xyz = xyz - centroid
    gt = gt - centroid
    cols = [xyz]
    if rgb is not None:
        cols.append((np.asarray(rgb)[point_idxs] / 255.) * 2 - 1)
    if use_height:
        cols.append(height[:, None])
    return torch.from_numpy(np.concatenate(cols, 1)).float(), centroid, torch.from_numpy(gt).float()


# ---------------------------------------------------------------------------------------------------------------------
# inference front end: workspace crop + voxel down-sample (evaluation/eval_simple_policy.py:151-175).
# PARITY UNPINNED for this function: open3d 0.18.0 (`voxel_down_sample_and_trace`) is a third-party dependency that is neither under
# /root/reference nor installed offline, so no reference output could be minted; this restates its published algorithm
# (open3d/geometry/PointCloud.cpp VoxelDownSampleAndTrace: voxel_min_bound = min_bound - voxel_size / 2, index = floor((p - voxel_min_bound)
# / voxel_size), output = mean of the voxel's points, trace = its point indices in input order) anchored on the call site, which keeps
# the colour of trace[v][0].  Voxel ORDER: open3d's is unordered_map iteration order; fixed here to first appearance.
def voxel_downsample(xyz, rgb, crop_lo, crop_hi, voxel_size):
    """xyz (n,3) float32, rgb (n,3) uint8 -> (mean xyz float64 (m,3), rgb uint8 (m,3), first index (m,))"""
    xyz = np.asarray(xyz).reshape(-1, 3)
    lo, hi = np.asarray(crop_lo, dtype=np.float64), np.asarray(crop_hi, dtype=np.float64)
    in_mask = np.all((xyz > lo) & (xyz < hi), 1)
    idx = np.nonzero(in_mask)[0]
    pts = xyz[idx].astype(np.float64)
    if len(pts) == 0:
        return np.zeros((0, 3)), np.zeros((0, 3), dtype=np.uint8), np.zeros((0,), dtype=np.int64)
    vmin = np.min(pts, 0) - voxel_size * 0.5
    vox = np.floor((pts - vmin) / voxel_size).astype(np.int64)
    _, inv = np.unique(vox, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    m = int(inv.max()) + 1
    first = np.full(m, np.iinfo(np.int64).max)
    np.minimum.at(first, inv, idx)
    cnt = np.bincount(inv, minlength=m).astype(np.float64)
    mean = np.stack([np.bincount(inv, weights=pts[:, a], minlength=m) / cnt for a in range(3)], 1)
    order = np.argsort(first)
    first = first[order]
    return mean[order], (np.asarray(rgb).reshape(-1, 3)[first] if rgb is not None else None), first
