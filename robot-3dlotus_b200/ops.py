"""torch-facing wrappers of the C ABI: output allocation (torch caching allocator), stream plumbing and
`torch.autograd.Function`s whose forward AND backward are C-ABI kernel calls.  torch provides memory,
streams and the tape; it computes nothing on this path.
"""
import ctypes

import torch

from .lib import load, LotusError  # noqa: F401


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _po(t, byte_off):
    return ctypes.c_void_p(t.data_ptr() + byte_off)


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t):
    assert t.dtype == torch.float32 and t.is_cuda, (t.dtype, t.device)
    return t if t.is_contiguous() else t.contiguous()


class DropCtx:
    """dropout bookkeeping of one forward: a seed plus a running stream id per dropout site."""

    def __init__(self, seed=0, enabled=False):
        self.seed, self.enabled, self.next = int(seed) & (2 ** 63 - 1), enabled, 1

    def site(self, p):
        if not self.enabled or p <= 0:
            return 0.0, 0, 0
        self.next += 1
        return float(p), self.seed, self.next


NO_DROP = (0.0, 0, 0)

# Gradient sinks: parameter storage address -> fp32 view inside the trainer's flat gradient buffer.  When a
# parameter has a sink, the backward kernels ACCUMULATE its gradient there directly ("flat grad
# buffer written directly by the backward kernels") and autograd receives None for it.
GRAD_SINK = {}


def _sink(param):
    return GRAD_SINK.get(param.data_ptr()) if param is not None else None


def _grad_buf(param, needs):
    """-> (buffer to accumulate into or None, value to hand back to autograd)"""
    if param is None or not needs:
        return None, None
    s = _sink(param)
    if s is not None:
        return s, None
    z = torch.zeros_like(param)
    return z, z

# ------------------------------------------------------------------------------------------ Linear


class LinearFn(torch.autograd.Function):
    """y = [residual +] drop(x @ w.T + b)"""

    @staticmethod
    def forward(ctx, x, w, b, residual, drop):
        L = load()
        x, w = _f32c(x), _f32c(w)
        M, K = x.shape
        N = w.shape[0]
        y = torch.empty((M, N), device=x.device, dtype=torch.float32)
        if residual is not None:
            residual = _f32c(residual)
        L.lotus_linear_fwd_f32(_p(x), _p(w), _p(b), _p(residual), _p(y), M, N, K, 0, None, drop[0], drop[1], drop[2], _st())
        ctx.save_for_backward(x, w)
        ctx.bias = b
        ctx.has_b, ctx.has_res, ctx.drop = b is not None, residual is not None, drop
        return y

    @staticmethod
    def backward(ctx, dy):
        L = load()
        x, w = ctx.saved_tensors
        dy = _f32c(dy)
        M, K = x.shape
        N = w.shape[0]
        dres = dy if ctx.has_res else None
        if ctx.drop[0] > 0:
            dyd = torch.empty_like(dy)
            L.lotus_dropout_f32(_p(dy), _p(dyd), dy.numel(), ctx.drop[0], ctx.drop[1], ctx.drop[2], _st())
            dy = dyd
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            L.lotus_linear_dgrad_f32(_p(dy), _p(w), None, _p(dx), M, N, K, _st())
        wbuf, dw = _grad_buf(w, ctx.needs_input_grad[1])
        if wbuf is not None:
            L.lotus_linear_wgrad_f32(_p(dy), _p(x), _p(wbuf), M, N, K, None, _st())
        bbuf, db = _grad_buf(ctx.bias, ctx.has_b and ctx.needs_input_grad[2])
        if bbuf is not None:
            L.lotus_colsum_f32(_p(dy), M, N, _p(bbuf), _st())
        return dx, dw, db, dres, None


def linear(x, w, b=None, residual=None, drop=NO_DROP):
    return LinearFn.apply(x, w, b, residual, drop)


# ------------------------------------------------------------------------------------------ pointwise


class ActFn(torch.autograd.Function):
    """y = drop(act(x)); act: 'gelu' | 'leaky' (slope 0.02)"""

    @staticmethod
    def forward(ctx, x, kind, drop):
        L = load()
        x = _f32c(x)
        y = torch.empty_like(x)
        op = 0 if kind == "gelu" else 2
        L.lotus_pointwise_f32(_p(x), None, _p(y), x.numel(), op, 0.02, drop[0], drop[1], drop[2], _st())
        ctx.save_for_backward(x)
        ctx.op, ctx.drop = op, drop
        return y

    @staticmethod
    def backward(ctx, dy):
        L = load()
        (x,) = ctx.saved_tensors
        dy = _f32c(dy)
        dx = torch.empty_like(x)
        L.lotus_pointwise_f32(_p(x), _p(dy), _p(dx), x.numel(), ctx.op + 1, 0.02, ctx.drop[0], ctx.drop[1], ctx.drop[2], _st())
        return dx, None, None


def act(x, kind, drop=NO_DROP):
    return ActFn.apply(x, kind, drop)


# ------------------------------------------------------------------------------------------ norms


class LayerNormFn(torch.autograd.Function):
    """y = [residual +] LN(x) over the last dim"""

    @staticmethod
    def forward(ctx, x, g, b, eps, residual):
        L = load()
        x = _f32c(x)
        rows, C = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        if residual is not None:
            residual = _f32c(residual)
        L.lotus_layernorm_fwd_f32(_p(x), 1, C, 0, _p(y), 1, C, 0, _p(residual), _p(g), _p(b), _p(mean), _p(rstd), rows, C,
                                  eps, _st())
        ctx.save_for_backward(x, g, mean, rstd)
        ctx.beta = b
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        L = load()
        x, g, mean, rstd = ctx.saved_tensors
        dy = _f32c(dy)
        rows, C = x.shape
        dx = torch.empty_like(x)
        gbuf, dg = _grad_buf(g, True)
        bbuf, db = _grad_buf(ctx.beta, True)
        L.lotus_layernorm_bwd_f32(_p(dy), 1, C, 0, _p(x), 1, C, 0, _p(dx), 1, C, 0, _p(g), _p(mean), _p(rstd), _p(gbuf), _p(bbuf),
                                  rows, C, 0, _st())
        return dx, dg, db, None, (dy if ctx.has_res else None)


def layer_norm(x, g, b, eps=1e-5, residual=None):
    return LayerNormFn.apply(x, g, b, eps, residual)


class BatchNormActFn(torch.autograd.Function):
    """y = [gelu](BatchNorm1d(x)); updates running stats in place when training"""

    @staticmethod
    def forward(ctx, x, g, b, run_mean, run_var, train, eps, momentum, gelu):
        L = load()
        x = _f32c(x)
        rows, C = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(C, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        scratch = torch.empty_like(mean)
        L.lotus_batchnorm_fwd_f32(_p(x), _p(y), rows, C, _p(g), _p(b), _p(run_mean), _p(run_var), _p(mean), _p(rstd),
                                  _p(scratch), int(train), eps, momentum, int(gelu), _st())
        ctx.save_for_backward(x, g, b, mean, rstd)
        ctx.train, ctx.gelu = train, gelu
        return y

    @staticmethod
    def backward(ctx, dy):
        L = load()
        x, g, b, mean, rstd = ctx.saved_tensors
        dy = _f32c(dy)
        rows, C = x.shape
        dx = torch.empty_like(x)
        dg = torch.zeros(C, device=x.device, dtype=torch.float32)
        db = torch.zeros_like(dg)
        L.lotus_batchnorm_bwd_f32(_p(dy), _p(x), rows, C, _p(g), _p(b), _p(mean), _p(rstd), _p(dx), _p(dg), _p(db),
                                  int(ctx.train), int(ctx.gelu), _st())
        for param, d in ((g, dg), (b, db)):
            s = _sink(param)
            if s is not None:
                L.lotus_pointwise_f32(_p(d), _p(s), _p(s), C, 4, 0.0, 0.0, 0, 0, _st())     # sink += batch sums
        if _sink(g) is not None:
            dg = db = None
        return dx, dg, db, None, None, None, None, None, None


def batch_norm_act(x, bn, train, gelu=True):
    if train:
        bn.num_batches_tracked.add_(1)
    return BatchNormActFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, train, bn.eps, bn.momentum, gelu)


# ------------------------------------------------------------------------------------------ sparse conv


class SubmConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, nbr, nbr_stride):
        L = load()
        x, w = _f32c(x), _f32c(w)
        n, cin = x.shape
        cout = w.shape[0]
        k3 = w.shape[1] * w.shape[2] * w.shape[3]
        y = torch.empty((n, cout), device=x.device, dtype=torch.float32)
        L.lotus_subm_conv_fwd_f32(_p(x), _p(w), _p(b), _p(y), _p(nbr), nbr_stride, n, cin, cout, k3, _st())
        ctx.save_for_backward(x, w, nbr)
        ctx.bias = b
        ctx.meta = (n, cin, cout, k3, nbr_stride, b is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = load()
        x, w, nbr = ctx.saved_tensors
        n, cin, cout, k3, nbr_stride, has_b = ctx.meta
        dy = _f32c(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            L.lotus_subm_conv_dgrad_f32(_p(dy), _p(w), _p(dx), _p(nbr), nbr_stride, n, cin, cout, k3, _st())
        wbuf, dw = _grad_buf(w, ctx.needs_input_grad[1])
        if wbuf is not None:
            L.lotus_subm_conv_wgrad_f32(_p(dy), _p(x), _p(wbuf), _p(nbr), nbr_stride, n, cin, cout, k3, _st())
        bbuf, db = _grad_buf(ctx.bias, has_b and ctx.needs_input_grad[2])
        if bbuf is not None:
            L.lotus_colsum_f32(_p(dy), n, cout, _p(bbuf), _st())
        return dx, dw, db, None, None


def subm_conv(x, w, b, nbr, nbr_stride):
    return SubmConvFn.apply(x, w, b, nbr, nbr_stride)


# ------------------------------------------------------------------------------------------ attention


class PatchAttnFn(torch.autograd.Function):
    """qkv (N,3C) raw projections -> per-head LN on q,k -> attention inside each serialized patch -> (N,C) in
    point order (gather by order[pad] and scatter by unpad[inverse] fused)."""

    @staticmethod
    def forward(ctx, qkv, qg, qb, kg, kb, rpe_table, plan, H, qk_norm, attn_drop, grid, pos_bnd):
        L = load()
        qkv = _f32c(qkv)
        N, C3 = qkv.shape
        C = C3 // 3
        D = C // H
        dev = qkv.device
        st = _st()
        if qk_norm:
            qh = torch.empty((N, C), device=dev, dtype=torch.float32)
            kh = torch.empty_like(qh)
            stats = torch.empty((4, N * H), device=dev, dtype=torch.float32)
            L.lotus_layernorm_fwd_f32(_p(qkv), H, C3, D, _p(qh), 1, D, 0, None, _p(qg), _p(qb), _p(stats[0]), _p(stats[1]),
                                      N * H, D, 1e-6, st)
            L.lotus_layernorm_fwd_f32(_po(qkv, C * 4), H, C3, D, _p(kh), 1, D, 0, None, _p(kg), _p(kb), _p(stats[2]),
                                      _p(stats[3]), N * H, D, 1e-6, st)
            q_ptr, k_ptr, ldq = _p(qh), _p(kh), C
        else:
            qh = kh = stats = None
            q_ptr, k_ptr, ldq = _p(qkv), _po(qkv, C * 4), C3
        out = torch.empty((N, C), device=dev, dtype=torch.float32)
        lse = torch.empty((plan.n_pad * H,), device=dev, dtype=torch.float32)
        L.lotus_patch_attn_fwd_f32(q_ptr, k_ptr, _po(qkv, 2 * C * 4), ldq, ldq, C3, _p(plan.gather_idx), _p(plan.slot_of_point),
                                   _p(plan.cu_seqlens), plan.n_patches, plan.max_len, H, D, float(D) ** -0.5,
                                   _p(rpe_table), _p(grid) if rpe_table is not None else None, pos_bnd, _p(out), C, _p(lse),
                                   attn_drop[0], attn_drop[1], attn_drop[2], st)
        ctx.save_for_backward(qkv, qg, kg, qh, kh, stats, out, lse, rpe_table, grid if rpe_table is not None else None)
        ctx.plan, ctx.H, ctx.qk_norm, ctx.attn_drop, ctx.pos_bnd = plan, H, qk_norm, attn_drop, pos_bnd
        ctx.nb = (qb, kb)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = load()
        qkv, qg, kg, qh, kh, stats, out, lse, rpe_table, grid = ctx.saved_tensors
        plan, H = ctx.plan, ctx.H
        dout = _f32c(dout)
        N, C3 = qkv.shape
        C = C3 // 3
        D = C // H
        dev = qkv.device
        st = _st()
        dqkv = torch.zeros_like(qkv)
        drpe = torch.zeros_like(rpe_table) if rpe_table is not None else None
        if ctx.qk_norm:
            dqh = torch.zeros((N, C), device=dev, dtype=torch.float32)
            dkh = torch.zeros_like(dqh)
            q_ptr, k_ptr, ldq = _p(qh), _p(kh), C
            dq_ptr, dk_ptr, lddq = _p(dqh), _p(dkh), C
        else:
            q_ptr, k_ptr, ldq = _p(qkv), _po(qkv, C * 4), C3
            dq_ptr, dk_ptr, lddq = _p(dqkv), _po(dqkv, C * 4), C3
        L.lotus_patch_attn_bwd_f32(q_ptr, k_ptr, _po(qkv, 2 * C * 4), ldq, ldq, C3, _p(plan.gather_idx), _p(plan.slot_of_point),
                                   _p(plan.cu_seqlens), plan.n_patches, plan.max_len, H, D, float(D) ** -0.5, _p(rpe_table),
                                   _p(grid), ctx.pos_bnd, _p(drpe), _p(out), _p(dout), C, _p(lse), dq_ptr, dk_ptr,
                                   _po(dqkv, 2 * C * 4), lddq, lddq, C3, ctx.attn_drop[0], ctx.attn_drop[1], ctx.attn_drop[2], st)
        dqg = dqb = dkg = dkb = None
        if ctx.qk_norm:
            (qgb, dqg), (qbb, dqb) = _grad_buf(qg, True), _grad_buf(ctx.nb[0], True)
            (kgb, dkg), (kbb, dkb) = _grad_buf(kg, True), _grad_buf(ctx.nb[1], True)
            L.lotus_layernorm_bwd_f32(_p(dqh), 1, D, 0, _p(qkv), H, C3, D, _p(dqkv), H, C3, D, _p(qg), _p(stats[0]), _p(stats[1]),
                                      _p(qgb), _p(qbb), N * H, D, 0, st)
            L.lotus_layernorm_bwd_f32(_p(dkh), 1, D, 0, _po(qkv, C * 4), H, C3, D, _po(dqkv, C * 4), H, C3, D, _p(kg),
                                      _p(stats[2]), _p(stats[3]), _p(kgb), _p(kbb), N * H, D, 0, st)
        return dqkv, dqg, dqb, dkg, dkb, drpe, None, None, None, None, None, None


def patch_attention(qkv, qn, kn, rpe, plan, H, qk_norm, attn_drop=NO_DROP, grid=None):
    qg = qn.weight if qk_norm else None
    qb = qn.bias if qk_norm else None
    kg = kn.weight if qk_norm else None
    kb = kn.bias if qk_norm else None
    table = rpe.rpe_table if rpe is not None else None
    pos_bnd = rpe.pos_bnd if rpe is not None else 0
    return PatchAttnFn.apply(qkv, qg, qb, kg, kb, table, plan, H, qk_norm, attn_drop, grid, pos_bnd)


class CrossAttnFn(torch.autograd.Function):
    """q (N,C) raw, kv (Lt,2C) raw -> per-head LN on q,k -> each cloud's points attend to its tokens -> (N,C)"""

    @staticmethod
    def forward(ctx, q, kv, qg, qb, kg, kb, geo, H, qk_norm, attn_drop):
        L = load()
        q, kv = _f32c(q), _f32c(kv)
        N, C = q.shape
        Lt = kv.shape[0]
        D = C // H
        dev = q.device
        st = _st()
        if qk_norm:
            qh = torch.empty_like(q)
            kh = torch.empty((Lt, C), device=dev, dtype=torch.float32)
            qs = torch.empty((2, N * H), device=dev, dtype=torch.float32)
            ks = torch.empty((2, Lt * H), device=dev, dtype=torch.float32)
            L.lotus_layernorm_fwd_f32(_p(q), 1, D, 0, _p(qh), 1, D, 0, None, _p(qg), _p(qb), _p(qs[0]), _p(qs[1]), N * H, D, 1e-6, st)
            L.lotus_layernorm_fwd_f32(_p(kv), H, 2 * C, D, _p(kh), 1, D, 0, None, _p(kg), _p(kb), _p(ks[0]), _p(ks[1]), Lt * H, D,
                                      1e-6, st)
            q_ptr, k_ptr, ldk = _p(qh), _p(kh), C
        else:
            qh = kh = qs = ks = None
            q_ptr, k_ptr, ldk = _p(q), _p(kv), 2 * C
        out = torch.empty_like(q)
        lse = torch.empty((N * H,), device=dev, dtype=torch.float32)
        L.lotus_cross_attn_fwd_f32(q_ptr, C, k_ptr, _po(kv, C * 4), ldk, 2 * C, _p(geo.pt_off), _p(geo.tok_off), geo.B,
                                   geo.max_points, geo.max_tokens, H, D, float(D) ** -0.5, _p(out), C, _p(lse),
                                   attn_drop[0], attn_drop[1], attn_drop[2], st)
        ctx.save_for_backward(q, kv, qg, kg, qh, kh, qs, ks, out, lse)
        ctx.geo, ctx.H, ctx.qk_norm, ctx.attn_drop = geo, H, qk_norm, attn_drop
        ctx.nb = (qb, kb)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = load()
        q, kv, qg, kg, qh, kh, qs, ks, out, lse = ctx.saved_tensors
        geo, H = ctx.geo, ctx.H
        dout = _f32c(dout)
        N, C = q.shape
        Lt = kv.shape[0]
        D = C // H
        dev = q.device
        st = _st()
        dq = torch.empty_like(q)
        dkv = torch.zeros_like(kv)
        if ctx.qk_norm:
            dqh = torch.empty_like(q)
            dkh = torch.zeros((Lt, C), device=dev, dtype=torch.float32)
            q_ptr, k_ptr, ldk = _p(qh), _p(kh), C
            dq_ptr, dk_ptr, lddk = _p(dqh), _p(dkh), C
        else:
            q_ptr, k_ptr, ldk = _p(q), _p(kv), 2 * C
            dq_ptr, dk_ptr, lddk = _p(dq), _p(dkv), 2 * C
        L.lotus_cross_attn_bwd_f32(q_ptr, C, k_ptr, _po(kv, C * 4), ldk, 2 * C, _p(geo.pt_off), _p(geo.tok_off), geo.B,
                                   geo.max_points, geo.max_tokens, H, D, float(D) ** -0.5, _p(out), _p(dout), C, _p(lse),
                                   dq_ptr, C, dk_ptr, _po(dkv, C * 4), lddk, 2 * C, ctx.attn_drop[0], ctx.attn_drop[1],
                                   ctx.attn_drop[2], st)
        dqg = dqb = dkg = dkb = None
        if ctx.qk_norm:
            (qgb, dqg), (qbb, dqb) = _grad_buf(qg, True), _grad_buf(ctx.nb[0], True)
            (kgb, dkg), (kbb, dkb) = _grad_buf(kg, True), _grad_buf(ctx.nb[1], True)
            L.lotus_layernorm_bwd_f32(_p(dqh), 1, D, 0, _p(q), 1, D, 0, _p(dq), 1, D, 0, _p(qg), _p(qs[0]), _p(qs[1]), _p(qgb),
                                      _p(qbb), N * H, D, 0, st)
            L.lotus_layernorm_bwd_f32(_p(dkh), 1, D, 0, _p(kv), H, 2 * C, D, _p(dkv), H, 2 * C, D, _p(kg), _p(ks[0]), _p(ks[1]),
                                      _p(kgb), _p(kbb), Lt * H, D, 0, st)
        return dq, dkv, dqg, dqb, dkg, dkb, None, None, None, None


def cross_attention(q, kv, qn, kn, geo, H, qk_norm, attn_drop=NO_DROP):
    qg = qn.weight if qk_norm else None
    qb = qn.bias if qk_norm else None
    kg = kn.weight if qk_norm else None
    kb = kn.bias if qk_norm else None
    return CrossAttnFn.apply(q, kv, qg, qb, kg, kb, geo, H, qk_norm, attn_drop)


# ------------------------------------------------------------------------------------------ pooling


class SegmentMaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, seg_ptr, member, M):
        L = load()
        x = _f32c(x)
        C = x.shape[1]
        out = torch.empty((M, C), device=x.device, dtype=torch.float32)
        arg = torch.empty((M, C), device=x.device, dtype=torch.int32)
        L.lotus_segment_max_fwd_f32(_p(x), C, _p(seg_ptr), _p(member), M, _p(out), _p(arg), _st())
        ctx.save_for_backward(arg)
        ctx.shape = x.shape
        return out

    @staticmethod
    def backward(ctx, dy):
        L = load()
        (arg,) = ctx.saved_tensors
        dy = _f32c(dy)
        dx = torch.zeros(ctx.shape, device=dy.device, dtype=torch.float32)
        L.lotus_segment_max_bwd_f32(_p(dy), _p(arg), arg.shape[0], arg.shape[1], _p(dx), _st())
        return dx, None, None, None


def segment_max(x, seg_ptr, member, M):
    return SegmentMaxFn.apply(x, seg_ptr, member, M)


def segment_mean(x, seg_ptr, member, M):
    """no-grad segment mean (pooled coordinates)"""
    L = load()
    x = _f32c(x)
    out = torch.empty((M, x.shape[1]), device=x.device, dtype=torch.float32)
    L.lotus_segment_sum_f32(_p(x), x.shape[1], _p(seg_ptr), _p(member), M, 1, _p(out), _st())
    return out


class UnpoolAddFn(torch.autograd.Function):
    """out[p] = skip[p] + child[cluster[p]]"""

    @staticmethod
    def forward(ctx, skip, child, cluster, seg_ptr, member):
        L = load()
        skip, child = _f32c(skip), _f32c(child)
        n, C = skip.shape
        out = torch.empty_like(skip)
        L.lotus_unpool_add_f32(_p(skip), _p(child), _p(cluster), n, C, _p(out), _st())
        ctx.save_for_backward(seg_ptr, member)
        ctx.M = child.shape[0]
        return out

    @staticmethod
    def backward(ctx, dy):
        L = load()
        seg_ptr, member = ctx.saved_tensors
        dy = _f32c(dy)
        dchild = torch.empty((ctx.M, dy.shape[1]), device=dy.device, dtype=torch.float32)
        L.lotus_segment_sum_f32(_p(dy), dy.shape[1], _p(seg_ptr), _p(member), ctx.M, 0, _p(dchild), _st())
        return dy, dchild, None, None, None


def unpool_add(skip, child, cluster, seg_ptr, member):
    return UnpoolAddFn.apply(skip, child, cluster, seg_ptr, member)


# ------------------------------------------------------------------------------------------ loss / decode


class PolicyLossFn(torch.autograd.Function):
    """-> losses (4,) = [pos, rot, open, total]  (simple_policy_ptv3.py:308-373)"""

    @staticmethod
    def forward(ctx, heat, action, prob, cloud_off, gt, B, NB, rot_bins, pos_w, rot_w):
        L = load()
        heat, action = _f32c(heat), _f32c(action)
        losses = torch.empty(4, device=heat.device, dtype=torch.float32)
        L.lotus_policy_loss_f32(_p(heat), _p(prob), _p(cloud_off), B, NB, _p(action), action.shape[1], rot_bins, _p(gt),
                                gt.shape[1], pos_w, rot_w, _p(losses), None, None, None, _st())
        ctx.save_for_backward(heat, action, prob, cloud_off, gt)
        ctx.cfg = (B, NB, rot_bins, pos_w, rot_w)
        return losses

    @staticmethod
    def backward(ctx, g):
        L = load()
        heat, action, prob, cloud_off, gt = ctx.saved_tensors
        B, NB, rot_bins, pos_w, rot_w = ctx.cfg
        g = _f32c(g)
        dheat = torch.empty_like(heat)
        dact = torch.empty_like(action)
        scratch = torch.empty(4, device=heat.device, dtype=torch.float32)
        L.lotus_policy_loss_f32(_p(heat), _p(prob), _p(cloud_off), B, NB, _p(action), action.shape[1], rot_bins, _p(gt),
                                gt.shape[1], pos_w, rot_w, _p(scratch), _p(dheat), _p(dact), _p(g), _st())
        return dheat, dact, None, None, None, None, None, None, None, None


def policy_loss(heat, action, prob, cloud_off, gt, B, NB, rot_bins, pos_w, rot_w):
    return PolicyLossFn.apply(heat, action, prob, cloud_off, gt, B, NB, rot_bins, float(pos_w), float(rot_w))


def decode_actions(heat, xyz, cloud_off, B, NB, bin_size, action, rot_bins, resolution, gt_pos):
    L = load()
    out = torch.empty((B, 8), device=heat.device, dtype=torch.float64)
    xyz_ld = xyz.stride(0)
    L.lotus_decode_actions(_p(heat), _p(xyz), xyz_ld, _p(cloud_off), B, NB, float(bin_size), _p(action), action.shape[1],
                           rot_bins, float(resolution), _p(gt_pos), gt_pos.stride(0) if gt_pos is not None else 0, _p(out), _st())
    return out
