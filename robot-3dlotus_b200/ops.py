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


def _actc(t):
    """activation tensor: fp32 (validation mode) or bf16 (tensor-core mode), contiguous, on the GPU"""
    assert t.dtype in (torch.float32, torch.bfloat16) and t.is_cuda, (t.dtype, t.device)
    return t if t.is_contiguous() else t.contiguous()


def _sfx(t):
    return "bf16" if t.dtype == torch.bfloat16 else "f32"


def _es(t):
    return t.element_size()


# bf16 weight staging for the tensor-core path: parameter address -> (w_bf16, wt_bf16)
STAGED = {}


def stage_weight(L, w, st):
    """fp32 master weight -> bf16 copy + transposed bf16 copy ([B][D][A]) used by the dgrad kernels"""
    if w.ndim == 2:
        A, D, B = w.shape[0], 1, w.shape[1]
    else:
        A, D, B = w.shape[0], w.shape[1] * w.shape[2] * w.shape[3], w.shape[4]
    ent = STAGED.get(w.data_ptr())
    if ent is None or ent[0].numel() != w.numel():
        ent = (torch.empty(w.shape, device=w.device, dtype=torch.bfloat16),
               torch.empty((B, D, A), device=w.device, dtype=torch.bfloat16))
        STAGED[w.data_ptr()] = ent
    L.lotus_cast_transpose_bf16(_p(w), _p(ent[0]), _p(ent[1]), A, D, B, st)
    return ent


def to_bf16(x):
    L = load()
    x = _f32c(x)
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    L.lotus_convert(_p(x), _p(y), x.numel(), 0, _st())
    return y


def to_f32(x):
    L = load()
    y = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    x = x.contiguous()                 # (held in a local: a temporary would return its memory before the kernel is enqueued)
    L.lotus_convert(_p(x), _p(y), x.numel(), 1, _st())
    return y


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
# bumped whenever a kernel rewrites parameters through raw pointers (the fused AdamW): invalidates cached bf16 weight copies
PARAM_EPOCH = [0]


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
    """y = [residual +] drop(x @ w.T + b).  fp32 x -> SIMT validation kernel; bf16 x -> tcgen05 kernel (bf16 or fp32 out)"""

    @staticmethod
    def forward(ctx, x, w, b, residual, drop, out_fp32):
        L = load()
        x = _actc(x)
        M, K = x.shape
        N = w.shape[0]
        tc = x.dtype == torch.bfloat16
        if residual is not None:
            residual = _actc(residual)
        if tc:
            wb, _ = STAGED[w.data_ptr()]
            y = torch.empty((M, N), device=x.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
            L.lotus_linear_fwd_bf16(_p(x), _p(wb), _p(b), _p(residual), _p(y), None, M, N, K, 0, int(out_fp32), drop[0], drop[1],
                                    drop[2], _st())
        else:
            y = torch.empty((M, N), device=x.device, dtype=torch.float32)
            L.lotus_linear_fwd_f32(_p(x), _p(w), _p(b), _p(residual), _p(y), M, N, K, 0, None, drop[0], drop[1], drop[2], _st())
        ctx.save_for_backward(x, w)
        ctx.bias = b
        ctx.has_b, ctx.has_res, ctx.drop, ctx.tc = b is not None, residual is not None, drop, tc
        return y

    @staticmethod
    def backward(ctx, dy):
        L = load()
        x, w = ctx.saved_tensors
        M, K = x.shape
        N = w.shape[0]
        st = _st()
        dres = dy if ctx.has_res else None
        if ctx.tc and dy.dtype == torch.float32:
            dy = to_bf16(dy)
        dy = _actc(dy)
        sfx = _sfx(dy)
        if ctx.drop[0] > 0:
            dyd = torch.empty_like(dy)
            getattr(L, "lotus_dropout_" + sfx)(_p(dy), _p(dyd), dy.numel(), ctx.drop[0], ctx.drop[1], ctx.drop[2], st)
            dy = dyd
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if ctx.tc:
                L.lotus_linear_dgrad_bf16(_p(dy), _p(STAGED[w.data_ptr()][1]), None, _p(dx), M, N, K, 0, st)
            else:
                L.lotus_linear_dgrad_f32(_p(dy), _p(w), None, _p(dx), M, N, K, st)
        wbuf, dw = _grad_buf(w, ctx.needs_input_grad[1])
        bbuf, db = _grad_buf(ctx.bias, ctx.has_b and ctx.needs_input_grad[2])
        if ctx.tc and N % 8 == 0 and K % 8 == 0:
            if wbuf is not None or bbuf is not None:      # weight and bias gradient in one GEMM (db = dy^T 1)
                L.lotus_linear_wgrad_bias_bf16(_p(dy), _p(x), None if wbuf is None else _p(wbuf), None if bbuf is None else _p(bbuf),
                                               M, N, K, st)
        else:
            if wbuf is not None:
                if ctx.tc:
                    L.lotus_linear_wgrad_bf16(_p(dy), _p(x), _p(wbuf), M, N, K, st)
                else:
                    L.lotus_linear_wgrad_f32(_p(dy), _p(x), _p(wbuf), M, N, K, None, st)
            if bbuf is not None:
                getattr(L, "lotus_colsum_" + sfx)(_p(dy), M, N, _p(bbuf), st)
        return dx, dw, db, dres, None, None


def linear(x, w, b=None, residual=None, drop=NO_DROP, out_fp32=False):
    return LinearFn.apply(x, w, b, residual, drop, out_fp32)


class MlpFn(torch.autograd.Function):
    """y = residual + drop_out(fc2(drop_act(gelu(fc1(x)))))  (model.py:560-583 inside Block / CABlock), tensor-core mode only:
    GELU + dropout ride in fc1's GEMM epilogue (pre-activation kept through a second TMA store) and the GELU backward rides
    in fc2's data-gradient GEMM epilogue -- the same launches the native executor issues."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual, drop_act, drop_out):
        L = load()
        x, residual = _actc(x), _actc(residual)
        assert x.dtype == torch.bfloat16
        M, C = x.shape
        Hd = w1.shape[0]
        st = _st()
        u = torch.empty((M, Hd), device=x.device, dtype=torch.bfloat16)
        g = torch.empty_like(u)
        L.lotus_linear_fwd_bf16(_p(x), _p(STAGED[w1.data_ptr()][0]), _p(b1), None, _p(g), _p(u), M, Hd, C, 1, 0, drop_act[0],
                                drop_act[1], drop_act[2], st)
        y = torch.empty_like(x)
        L.lotus_linear_fwd_bf16(_p(g), _p(STAGED[w2.data_ptr()][0]), _p(b2), _p(residual), _p(y), None, M, C, Hd, 0, 0, drop_out[0],
                                drop_out[1], drop_out[2], st)
        ctx.save_for_backward(x, u, g, w1, w2)
        ctx.b1, ctx.b2, ctx.drop_act, ctx.drop_out = b1, b2, drop_act, drop_out
        return y

    @staticmethod
    def backward(ctx, dy):
        L = load()
        x, u, g, w1, w2 = ctx.saved_tensors
        M, C = x.shape
        Hd = w1.shape[0]
        st = _st()
        dres = dy
        dy = _actc(to_bf16(dy) if dy.dtype == torch.float32 else dy)
        if ctx.drop_out[0] > 0:
            dyd = torch.empty_like(dy)
            L.lotus_dropout_bf16(_p(dy), _p(dyd), dy.numel(), ctx.drop_out[0], ctx.drop_out[1], ctx.drop_out[2], st)
            dy = dyd
        w2buf, dw2 = _grad_buf(w2, ctx.needs_input_grad[3])
        b2buf, db2 = _grad_buf(ctx.b2, ctx.needs_input_grad[4])
        if w2buf is not None or b2buf is not None:
            L.lotus_linear_wgrad_bias_bf16(_p(dy), _p(g), None if w2buf is None else _p(w2buf), None if b2buf is None else _p(b2buf),
                                           M, C, Hd, st)
        du = torch.empty_like(u)
        L.lotus_linear_dgrad_act_bf16(_p(dy), _p(STAGED[w2.data_ptr()][1]), _p(u), _p(du), M, C, Hd, 1, ctx.drop_act[0],
                                      ctx.drop_act[1], ctx.drop_act[2], st)
        w1buf, dw1 = _grad_buf(w1, ctx.needs_input_grad[1])
        b1buf, db1 = _grad_buf(ctx.b1, ctx.needs_input_grad[2])
        if w1buf is not None or b1buf is not None:
            L.lotus_linear_wgrad_bias_bf16(_p(du), _p(x), None if w1buf is None else _p(w1buf), None if b1buf is None else _p(b1buf),
                                           M, Hd, C, st)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            L.lotus_linear_dgrad_bf16(_p(du), _p(STAGED[w1.data_ptr()][1]), None, _p(dx), M, Hd, C, 0, st)
        return dx, dw1, db1, dw2, db2, dres, None, None


def mlp(x, fc1, fc2, residual, drop_act=NO_DROP, drop_out=NO_DROP):
    """transformer MLP tail; bf16 activations take the fused path, fp32 the three validation-mode kernels"""
    if x.dtype == torch.bfloat16:
        return MlpFn.apply(x, fc1.weight, fc1.bias, fc2.weight, fc2.bias, residual, drop_act, drop_out)
    u = linear(x, fc1.weight, fc1.bias)
    g = act(u, "gelu", drop_act)
    return linear(g, fc2.weight, fc2.bias, residual=residual, drop=drop_out)


# ------------------------------------------------------------------------------------------ pointwise


class ActFn(torch.autograd.Function):
    """y = drop(act(x)); act: 'gelu' | 'leaky' (slope 0.02) | 'silu' (no dropout)"""

    @staticmethod
    def forward(ctx, x, kind, drop):
        L = load()
        x = _actc(x)
        y = torch.empty_like(x)
        op = {"gelu": 0, "leaky": 2, "silu": 7}[kind]
        getattr(L, "lotus_pointwise_" + _sfx(x))(_p(x), None, _p(y), x.numel(), op, 0.02, drop[0], drop[1], drop[2], _st())
        ctx.save_for_backward(x)
        ctx.op, ctx.drop = op, drop
        return y

    @staticmethod
    def backward(ctx, dy):
        L = load()
        (x,) = ctx.saved_tensors
        dy = _actc(dy)
        dx = torch.empty_like(x)
        getattr(L, "lotus_pointwise_" + _sfx(x))(_p(x), _p(dy), _p(dx), x.numel(), ctx.op + 1, 0.02, ctx.drop[0], ctx.drop[1],
                                                 ctx.drop[2], _st())
        return dx, None, None


def act(x, kind, drop=NO_DROP):
    return ActFn.apply(x, kind, drop)


# ------------------------------------------------------------------------------------------ norms


class LayerNormFn(torch.autograd.Function):
    """y = [residual +] LN(x) over the last dim"""

    @staticmethod
    def forward(ctx, x, g, b, eps, residual):
        L = load()
        x = _actc(x)
        rows, C = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        if residual is not None:
            residual = _actc(residual)
        getattr(L, "lotus_layernorm_fwd_" + _sfx(x))(_p(x), 1, C, 0, _p(y), 1, C, 0, _p(residual), _p(g), _p(b), _p(mean), _p(rstd),
                                                     rows, C, eps, _st())
        ctx.save_for_backward(x, g, mean, rstd)
        ctx.beta = b
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        L = load()
        x, g, mean, rstd = ctx.saved_tensors
        dy = _actc(dy)
        rows, C = x.shape
        dx = torch.empty_like(x)
        gbuf, dg = _grad_buf(g, True)
        bbuf, db = _grad_buf(ctx.beta, True)
        if x.dtype == torch.bfloat16:
            L.lotus_layernorm_bwd_bf16(_p(dy), int(dy.dtype == torch.float32), 1, C, 0, _p(x), 1, C, 0, _p(dx), 1, C, 0, _p(g),
                                       _p(mean), _p(rstd), _p(gbuf), _p(bbuf), rows, C, 0, _st())
        else:
            L.lotus_layernorm_bwd_f32(_p(dy), 1, C, 0, _p(x), 1, C, 0, _p(dx), 1, C, 0, _p(g), _p(mean), _p(rstd), _p(gbuf),
                                      _p(bbuf), rows, C, 0, _st())
        return dx, dg, db, None, (dy if ctx.has_res else None)


def layer_norm(x, g, b, eps=1e-5, residual=None):
    return LayerNormFn.apply(x, g, b, eps, residual)


class BatchNormActFn(torch.autograd.Function):
    """y = [gelu](BatchNorm1d(x)); updates running stats in place when training.  out_bf16: y (and dy) bf16 while x may
    be fp32 (the stem conv runs in fp32)"""

    @staticmethod
    def forward(ctx, x, g, b, run_mean, run_var, train, eps, momentum, gelu, out_bf16):
        L = load()
        x = _actc(x)
        rows, C = x.shape
        bf = out_bf16 or x.dtype == torch.bfloat16
        y = torch.empty((rows, C), device=x.device, dtype=torch.bfloat16 if bf else torch.float32)
        mean = torch.empty(C, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        scratch = torch.empty(2 * C, device=x.device, dtype=torch.float32)
        if bf:
            L.lotus_batchnorm_fwd_bf16(_p(x), int(x.dtype == torch.float32), _p(y), rows, C, _p(g), _p(b), _p(run_mean), _p(run_var),
                                       _p(mean), _p(rstd), _p(scratch), int(train), eps, momentum, int(gelu), _st())
        else:
            L.lotus_batchnorm_fwd_f32(_p(x), _p(y), rows, C, _p(g), _p(b), _p(run_mean), _p(run_var), _p(mean), _p(rstd),
                                      _p(scratch), int(train), eps, momentum, int(gelu), _st())
        ctx.save_for_backward(x, g, b, mean, rstd)
        ctx.train, ctx.gelu, ctx.bf = train, gelu, bf
        return y

    @staticmethod
    def backward(ctx, dy):
        L = load()
        x, g, b, mean, rstd = ctx.saved_tensors
        dy = _actc(dy)
        rows, C = x.shape
        dx = torch.empty_like(x)
        dg = torch.zeros(C, device=x.device, dtype=torch.float32)
        db = torch.zeros_like(dg)
        if ctx.bf:
            if dy.dtype == torch.float32:
                dy = to_bf16(dy)
            L.lotus_batchnorm_bwd_bf16(_p(dy), _p(x), int(x.dtype == torch.float32), rows, C, _p(g), _p(b), _p(mean), _p(rstd),
                                       _p(dx), _p(dg), _p(db), int(ctx.train), int(ctx.gelu), _st())
        else:
            L.lotus_batchnorm_bwd_f32(_p(dy), _p(x), rows, C, _p(g), _p(b), _p(mean), _p(rstd), _p(dx), _p(dg), _p(db),
                                      int(ctx.train), int(ctx.gelu), _st())
        for param, d in ((g, dg), (b, db)):
            s = _sink(param)
            if s is not None:
                L.lotus_pointwise_f32(_p(d), _p(s), _p(s), C, 4, 0.0, 0.0, 0, 0, _st())     # sink += batch sums
        if _sink(g) is not None:
            dg = db = None
        return dx, dg, db, None, None, None, None, None, None, None


def batch_norm_act(x, bn, train, gelu=True, out_bf16=False):
    if train:
        bn.num_batches_tracked.add_(1)
    return BatchNormActFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, train, bn.eps, bn.momentum, gelu, out_bf16)


# ------------------------------------------------------------------------------------------ sparse conv


class SubmConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, nbr, nbr_stride, rule=None):
        L = load()
        x = _actc(x)
        n, cin = x.shape
        cout = w.shape[0]
        k3 = w.shape[1] * w.shape[2] * w.shape[3]
        tc = x.dtype == torch.bfloat16
        y = torch.empty((n, cout), device=x.device, dtype=x.dtype)
        if tc:
            L.lotus_subm_conv_fwd_bf16(_p(x), _p(STAGED[w.data_ptr()][0]), _p(b), _p(y), _p(nbr), nbr_stride, n, cin, cout, k3, 0, _st())
        else:
            L.lotus_subm_conv_fwd_f32(_p(x), _p(w), _p(b), _p(y), _p(nbr), nbr_stride, n, cin, cout, k3, _st())
        ctx.save_for_backward(x, w, nbr)
        ctx.bias = b
        ctx.meta = (n, cin, cout, k3, nbr_stride, b is not None, tc)
        ctx.rule = rule
        return y

    @staticmethod
    def backward(ctx, dy):
        L = load()
        x, w, nbr = ctx.saved_tensors
        n, cin, cout, k3, nbr_stride, has_b, tc = ctx.meta
        dy = _actc(dy)
        st = _st()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if tc:
                L.lotus_subm_conv_dgrad_bf16(_p(dy), _p(STAGED[w.data_ptr()][1]), None, _p(dx), _p(nbr), nbr_stride, n, cin, cout, k3, 0, st)
            else:
                L.lotus_subm_conv_dgrad_f32(_p(dy), _p(w), None, _p(dx), _p(nbr), nbr_stride, n, cin, cout, k3, st)
        wbuf, dw = _grad_buf(w, ctx.needs_input_grad[1])
        if wbuf is not None:
            if tc and ctx.rule is not None and ctx.rule.pairs_out is not None:
                r = ctx.rule      # compact per-tap pair lists: contract only over existing (voxel, neighbour) pairs
                L.lotus_subm_conv_wgrad_pairs_bf16(_p(dy), _p(x), _p(wbuf), _p(r.pairs_out), _p(r.pairs_in), _p(r.pair_cnt),
                                                   r.stride_alloc, n, cin, cout, k3, st)
            elif tc:
                L.lotus_subm_conv_wgrad_bf16(_p(dy), _p(x), _p(wbuf), _p(nbr), nbr_stride, n, cin, cout, k3, st)
            else:
                L.lotus_subm_conv_wgrad_f32(_p(dy), _p(x), _p(wbuf), _p(nbr), nbr_stride, n, cin, cout, k3, st)
        bbuf, db = _grad_buf(ctx.bias, has_b and ctx.needs_input_grad[2])
        if bbuf is not None:
            getattr(L, "lotus_colsum_" + _sfx(dy))(_p(dy), n, cout, _p(bbuf), st)
        return dx, dw, db, None, None, None


def subm_conv(x, w, b, nbr, nbr_stride, rule=None):
    return SubmConvFn.apply(x, w, b, nbr, nbr_stride, rule)


# ------------------------------------------------------------------------------------------ attention


def _fused_attn_ok(qkv, plan, D):
    """bf16 activations, patches of at most 128 tokens, head dim a multiple of 8 up to 64 -> the fused tcgen05 kernels
    (LOTUS_ATTN_SIMT=1 keeps the SIMT kernels: A/B tests)"""
    import os
    return (qkv.dtype == torch.bfloat16 and plan.max_len <= 128 and D % 8 == 0 and 8 <= D <= 64 and qkv.shape[1] % 8 == 0
            and os.environ.get("LOTUS_ATTN_SIMT") is None)


class PatchAttnFn(torch.autograd.Function):
    """qkv (N,3C) raw projections -> per-head LN on q,k -> attention inside each serialized patch -> (N,C) in
    point order (gather by order[pad] and scatter by unpad[inverse] fused).  fp32 or bf16 activations."""

    @staticmethod
    def forward(ctx, qkv, qg, qb, kg, kb, rpe_table, plan, H, qk_norm, attn_drop, grid, pos_bnd):
        L = load()
        qkv = _actc(qkv)
        sfx, es = _sfx(qkv), _es(qkv)
        N, C3 = qkv.shape
        C = C3 // 3
        D = C // H
        dev = qkv.device
        st = _st()
        fused = _fused_attn_ok(qkv, plan, D)
        if fused:
            # bf16: gather + per-head q/k LayerNorm + QK^T + (RPE) softmax + PV + scatter in ONE tcgen05 kernel (attention_fused.cu)
            out = torch.empty((N, C), device=dev, dtype=qkv.dtype)
            lse = torch.empty((plan.n_pad * H,), device=dev, dtype=torch.float32)
            L.lotus_patch_attn_fused_fwd_bf16(
                _p(qkv), _po(qkv, C * es), _po(qkv, 2 * C * es), C3, C3, C3, _p(qg) if qk_norm else None, _p(qb) if qk_norm else None,
                _p(kg) if qk_norm else None, _p(kb) if qk_norm else None, _p(plan.gather_idx), _p(plan.slot_of_point), _p(plan.cu_seqlens),
                plan.n_patches, plan.max_len, H, D, float(D) ** -0.5, _p(rpe_table), _p(grid) if rpe_table is not None else None, pos_bnd,
                _p(out), C, _p(lse), attn_drop[0], attn_drop[1], attn_drop[2], st)
            ctx.save_for_backward(qkv, qg, kg, None, None, None, out, lse, rpe_table, grid if rpe_table is not None else None)
            ctx.plan, ctx.H, ctx.qk_norm, ctx.attn_drop, ctx.pos_bnd = plan, H, qk_norm, attn_drop, pos_bnd
            ctx.nb = (qb, kb)
            ctx.fused = True
            return out
        ctx.fused = False
        ln_fwd = getattr(L, "lotus_layernorm_fwd_" + sfx)
        if qk_norm:
            qh = torch.empty((N, C), device=dev, dtype=qkv.dtype)
            kh = torch.empty_like(qh)
            stats = torch.empty((4, N * H), device=dev, dtype=torch.float32)
            ln_fwd(_p(qkv), H, C3, D, _p(qh), 1, D, 0, None, _p(qg), _p(qb), _p(stats[0]), _p(stats[1]), N * H, D, 1e-6, st)
            ln_fwd(_po(qkv, C * es), H, C3, D, _p(kh), 1, D, 0, None, _p(kg), _p(kb), _p(stats[2]), _p(stats[3]), N * H, D, 1e-6, st)
            q_ptr, k_ptr, ldq = _p(qh), _p(kh), C
        else:
            qh = kh = stats = None
            q_ptr, k_ptr, ldq = _p(qkv), _po(qkv, C * es), C3
        out = torch.empty((N, C), device=dev, dtype=qkv.dtype)
        lse = torch.empty((plan.n_pad * H,), device=dev, dtype=torch.float32)
        getattr(L, "lotus_patch_attn_fwd_" + sfx)(
            q_ptr, k_ptr, _po(qkv, 2 * C * es), ldq, ldq, C3, _p(plan.gather_idx), _p(plan.slot_of_point), _p(plan.cu_seqlens),
            plan.n_patches, plan.max_len, H, D, float(D) ** -0.5, _p(rpe_table), _p(grid) if rpe_table is not None else None,
            pos_bnd, _p(out), C, _p(lse), attn_drop[0], attn_drop[1], attn_drop[2], st)
        ctx.save_for_backward(qkv, qg, kg, qh, kh, stats, out, lse, rpe_table, grid if rpe_table is not None else None)
        ctx.plan, ctx.H, ctx.qk_norm, ctx.attn_drop, ctx.pos_bnd = plan, H, qk_norm, attn_drop, pos_bnd
        ctx.nb = (qb, kb)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = load()
        qkv, qg, kg, qh, kh, stats, out, lse, rpe_table, grid = ctx.saved_tensors
        plan, H = ctx.plan, ctx.H
        dout = _actc(dout)
        bf = qkv.dtype == torch.bfloat16
        sfx, es = _sfx(qkv), _es(qkv)
        N, C3 = qkv.shape
        C = C3 // 3
        D = C // H
        dev = qkv.device
        st = _st()
        f32 = dict(device=dev, dtype=torch.float32)
        drpe = torch.zeros_like(rpe_table) if rpe_table is not None else None
        if ctx.fused:
            dqkv = torch.empty_like(qkv)
            alias = torch.empty((plan.n_pad, 2, C), device=dev, dtype=torch.bfloat16)
            dqg = dqb = dkg = dkb = None
            qgb = qbb = kgb = kbb = None
            if ctx.qk_norm:
                (qgb, dqg), (qbb, dqb) = _grad_buf(qg, True), _grad_buf(ctx.nb[0], True)
                (kgb, dkg), (kbb, dkb) = _grad_buf(kg, True), _grad_buf(ctx.nb[1], True)
            L.lotus_patch_attn_fused_bwd_bf16(
                _p(qkv), _po(qkv, C * es), _po(qkv, 2 * C * es), C3, C3, C3, _p(qg) if ctx.qk_norm else None,
                _p(ctx.nb[0]) if ctx.qk_norm else None, _p(kg) if ctx.qk_norm else None, _p(ctx.nb[1]) if ctx.qk_norm else None,
                _p(plan.gather_idx), _p(plan.slot_of_point), _p(plan.cu_seqlens), plan.n_patches, plan.n_pad, plan.max_len, H, D,
                float(D) ** -0.5, _p(rpe_table), _p(grid), ctx.pos_bnd, _p(drpe), _p(out), _p(dout), C, _p(lse), _p(dqkv),
                _po(dqkv, C * es), _po(dqkv, 2 * C * es), C3, C3, C3, _p(qgb), _p(qbb), _p(kgb), _p(kbb), _p(alias),
                ctx.attn_drop[0], ctx.attn_drop[1], ctx.attn_drop[2], st)
            return dqkv, dqg, dqb, dkg, dkb, drpe, None, None, None, None, None, None
        # fp32 gradient accumulators of the attention core (alias slots add a second contribution -> atomics)
        if bf or ctx.qk_norm:
            dqh, dkh = torch.zeros((N, C), **f32), torch.zeros((N, C), **f32)
        if bf:
            dvh = torch.zeros((N, C), **f32)
            dqkv = torch.empty_like(qkv)
            dv_ptr, lddv = _p(dvh), C
        else:
            dqkv = torch.zeros_like(qkv)
            dv_ptr, lddv = _po(dqkv, 2 * C * 4), C3
        if ctx.qk_norm:
            q_ptr, k_ptr, ldq = _p(qh), _p(kh), C
            dq_ptr, dk_ptr, lddq = _p(dqh), _p(dkh), C
        elif bf:
            q_ptr, k_ptr, ldq = _p(qkv), _po(qkv, C * es), C3
            dq_ptr, dk_ptr, lddq = _p(dqh), _p(dkh), C
        else:
            q_ptr, k_ptr, ldq = _p(qkv), _po(qkv, C * 4), C3
            dq_ptr, dk_ptr, lddq = _p(dqkv), _po(dqkv, C * 4), C3
        getattr(L, "lotus_patch_attn_bwd_" + sfx)(
            q_ptr, k_ptr, _po(qkv, 2 * C * es), ldq, ldq, C3, _p(plan.gather_idx), _p(plan.slot_of_point), _p(plan.cu_seqlens),
            plan.n_patches, plan.max_len, H, D, float(D) ** -0.5, _p(rpe_table), _p(grid), ctx.pos_bnd, _p(drpe), _p(out), _p(dout),
            C, _p(lse), dq_ptr, dk_ptr, dv_ptr, lddq, lddq, lddv, ctx.attn_drop[0], ctx.attn_drop[1], ctx.attn_drop[2], st)
        dqg = dqb = dkg = dkb = None
        if ctx.qk_norm:
            (qgb, dqg), (qbb, dqb) = _grad_buf(qg, True), _grad_buf(ctx.nb[0], True)
            (kgb, dkg), (kbb, dkb) = _grad_buf(kg, True), _grad_buf(ctx.nb[1], True)
            if bf:
                L.lotus_layernorm_bwd_bf16(_p(dqh), 1, 1, D, 0, _p(qkv), H, C3, D, _p(dqkv), H, C3, D, _p(qg), _p(stats[0]),
                                           _p(stats[1]), _p(qgb), _p(qbb), N * H, D, 0, st)
                L.lotus_layernorm_bwd_bf16(_p(dkh), 1, 1, D, 0, _po(qkv, C * es), H, C3, D, _po(dqkv, C * es), H, C3, D, _p(kg),
                                           _p(stats[2]), _p(stats[3]), _p(kgb), _p(kbb), N * H, D, 0, st)
            else:
                L.lotus_layernorm_bwd_f32(_p(dqh), 1, D, 0, _p(qkv), H, C3, D, _p(dqkv), H, C3, D, _p(qg), _p(stats[0]), _p(stats[1]),
                                          _p(qgb), _p(qbb), N * H, D, 0, st)
                L.lotus_layernorm_bwd_f32(_p(dkh), 1, D, 0, _po(qkv, C * 4), H, C3, D, _po(dqkv, C * 4), H, C3, D, _p(kg),
                                          _p(stats[2]), _p(stats[3]), _p(kgb), _p(kbb), N * H, D, 0, st)
        elif bf:
            L.lotus_convert_rows_f32_to_bf16(_p(dqh), C, _p(dqkv), C3, N, C, st)
            L.lotus_convert_rows_f32_to_bf16(_p(dkh), C, _po(dqkv, C * es), C3, N, C, st)
        if bf:
            L.lotus_convert_rows_f32_to_bf16(_p(dvh), C, _po(dqkv, 2 * C * es), C3, N, C, st)
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
        q, kv = _actc(q), _actc(kv)
        sfx, es = _sfx(q), _es(q)
        N, C = q.shape
        Lt = kv.shape[0]
        D = C // H
        dev = q.device
        st = _st()
        ln_fwd = getattr(L, "lotus_layernorm_fwd_" + sfx)
        if qk_norm:
            qh = torch.empty_like(q)
            kh = torch.empty((Lt, C), device=dev, dtype=q.dtype)
            qs = torch.empty((2, N * H), device=dev, dtype=torch.float32)
            ks = torch.empty((2, Lt * H), device=dev, dtype=torch.float32)
            ln_fwd(_p(q), 1, D, 0, _p(qh), 1, D, 0, None, _p(qg), _p(qb), _p(qs[0]), _p(qs[1]), N * H, D, 1e-6, st)
            ln_fwd(_p(kv), H, 2 * C, D, _p(kh), 1, D, 0, None, _p(kg), _p(kb), _p(ks[0]), _p(ks[1]), Lt * H, D, 1e-6, st)
            q_ptr, k_ptr, ldk = _p(qh), _p(kh), C
        else:
            qh = kh = qs = ks = None
            q_ptr, k_ptr, ldk = _p(q), _p(kv), 2 * C
        out = torch.empty_like(q)
        lse = torch.empty((N * H,), device=dev, dtype=torch.float32)
        getattr(L, "lotus_cross_attn_fwd_" + sfx)(q_ptr, C, k_ptr, _po(kv, C * es), ldk, 2 * C, _p(geo.pt_off), _p(geo.tok_off), geo.B,
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
        dout = _actc(dout)
        bf = q.dtype == torch.bfloat16
        sfx, es = _sfx(q), _es(q)
        N, C = q.shape
        Lt = kv.shape[0]
        D = C // H
        dev = q.device
        st = _st()
        f32 = dict(device=dev, dtype=torch.float32)
        dq = torch.empty_like(q)
        if bf or ctx.qk_norm:
            dqh, dkh = torch.empty((N, C), **f32), torch.zeros((Lt, C), **f32)
        if bf:
            dvh = torch.zeros((Lt, C), **f32)
            dkv = torch.empty_like(kv)
            dv_ptr, lddv = _p(dvh), C
        else:
            dkv = torch.zeros_like(kv)
            dv_ptr, lddv = _po(dkv, C * 4), 2 * C
        if ctx.qk_norm:
            q_ptr, k_ptr, ldk = _p(qh), _p(kh), C
            dq_ptr, dk_ptr, lddk = _p(dqh), _p(dkh), C
        elif bf:
            q_ptr, k_ptr, ldk = _p(q), _p(kv), 2 * C
            dq_ptr, dk_ptr, lddk = _p(dqh), _p(dkh), C
        else:
            q_ptr, k_ptr, ldk = _p(q), _p(kv), 2 * C
            dq_ptr, dk_ptr, lddk = _p(dq), _p(dkv), 2 * C
        getattr(L, "lotus_cross_attn_bwd_" + sfx)(q_ptr, C, k_ptr, _po(kv, C * es), ldk, 2 * C, _p(geo.pt_off), _p(geo.tok_off), geo.B,
                                                  geo.max_points, geo.max_tokens, H, D, float(D) ** -0.5, _p(out), _p(dout), C,
                                                  _p(lse), dq_ptr, C, dk_ptr, dv_ptr, lddk, lddv, ctx.attn_drop[0],
                                                  ctx.attn_drop[1], ctx.attn_drop[2], st)
        dqg = dqb = dkg = dkb = None
        if ctx.qk_norm:
            (qgb, dqg), (qbb, dqb) = _grad_buf(qg, True), _grad_buf(ctx.nb[0], True)
            (kgb, dkg), (kbb, dkb) = _grad_buf(kg, True), _grad_buf(ctx.nb[1], True)
            if bf:
                L.lotus_layernorm_bwd_bf16(_p(dqh), 1, 1, D, 0, _p(q), 1, D, 0, _p(dq), 1, D, 0, _p(qg), _p(qs[0]), _p(qs[1]), _p(qgb),
                                           _p(qbb), N * H, D, 0, st)
                L.lotus_layernorm_bwd_bf16(_p(dkh), 1, 1, D, 0, _p(kv), H, 2 * C, D, _p(dkv), H, 2 * C, D, _p(kg), _p(ks[0]),
                                           _p(ks[1]), _p(kgb), _p(kbb), Lt * H, D, 0, st)
            else:
                L.lotus_layernorm_bwd_f32(_p(dqh), 1, D, 0, _p(q), 1, D, 0, _p(dq), 1, D, 0, _p(qg), _p(qs[0]), _p(qs[1]), _p(qgb),
                                          _p(qbb), N * H, D, 0, st)
                L.lotus_layernorm_bwd_f32(_p(dkh), 1, D, 0, _p(kv), H, 2 * C, D, _p(dkv), H, 2 * C, D, _p(kg), _p(ks[0]), _p(ks[1]),
                                          _p(kgb), _p(kbb), Lt * H, D, 0, st)
        elif bf:
            L.lotus_convert_rows_f32_to_bf16(_p(dqh), C, _p(dq), C, N, C, st)
            L.lotus_convert_rows_f32_to_bf16(_p(dkh), C, _p(dkv), 2 * C, Lt, C, st)
        if bf:
            L.lotus_convert_rows_f32_to_bf16(_p(dvh), C, _po(dkv, C * es), 2 * C, Lt, C, st)
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
        x = _actc(x)
        C = x.shape[1]
        out = torch.empty((M, C), device=x.device, dtype=x.dtype)
        arg = torch.empty((M, C), device=x.device, dtype=torch.int32)
        getattr(L, "lotus_segment_max_fwd_" + _sfx(x))(_p(x), C, _p(seg_ptr), _p(member), M, _p(out), _p(arg), _st())
        ctx.save_for_backward(arg)
        ctx.shape, ctx.dtype = x.shape, x.dtype
        return out

    @staticmethod
    def backward(ctx, dy):
        L = load()
        (arg,) = ctx.saved_tensors
        dy = _actc(dy)
        if dy.dtype != ctx.dtype:
            dy = to_bf16(dy) if ctx.dtype == torch.bfloat16 else to_f32(dy)
        dx = torch.zeros(ctx.shape, device=dy.device, dtype=ctx.dtype)
        getattr(L, "lotus_segment_max_bwd_" + _sfx(dy))(_p(dy), _p(arg), arg.shape[0], arg.shape[1], _p(dx), 0, _st())
        return dx, None, None, None


def segment_max(x, seg_ptr, member, M):
    return SegmentMaxFn.apply(x, seg_ptr, member, M)


def segment_mean(x, seg_ptr, member, M):
    """no-grad segment mean (pooled coordinates, fp32)"""
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
        skip, child = _actc(skip), _actc(child)
        n, C = skip.shape
        out = torch.empty_like(skip)
        getattr(L, "lotus_unpool_add_" + _sfx(skip))(_p(skip), _p(child), _p(cluster), n, C, _p(out), _st())
        ctx.save_for_backward(seg_ptr, member)
        ctx.M = child.shape[0]
        return out

    @staticmethod
    def backward(ctx, dy):
        L = load()
        seg_ptr, member = ctx.saved_tensors
        dy = _actc(dy)
        dchild = torch.empty((ctx.M, dy.shape[1]), device=dy.device, dtype=dy.dtype)
        getattr(L, "lotus_segment_sum_" + _sfx(dy))(_p(dy), dy.shape[1], _p(seg_ptr), _p(member), ctx.M, 0, _p(dchild), _st())
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



# ------------------------------------------------------------------------------------------ variants (AdaNorm / planner)


class CastFn(torch.autograd.Function):
    """differentiable fp32 <-> bf16 conversion (the gradient is converted back)"""

    @staticmethod
    def forward(ctx, x, to_bf):
        ctx.to_bf = to_bf
        return to_bf16(x) if to_bf else to_f32(x)

    @staticmethod
    def backward(ctx, dy):
        return (to_f32(dy) if ctx.to_bf else to_bf16(dy.float() if dy.dtype != torch.float32 else dy)), None


def cast(x, dtype):
    if x.dtype == dtype:
        return x
    return CastFn.apply(x, dtype == torch.bfloat16)


class AddFn(torch.autograd.Function):
    """y = a + b (same shape / dtype), through the pointwise kernel"""

    @staticmethod
    def forward(ctx, a, b):
        L = load()
        a, b = _actc(a), _actc(b)
        y = torch.empty_like(a)
        getattr(L, "lotus_pointwise_" + _sfx(a))(_p(b), _p(a), _p(y), a.numel(), 4, 0.0, 0.0, 0, 0, _st())
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


def add(a, b):
    return AddFn.apply(a, b)


class ModulateFn(torch.autograd.Function):
    """PDNorm modulation y = x * (1 + scale[b]) + shift[b] with per-cloud shift_scale (B, 2C) (model.py:293-303)"""

    @staticmethod
    def forward(ctx, x, ss, cloud_off, B):
        L = load()
        x, ss = _f32c(x), _f32c(ss)
        rows, C = x.shape
        assert ss.shape == (B, 2 * C)
        y = torch.empty_like(x)
        L.lotus_pdnorm_modulate_fwd_f32(_p(x), _p(ss), _p(cloud_off), B, rows, C, _p(y), _st())
        ctx.save_for_backward(x, ss, cloud_off)
        ctx.B = B
        return y

    @staticmethod
    def backward(ctx, dy):
        L = load()
        x, ss, cloud_off = ctx.saved_tensors
        dy = _f32c(dy)
        dx = torch.empty_like(x)
        dss = torch.empty_like(ss)
        L.lotus_pdnorm_modulate_bwd_f32(_p(dy), _p(x), _p(ss), _p(cloud_off), ctx.B, x.shape[1], _p(dx), _p(dss), _st())
        return dx, dss, None, None


def pdnorm_modulate(x, ss, cloud_off, B):
    return ModulateFn.apply(x, ss, cloud_off, B)


class TrajLossFn(torch.autograd.Function):
    """-> losses (5,) = [pos, rot, open, stop, total]  (motion_planner_ptv3.py:307-397)"""

    @staticmethod
    def forward(ctx, heat, action, prob, cloud_off, gt, stop, mask, B, T, NB, rot_bins, pos_w, rot_w):
        L = load()
        heat, action = _f32c(heat), _f32c(action)
        losses = torch.empty(5, device=heat.device, dtype=torch.float32)
        L.lotus_traj_loss_f32(_p(heat), _p(prob), _p(cloud_off), B, T, NB, _p(action), action.shape[1], rot_bins, _p(gt),
                              gt.shape[1], _p(stop), _p(mask), pos_w, rot_w, _p(losses), None, None, None, _st())
        ctx.save_for_backward(heat, action, prob, cloud_off, gt, stop, mask)
        ctx.cfg = (B, T, NB, rot_bins, pos_w, rot_w)
        return losses

    @staticmethod
    def backward(ctx, g):
        L = load()
        heat, action, prob, cloud_off, gt, stop, mask = ctx.saved_tensors
        B, T, NB, rot_bins, pos_w, rot_w = ctx.cfg
        g = _f32c(g)
        dheat = torch.empty_like(heat)
        dact = torch.empty_like(action)
        scratch = torch.empty(5, device=heat.device, dtype=torch.float32)
        L.lotus_traj_loss_f32(_p(heat), _p(prob), _p(cloud_off), B, T, NB, _p(action), action.shape[1], rot_bins, _p(gt),
                              gt.shape[1], _p(stop), _p(mask), pos_w, rot_w, _p(scratch), _p(dheat), _p(dact), _p(g), _st())
        return (dheat, dact) + (None,) * 11


def traj_loss(heat, action, prob, cloud_off, gt, stop, mask, B, T, NB, rot_bins, pos_w, rot_w):
    return TrajLossFn.apply(heat, action, prob, cloud_off, gt, stop, mask, B, T, NB, rot_bins, float(pos_w), float(rot_w))


def decode_trajs(heat, xyz, cloud_off, B, T, NB, bin_size, action, rot_bins, resolution, gt_pos):
    L = load()
    out = torch.empty((B, T, 9), device=heat.device, dtype=torch.float64)
    L.lotus_decode_trajs(_p(heat), _p(xyz), xyz.stride(0), _p(cloud_off), B, T, NB, float(bin_size), _p(action), action.shape[1],
                         rot_bins, float(resolution), _p(gt_pos), gt_pos.stride(0) if gt_pos is not None else 0, _p(out), _st())
    return out


def disc_pos_prob(pc_fts, cloud_off, B, max_cloud, gt_actions, pos_bins, bin_size, heatmap_type="dist"):
    """packed soft position labels (sum_b 3 * n_b * 2*pos_bins,) fp32 on the device, from the packed cloud and the ground-truth
    positions (utils/action_position_utils.py:7-46); replaces the host-computed `disc_pos_probs` of the collate"""
    L = load()
    assert pc_fts.dtype == torch.float32 and pc_fts.is_cuda and gt_actions.dtype == torch.float32
    N = pc_fts.shape[0]
    out = torch.empty(3 * 2 * pos_bins * N, device=pc_fts.device, dtype=torch.float32)
    scratch = torch.empty((int(L.cdll.lotus_disc_pos_prob_scratch_bytes(B)) + 7) // 8, device=pc_fts.device, dtype=torch.int64)
    L.lotus_disc_pos_prob(_p(pc_fts), pc_fts.stride(0), _p(cloud_off), B, max_cloud, _p(gt_actions), gt_actions.stride(0), pos_bins,
                          float(bin_size), int(heatmap_type == "plain"), _p(out), _p(scratch), _st())
    return out


# ---------------------------------------------------------------------------------------------------- batch producer (device side)
def subsample_keysteps(raw_off, counts, seed):
    """uniform random `counts[b]`-subset (random order) of every raw cloud: train/datasets/simple_policy_dataset.py:281-294.
    raw_off (B+1,) int32 device offsets of the raw clouds; counts: python list.  -> (sel_idx (sum counts,) int32, out_off (B+1,) int32)"""
    L = load()
    dev = raw_off.device
    B = raw_off.numel() - 1
    out_off = torch.tensor([0] + list(torch.tensor(counts).cumsum(0).tolist()), dtype=torch.int32, device=dev)
    sel = torch.empty(int(sum(counts)), dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    raw = raw_off.cpu()
    max_raw = int((raw[1:] - raw[:-1]).max()) if B else 0
    L.lotus_subsample_keysteps(_p(raw_off), _p(out_off), B, max_raw, int(seed) & (2 ** 63 - 1), _p(sel), _p(err), _st())
    if int(err.item()):
        raise RuntimeError("lotus_subsample_keysteps: a raw cloud exceeds 16384 points or a count exceeds its cloud")
    return sel, out_off


def augment_keysteps(raw_xyz, raw_rgb, sel_idx, out_off, angle, noise=None, noise_amp=0.002, seed=0, table_height=0.7505, center=True,
                     gt_pos=None, use_height=True):
    """the per-keystep augmentation + feature assembly of the reference's DataLoader workers on the device
    (train/datasets/simple_policy_dataset.py:158-181, 296, 321-346): -> (pc_fts (N, 7|6|4|3) f32, centroid (B,3) f64, gt_pos (B,3) f32 or None)"""
    L = load()
    dev = raw_xyz.device
    B = out_off.numel() - 1
    oc = out_off.cpu()
    N = int(oc[-1])
    max_cloud = int((oc[1:] - oc[:-1]).max()) if B else 0
    width = 3 + (3 if raw_rgb is not None else 0) + (1 if use_height else 0)
    fts = torch.empty((N, width), dtype=torch.float32, device=dev)
    centroid = torch.empty((B, 3), dtype=torch.float64, device=dev)
    gt_out = torch.empty((B, 3), dtype=torch.float32, device=dev) if gt_pos is not None else None
    scratch = torch.empty(3 * B, dtype=torch.float64, device=dev)
    # (every converted input is held in a local until the call is enqueued: a temporary would hand its memory back to the allocator)
    angle = angle.to(dev, torch.float64).contiguous()
    raw_xyz = raw_xyz.contiguous()
    raw_rgb = raw_rgb.contiguous() if raw_rgb is not None else None
    noise = noise.to(dev, torch.float64).contiguous() if noise is not None else None
    gt_pos = gt_pos.to(dev, torch.float64).contiguous() if gt_pos is not None else None
    assert raw_xyz.dtype == torch.float32 and (raw_rgb is None or raw_rgb.dtype == torch.uint8) and sel_idx is None or sel_idx.dtype == torch.int32
    L.lotus_augment_keysteps(_p(raw_xyz), _p(raw_rgb), _p(sel_idx), _p(out_off), B, max_cloud, _p(angle), _p(noise), float(noise_amp),
                             int(seed) & (2 ** 63 - 1), float(table_height), int(center), _p(gt_pos), _p(fts), width, _p(centroid),
                             _p(gt_out), 3, _p(scratch), _st())
    return fts, centroid, gt_out


def voxel_downsample(xyz, rgb, crop_lo, crop_hi, voxel_size=0.01):
    """workspace crop + 1 cm voxel down-sample of the merged camera clouds on the device (Actioner.process_point_clouds,
    evaluation/eval_simple_policy.py:151-175; open3d voxel_down_sample_and_trace restated): xyz (n,3) f32, rgb (n,3) u8 or None ->
    (voxel means (m,3) f64, colours (m,3) u8 of each voxel's first input point, first input index (m,) int32); voxels in order of first appearance"""
    L = load()
    dev = xyz.device
    xyz = xyz.reshape(-1, 3).contiguous()
    rgb = rgb.reshape(-1, 3).contiguous() if rgb is not None else None
    assert xyz.dtype == torch.float32 and (rgb is None or rgb.dtype == torch.uint8)
    n = xyz.shape[0]
    crop = torch.tensor(list(crop_lo) + list(crop_hi), dtype=torch.float64, device=dev)
    bounds = torch.empty(9, dtype=torch.float64, device=dev)
    mm = torch.empty(6, dtype=torch.int64, device=dev)
    st = _st()
    L.lotus_voxel_bounds(_p(xyz), n, _p(crop), float(voxel_size), _p(bounds), _p(mm), st)
    out_xyz = torch.empty((n, 3), dtype=torch.float64, device=dev)
    out_rgb = torch.empty((n, 3), dtype=torch.uint8, device=dev) if rgb is not None else None
    out_first = torch.empty(n, dtype=torch.int32, device=dev)
    meta = torch.zeros(4, dtype=torch.int32, device=dev)
    scratch = torch.empty((int(L.cdll.lotus_voxel_downsample_scratch_bytes(n)) + 7) // 8, dtype=torch.int64, device=dev)
    L.lotus_voxel_downsample(_p(xyz), _p(rgb), n, _p(bounds), float(voxel_size), _p(out_xyz), _p(out_rgb), _p(out_first), _p(meta), _p(scratch), st)
    m = int(meta[0].item())            # one size read-back (the reference holds the whole front end on the host)
    return out_xyz[:m], (out_rgb[:m] if rgb is not None else None), out_first[:m]
