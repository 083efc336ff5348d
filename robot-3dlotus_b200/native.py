"""Python side of the native step executor (`lotus_policy_step`, csrc/engine_native.cu): packs the model's
parameter / gradient / BatchNorm pointers and the tokeniser's stage tables into plain C arrays and issues ONE
C-ABI call per forward(+backward).  Used by the trainer and by inference for the production topology; every other
configuration goes through the per-op Python engine.
"""
import ctypes

import torch

from .lib import load
from .ops import GRAD_SINK, PARAM_EPOCH

VP = ctypes.c_void_p


def supported(model):
    pt = model.ptv3_model
    ac = model.config.action_config
    return (model.variant == "ca" and pt.with_ca and not (pt.pdnorm_bn or pt.pdnorm_ln) and not pt.enable_rpe
            and all(d == 1 for d in pt.enc_depths) and all(d == 1 for d in pt.dec_depths)
            and not ac.use_ee_pose and not ac.use_step_id and ac.pos_pred_type == "heatmap_disc"
            and ac.rot_pred_type == "euler_disc" and ac.reduce == "max"
            and len(set(pt.enc_patch_size + pt.dec_patch_size)) == 1
            and _no_drop_path(pt))


def _no_drop_path(pt):
    """the C++ executor has no stochastic depth (model.py:659-680 DropPath; 0 in every shipped job script, 0.1 in the bare yaml)"""
    from .modules import Block
    return all(getattr(b, "drop_path_p", 0.0) == 0 for b in pt.modules() if isinstance(b, Block) or hasattr(b, "drop_path_p"))


def _f32_dev(t, dev, name):
    """-> contiguous float32 tensor on `dev` (no copy when it already is one)"""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"native step: batch['{name}'] must be a tensor, got {type(t).__name__}")
    if t.device != dev:
        t = t.to(dev, non_blocking=True)
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class NativeStep:
    def __init__(self, model):
        assert supported(model)
        self.m = model
        pt = model.ptv3_model
        ac = model.config.action_config
        S = pt.num_stages
        cfg = [S, pt.embedding.stem.conv.in_channels, ac.txt_ft_size, ac.context_channels, ac.pos_bins,
               model.act_proj_head.euler_bins]
        cfg += list(pt.enc_channels) + list(pt.enc_num_head) + list(pt.dec_channels[:S - 1]) + list(pt.dec_num_head)
        self.cfg = (ctypes.c_int64 * len(cfg))(*cfg)
        self.S, self.K, self.qk_norm = S, pt.enc_patch_size[0], int(pt.qk_norm)
        self.params = [p for _, p in model.named_parameters()]
        bns = [pt.embedding.stem.norm] + [getattr(pt.enc, f"enc{s}").down.norm[0] for s in range(1, S)]
        for s in reversed(range(S - 1)):
            up = getattr(pt.dec, f"dec{s}").up
            bns += [up.proj[1], up.proj_skip[1]]
        self.bns = bns
        self._ptr_key = None
        self._staged_key = None
        self.arena = None
        self.losses = None
        self.final = None

    def _pointer_arrays(self, with_grad):
        key = (tuple(p.data_ptr() for p in self.params[:4]), with_grad, len(GRAD_SINK))
        if key != self._ptr_key:
            n = len(self.params)
            self.c_params = (VP * n)(*[p.data_ptr() for p in self.params])
            sinks = [GRAD_SINK.get(p.data_ptr()) for p in self.params]
            if with_grad and any(s is None for s in sinks):
                # no trainer-owned flat buffer: give every parameter a .grad to accumulate into (reference-style usage)
                for p in self.params:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                sinks = [p.grad for p in self.params]
            self.sinks = sinks
            self.c_grads = (VP * n)(*[s.data_ptr() for s in sinks]) if with_grad else None
            stats = []
            for bn in self.bns:
                stats += [bn.running_mean.data_ptr(), bn.running_var.data_ptr()]
            self.c_bn = (VP * len(stats))(*stats)
            self._ptr_key = key
        return self.c_params, (self.c_grads if with_grad else None), self.c_bn

    def run(self, batch, stages, tok_off_dev, tok_off_host, train, bf16, compute_loss, use_gt_pos, with_grad, seed, slice_events=None):
        L = load()
        m = self.m
        ac = m.config.action_config
        lc = m.config.loss_config
        pt = m.ptv3_model
        dev = batch["pc_fts"].device
        B = len(batch["npoints_in_batch"])
        Lt = tok_off_host[-1]
        max_tok = max(b - a for a, b in zip(tok_off_host[:-1], tok_off_host[1:]))
        run = (ctypes.c_int64 * 11)(B, Lt, max_tok, self.K, self.qk_norm, int(train), int(bf16), int(compute_loss), int(use_gt_pos), 0, 0)
        sizes, ptrs = [], []
        for s, stg in enumerate(stages):
            pl = stg.plans[0]
            sizes += [stg.n, pl.n_pad, pl.n_patches, stg.max_cloud, stg.stride_alloc]
            ptrs += [stg.nbr3.data_ptr(), *(t.data_ptr() if t is not None else 0 for t in (stg.pairs_out, stg.pairs_in, stg.pair_cnt)),
                     pl.gather_idx.data_ptr(), pl.slot_of_point.data_ptr(), pl.cu_seqlens.data_ptr(), stg.cloud_off.data_ptr(),
                     stg.cluster.data_ptr() if stg.cluster is not None else 0, stg.seg_ptr.data_ptr() if stg.seg_ptr is not None else 0,
                     stg.member.data_ptr() if stg.member is not None else 0, stg.nbr5.data_ptr() if stg.nbr5 is not None else 0]
        c_sizes = (ctypes.c_int64 * len(sizes))(*sizes)
        c_ptrs = (VP * len(ptrs))(*ptrs)
        c_params, c_grads, c_bn = self._pointer_arrays(with_grad)
        fl = (ctypes.c_float * 6)(float(pt.enc.enc0.block0.attn.attn_drop), float(pt.enc.enc0.block0.proj_drop),
                                  float(m.act_proj_head.dropout), float(lc.pos_weight), float(lc.rot_weight),
                                  float(m.act_proj_head.euler_resolution))
        # the C ABI takes raw `const float*`: coerce what a reference collate may hand over (float64 soft labels from
        # numpy arithmetic, `action_position_utils.py:7-46`; non-contiguous slices) BEFORE any pointer is taken
        feat = _f32_dev(batch["pc_fts"], dev, "pc_fts")
        txt = _f32_dev(batch["txt_embeds"], dev, "txt_embeds")
        gt = batch.get("gt_actions")
        if gt is not None:
            gt = _f32_dev(gt, dev, "gt_actions")
        prob = batch.get("disc_pos_probs_packed")
        if prob is not None:
            prob = _f32_dev(prob, dev, "disc_pos_probs_packed")
        if compute_loss and prob is None:
            if "disc_pos_probs" in batch:
                prob = _f32_dev(torch.cat([p.reshape(-1) for p in batch["disc_pos_probs"]]), dev, "disc_pos_probs")
            else:   # device-side batch producer: soft labels from (xyz, gt position) instead of a host tensor
                from . import ops
                prob = ops.disc_pos_prob(feat, stages[0].cloud_off, B, stages[0].max_cloud, gt, ac.pos_bins, ac.pos_bin_size,
                                         getattr(m.engine(), "pos_heatmap_type", "dist"))
        if self.losses is None or self.losses.device != dev:
            self.losses = torch.zeros(4, device=dev, dtype=torch.float32)
        final = torch.empty((B, 8), device=dev, dtype=torch.float64)
        peak = ctypes.c_int64(0)
        st = VP(torch.cuda.current_stream().cuda_stream)
        common = (self.cfg, run, c_sizes, c_ptrs, c_params, c_grads, c_bn, VP(feat.data_ptr()), VP(txt.data_ptr()),
                  VP(tok_off_dev.data_ptr()), VP(gt.data_ptr()) if gt is not None else None,
                  VP(prob.data_ptr()) if prob is not None else None, fl, float(ac.pos_bin_size), int(seed) & (2 ** 63 - 1))
        # dry run sizes the arena (host-only, nothing is launched)
        run[9] = 1
        ev = None
        if slice_events is not None:
            # torch.cuda.Event creates its CUDA event lazily: record once so the handle exists, then hand the raw cudaEvent_t over
            for e in slice_events:
                if not getattr(e, "_lotus_ready", False):
                    e.record()
                    e._lotus_ready = True
            ev = (VP * len(slice_events))(*[e.cuda_event for e in slice_events])
        L.lotus_policy_step(*common, None, 0, ctypes.byref(peak), VP(self.losses.data_ptr()), VP(final.data_ptr()), None, st)
        need = int(peak.value) + 4096
        if self.arena is None or self.arena.numel() < need or self.arena.device != dev:
            self.arena = None
            self._staged_key = None
            self.arena = torch.empty(int(need * 1.1), device=dev, dtype=torch.uint8)
        run[9] = 0
        # inference: the staged bf16 weights at the head of the arena stay valid while the parameters are untouched
        wkey = None
        if not train and bf16:
            wkey = (self.arena.data_ptr(), PARAM_EPOCH[0], sum(p._version for p in self.params))
            run[10] = int(wkey == self._staged_key)
        L.lotus_policy_step(*common, VP(self.arena.data_ptr()), self.arena.numel(), ctypes.byref(peak), VP(self.losses.data_ptr()),
                            VP(final.data_ptr()), ev, st)
        self._staged_key = wkey
        if train:
            torch._foreach_add_([bn.num_batches_tracked for bn in self.bns], 1)
        return final, self.losses


def profile_step(run_step):
    """Run `run_step()` (one native step) in profile mode and return [(c_abi_name, args tuple, device_ms)], one entry per
    launcher call of `lotus_policy_step` in issue order.  Arguments are in C-ABI order: pointers as integer addresses
    (0 = null), integers as given, floats as 0 / 1 (zero / non-zero)."""
    L = load()
    L.cdll.lotus_policy_step_profile(1)
    try:
        run_step()
        torch.cuda.synchronize()
        L.cdll.lotus_policy_step_profile_count.restype = ctypes.c_int64
        n = int(L.cdll.lotus_policy_step_profile_count())
        name = ctypes.create_string_buffer(64)
        ms = ctypes.c_float(0)
        args = (ctypes.c_int64 * 48)()
        na = ctypes.c_int(0)
        out = []
        for i in range(n):
            rc = L.cdll.lotus_policy_step_profile_read(ctypes.c_int64(i), name, ctypes.byref(ms), args, ctypes.byref(na))
            if rc:
                raise RuntimeError(L.cdll.lotus_last_error().decode())
            out.append((name.value.decode(), tuple(int(args[j]) for j in range(na.value)), float(ms.value)))
        return out
    finally:
        L.cdll.lotus_policy_step_profile(0)


# ---------------------------------------------------------------------------------------------------- CUDA-graph replay (inference)
def _round_up(x, m):
    return (int(x) + m - 1) // m * m


class GraphedForward:
    """CUDA-graph capture of the WHOLE inference forward for one size bucket: integer tokeniser (all five stages) + lotus_policy_step
    (forward + final-action decode), ~370 kernel nodes replayed with one launch.  SURVEY §7.10 / §8f row 1 ("CUDA-graph capture with
    shape bucketing"); reference hot loop: evaluation/eval_simple_policy.py:316-321 (one model call per keystep).

    Shape bucketing: the graph is captured over STATIC buffers sized by the bucket (points rounded up to 512, text tokens to 16);
    the actual cloud / token offsets are device data.  Every kernel runs over the bucket's row bounds; rows between a stage's actual
    size and its bound are inert (lotus_stage_sanitize) and compute garbage no real row reads (BatchNorm uses running statistics in
    eval mode, so no kernel reduces over rows).  The deeper stages' bounds are fractions of the point bucket (RATIOS, ~1.5-7x the
    sizes measured on voxel1cm-like clouds); after every replay the actual sizes come back with the actions and an overflow falls
    back to the eager path.  The per-forward `shuffle_orders` draw stays a run-time input (curve ids in device memory)."""
    RATIOS = (1.0, 0.65, 0.32, 0.16, 0.09)

    def __init__(self, model, B, n_bucket, lt_bucket, cloud_bucket, tok_bucket):
        from . import tokenizer as tk
        self.tk, self.model, self.eng = tk, model, model.engine()
        self.ns = NativeStep(model)          # a private executor: its arena (and the bf16 weights staged in it) is baked into the graph
        self.tokb = tok_bucket
        pt = model.ptv3_model
        self.B, self.N0b, self.Ltb, self.S = B, n_bucket, lt_bucket, pt.num_stages
        dev = next(model.parameters()).device
        self.dev = dev
        inC = pt.embedding.stem.conv.in_channels
        ratios = list(self.RATIOS) + [self.RATIOS[-1]] * max(0, self.S - len(self.RATIOS))
        self.bounds = [max(B, min(n_bucket, int(n_bucket * ratios[s]) + B)) for s in range(self.S)]
        self.cloud_bounds = [max(1, min(self.bounds[s], int(cloud_bucket * ratios[s]) + 1)) for s in range(self.S)]
        self.fts = torch.zeros((n_bucket, inC), device=dev, dtype=torch.float32)
        self.txt = torch.zeros((lt_bucket, model.config.action_config.txt_ft_size), device=dev, dtype=torch.float32)
        self.off0 = torch.zeros(B + 1, device=dev, dtype=torch.int32)
        self.tok_off = torch.zeros(B + 1, device=dev, dtype=torch.int32)
        self.curves = torch.zeros((self.S, 4), device=dev, dtype=torch.int32)
        self.h_small = torch.zeros(2 * (B + 1) + 4 * self.S, dtype=torch.int32).pin_memory()
        self.graph, self.final, self.sizes, self.h_out = None, None, None, None
        self.param_epoch = None

    def _stage_inputs(self, batch, perms):
        """copy this call's inputs into the static buffers (stream-ordered H2D copies from the caller's tensors)"""
        B, S = self.B, self.S
        npts = [int(v) for v in batch["npoints_in_batch"]]
        lens = [int(v) for v in batch["txt_lens"]]
        N, Lt = sum(npts), sum(lens)
        h = self.h_small
        off = 0
        for i, c in enumerate(npts):
            h[i] = off
            off += c
        h[B] = off
        off = 0
        for i, l in enumerate(lens):
            h[B + 1 + i] = off
            off += l
        h[2 * B + 1] = off
        pt = self.model.ptv3_model
        curves = [self.tk.CURVE_ID[o] for o in pt.order]
        for s in range(S):
            if perms is not None:
                curves = [curves[int(i)] for i in perms[s]]
            for r in range(4):
                h[2 * (B + 1) + 4 * s + r] = curves[r] if r < len(curves) else 0
        self.off0.copy_(h[: B + 1], non_blocking=True)
        self.tok_off.copy_(h[B + 1: 2 * (B + 1)], non_blocking=True)
        self.curves.copy_(h[2 * (B + 1):].view(S, 4), non_blocking=True)
        self.fts[:N].copy_(batch["pc_fts"], non_blocking=True)
        self.txt[:Lt].copy_(batch["txt_embeds"], non_blocking=True)
        return npts, lens

    def _forward_static(self):
        """tokeniser + native step over the static buffers (this is what gets captured)"""
        m, pt, eng = self.model, self.model.ptv3_model, self.eng
        S, B = self.S, self.B
        static = {"off0": self.off0, "bounds": self.bounds, "cloud_bounds": self.cloud_bounds, "curves_dev": self.curves}
        pend = self.tk.tokenize_launch(self.fts[:, :3], [1] * B, pt.order, pt.stride, list(pt.enc_patch_size), [[0]] * S, None,
                                       m.config.action_config.voxel_size, eng.div_mode, need_pairs=False, static=static)
        stages = self.tk.tokenize_finish(pend)
        batch = {"pc_fts": self.fts, "txt_embeds": self.txt, "npoints_in_batch": [1] * B, "txt_lens": [1] * B}
        # host-side token geometry of the bucket: total = Ltb rows, longest list = tokb (the kernels read the real offsets on the device)
        tok_host = [min(i * self.tokb, self.Ltb) for i in range(B)] + [self.Ltb]
        final, _ = self.ns.run(batch, stages, self.tok_off, tok_host, False, True, False, False, False, 0)
        return final, pend.dev_sizes

    def capture(self, batch, perms):
        from .ops import PARAM_EPOCH
        self._stage_inputs(batch, perms)
        # eager warm-up over the static buffers: first-use initialisation inside the library (function attributes, side stream),
        # arena sizing and bf16 weight staging happen OUTSIDE the capture
        self._forward_static()
        self._forward_static()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.final, self.sizes = self._forward_static()
        self.h_out = torch.empty(self.final.numel() * 2 + self.sizes.numel(), dtype=torch.int32).pin_memory()
        self.param_epoch = (PARAM_EPOCH[0], sum(p._version for p in self.ns.params))

    def stale(self):
        from .ops import PARAM_EPOCH
        return self.param_epoch != (PARAM_EPOCH[0], sum(p._version for p in self.ns.params))

    def __call__(self, batch, perms):
        """-> final actions (B, 8) float64 on the HOST (pinned), or None when a stage outgrew its bound (caller falls back to eager)"""
        self._stage_inputs(batch, perms)
        self.graph.replay()
        nf = self.final.numel()
        out_f = self.h_out[: 2 * nf].view(torch.float64)
        out_f.copy_(self.final.reshape(-1), non_blocking=True)
        self.h_out[2 * nf:].copy_(self.sizes, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        S, B = self.S, self.B
        sizes = self.h_out[2 * nf:]
        meta = sizes[: S * 8].view(S, 8)
        if int(sizes[-1]) != 0:
            return None
        offs = sizes[S * 8: S * 8 + S * (B + 1)].view(S, B + 1)
        for s in range(S):
            cl = offs[s][1:] - offs[s][:-1]
            if int(meta[s][0]) > self.bounds[s] or int(cl.max()) > self.cloud_bounds[s]:
                return None
        return out_f.view(B, 8).clone()
