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
