"""Execution engine of the drop-in policy: walks the parameter tree of `policy.py` / `modules.py` and issues
the C-ABI kernels (through `ops.py`) in the order of the reference forward
(reference: models/simple_policy_ptv3.py:225-306,403-431; models/PointTransformerV3/model_ca.py:383-412;
model.py:659-680,713-790,817-828,859-861).

No CPU / PyTorch-math fallback exists: a missing library or a non-CUDA device raises.
"""
import torch

from . import ops
from .lib import load
from .modules import PDNorm
from .tokenizer import tokenize, tokenize_launch, tokenize_finish, CrossGeo


class PolicyEngine:
    def __init__(self, model):
        self.m = model
        self.forced_perms = None       # tests: tensor (S,4) of recorded shuffle_orders permutations
        self.div_mode = 0              # 0: torch-CPU grid semantics (true division), 1: torch-CUDA (reciprocal multiply)
        self.seed_counter = 0
        self.last = {}                 # intermediates of the last forward when `self.record` is on
        self.record = False
        # "fp32": validation mode (SIMT kernels, 1e-4 parity with the reference); "bf16": tensor-core mode
        # (tcgen05 GEMMs / implicit-GEMM conv, bf16 activations, fp32 master weights, statistics and gradients)
        self.precision = "fp32"
        self.pos_heatmap_type = "dist"  # TRAIN_DATASET.pos_heatmap_type of the job scripts ('plain' | 'dist')
        self.use_native = False        # one-call C++ executor (csrc/engine_native.cu) for the production topology
        self._native = None
        self._pf_stream = None
        self._prefetched = None
        self._tc_weights = None
        self.use_graph = False         # inference: replay the whole forward as one CUDA graph per size bucket (native.GraphedForward)
        self._graphs = {}
        self.graph_stats = {"replays": 0, "captures": 0, "fallbacks": 0}
        load()                         # fail loudly if the CUDA library is absent

    def _stage_weights(self):
        """bf16 (+ transposed) copies of every weight the tensor-core kernels read, refreshed each forward"""
        import torch.nn as nn
        from .modules import SubMConv3d
        if self._tc_weights is None:
            ws = []
            head = self.m.act_proj_head
            skip = {id(head.heatmap_mlp[3]), id(head.action_mlp[3]), id(self.m.ptv3_model.embedding.stem.conv)}
            for mod in self.m.modules():
                if id(mod) in skip:
                    continue
                if isinstance(mod, nn.Linear) and (mod.in_features % 8 or mod.out_features % 8):
                    continue                 # pose / txt-attention projections stay fp32 SIMT
                if isinstance(mod, (nn.Linear, SubMConv3d)):
                    ws.append(mod.weight)
            self._tc_weights = ws
        L = load()
        st = ops._st()
        for w in self._tc_weights:
            ops.stage_weight(L, w, st)

    # ------------------------------------------------------------------ helpers
    def _rec(self, name, t):
        if self.record:
            self.last[name] = t

    def _perms(self, S, R, shuffle):
        if self.forced_perms is not None:
            return [[int(v) for v in p] for p in self.forced_perms]
        if not shuffle:
            return None
        return [torch.randperm(R).tolist() for _ in range(S)]      # CPU RNG, like model.py:131,751

    # ------------------------------------------------------------------ norms (plain or PDNorm, model.py:257-304)
    def _modulated(self, mod, y, pd, stg):
        if not mod.adaptive:
            return y
        ss = ops.linear(pd["sctx"], mod.modulation[1].weight, mod.modulation[1].bias)       # (B, 2C) = [shift | scale]
        return ops.pdnorm_modulate(y, ss, stg.cloud_off, pd["B"])

    def _ln(self, mod, x, pd, stg, residual=None):
        if isinstance(mod, PDNorm):
            assert mod.norm.weight is not None, "pdnorm_affine False is not used by any shipped config"
            y = self._modulated(mod, ops.layer_norm(x, mod.norm.weight, mod.norm.bias, mod.norm.eps), pd, stg)
            return y if residual is None else ops.add(residual, y)
        return ops.layer_norm(x, mod.weight, mod.bias, mod.eps, residual=residual)

    def _bn_gelu(self, mod, x, train, pd, stg, out_bf16=False):
        if isinstance(mod, PDNorm):
            assert mod.norm.weight is not None, "pdnorm_affine False is not used by any shipped config"
            y = self._modulated(mod, ops.batch_norm_act(x, mod.norm, train, gelu=False), pd, stg)
            return ops.act(y, "gelu")
        return ops.batch_norm_act(x, mod, train, gelu=True, out_bf16=out_bf16)

    # ------------------------------------------------------------------ blocks
    def _block(self, blk, x, stg, drop, conv_in=None, pd=None):
        if blk.drop_path_p > 0 and drop.enabled:
            raise NotImplementedError("engine: stochastic depth (drop_path > 0) in training mode is not wired "
                                      "(0 in the production configuration); use eval mode or drop_path 0")
        a = blk.attn
        c = ops.subm_conv(x if conv_in is None else conv_in, blk.cpe[0].weight, blk.cpe[0].bias, stg.nbr3, stg.stride_alloc, rule=stg)
        c = ops.linear(c, blk.cpe[1].weight, blk.cpe[1].bias)
        x = self._ln(blk.cpe[2], c, pd, stg, residual=x)
        h = self._ln(blk.norm1[0], x, pd, stg)
        qkv = ops.linear(h, a.qkv.weight, a.qkv.bias)
        plan = stg.plans[a.order_index]
        o = ops.patch_attention(qkv, a.q_norm, a.k_norm, a.rpe, plan, a.num_heads, a.qk_norm, drop.site(a.attn_drop),
                                grid=stg.grid)
        x = ops.linear(o, a.proj.weight, a.proj.bias, residual=x, drop=drop.site(a.proj_drop_p))
        h = self._ln(blk.norm2[0], x, pd, stg)
        mlp = blk.mlp[0]
        return ops.mlp(h, mlp.fc1, mlp.fc2, x, drop.site(blk.proj_drop), drop.site(blk.proj_drop))

    def _ca_block(self, blk, x, ctx, geo, drop):
        a = blk.attn
        h = ops.layer_norm(x, blk.norm1[0].weight, blk.norm1[0].bias, blk.norm1[0].eps)
        q = ops.linear(h, a.q.weight, a.q.bias)
        kv = ops.linear(ctx, a.kv.weight, a.kv.bias)
        o = ops.cross_attention(q, kv, a.q_norm, a.k_norm, geo, a.num_heads, a.qk_norm, drop.site(a.attn_drop))
        x = ops.linear(o, a.proj.weight, a.proj.bias, residual=x, drop=drop.site(a.proj_drop_p))
        h = ops.layer_norm(x, blk.norm2[0].weight, blk.norm2[0].bias, blk.norm2[0].eps)
        mlp = blk.mlp[0]
        return ops.mlp(h, mlp.fc1, mlp.fc2, x, drop.site(blk.proj_drop), drop.site(blk.proj_drop))

    # ------------------------------------------------------------------ backbone
    def backbone(self, feat, coord, npoints, ctx, tok_off_host, train, drop, pd=None):
        """PointTransformerV3[CA].forward(return_dec_layers=True) (model.py:1072-1101, model_ca.py:383-412):
        returns (last decoder feat, stages).  `pd` carries the per-cloud context of the PDNorm (AdaNorm) variants."""
        pt = self.m.ptv3_model
        ca = pt.with_ca
        if (pt.pdnorm_bn or pt.pdnorm_ln) and pt.pdnorm_adaptive and pd is None:
            raise ValueError("adaptive PDNorm needs a per-cloud context")
        S = pt.num_stages
        R = len(pt.order)
        rows = [sorted({i % R for i in range(pt.enc_depths[s])} |
                       ({i % R for i in range(pt.dec_depths[s])} if s < S - 1 else set())) for s in range(S)]
        psz = [pt.enc_patch_size[s] for s in range(S)]
        for s in range(S - 1):
            assert pt.dec_patch_size[s] == psz[s], "enc/dec patch sizes of a stage must match (shared patch plan)"
        stages = tokenize(coord, npoints, pt.order, pt.stride, psz, rows, self._perms(S, R, pt.shuffle_orders),
                          self.m.config.action_config.voxel_size, self.div_mode, need_pairs=torch.is_grad_enabled())
        self.last_stages = stages
        dev = feat.device
        B = len(npoints)
        geos = None
        if ca:
            tok_off = torch.tensor(tok_off_host, dtype=torch.int32).to(dev, non_blocking=True)
            max_tok = max(b - a for a, b in zip(tok_off_host[:-1], tok_off_host[1:]))
            geos = [CrossGeo(stg.cloud_off, tok_off, B, stg.max_cloud, max_tok) for stg in stages]

        stem = pt.embedding.stem
        x = ops.subm_conv(feat, stem.conv.weight, None, stages[0].nbr5, stages[0].stride_alloc)     # narrow Cin: fp32 SIMT in both modes
        x = self._bn_gelu(stem.norm, x, train, pd, stages[0], out_bf16=(self.precision == "bf16"))
        self._rec("embedding", x)
        skips = []
        for s in range(S):
            enc = getattr(pt.enc, f"enc{s}")
            stg = stages[s]
            if s > 0:
                down = enc.down
                skips.append(x)
                proj = ops.linear(x, down.proj.weight, down.proj.bias)
                pooled = ops.segment_max(proj, stg.seg_ptr, stg.member, stg.n)
                x = self._bn_gelu(down.norm[0], pooled, train, pd, stg)
                self._rec(f"enc{s}.down", x)
            for i in range(pt.enc_depths[s]):
                x = self._block(getattr(enc, f"block{i}"), x, stg, drop, pd=pd)
                self._rec(f"enc{s}.block{i}", x)
                if ca and hasattr(enc, f"ca_block{i}"):
                    x = self._ca_block(getattr(enc, f"ca_block{i}"), x, ctx, geos[s], drop)
                    self._rec(f"enc{s}.ca_block{i}", x)
        for s in reversed(range(S - 1)):
            dec = getattr(pt.dec, f"dec{s}")
            stg, child = stages[s], stages[s + 1]
            up = dec.up
            a = self._bn_gelu(up.proj[1], ops.linear(x, up.proj[0].weight, up.proj[0].bias), train, pd, child)
            b = self._bn_gelu(up.proj_skip[1], ops.linear(skips[s], up.proj_skip[0].weight, up.proj_skip[0].bias), train, pd, stg)
            x = ops.unpool_add(b, a, child.cluster, child.seg_ptr, child.member)
            self._rec(f"dec{s}.up", x)
            conv_in = b          # sparse_conv_feat is NOT refreshed by the add (model.py:824 vs :235-238)
            for i in range(pt.dec_depths[s]):
                x = self._block(getattr(dec, f"block{i}"), x, stg, drop, conv_in=conv_in, pd=pd)
                conv_in = None
                self._rec(f"dec{s}.block{i}", x)
                if ca:
                    x = self._ca_block(getattr(dec, f"ca_block{i}"), x, ctx, geos[s], drop)
                    self._rec(f"dec{s}.ca_block{i}", x)
        return x, stages

    # ------------------------------------------------------------------ native one-call step
    def _launch_tokens(self, fts, npoints, need_pairs):
        pt = self.m.ptv3_model
        S, R = pt.num_stages, len(pt.order)
        return tokenize_launch(fts[:, :3], npoints, pt.order, pt.stride, list(pt.enc_patch_size), [[0]] * S,
                               self._perms(S, R, pt.shuffle_orders), self.m.config.action_config.voxel_size, self.div_mode,
                               need_pairs=need_pairs)

    def prefetch(self, batch, need_pairs=True):
        """Input pipelining for the native step: copy the NEXT batch to the device and run its integer tokeniser on a side stream
        while the current step executes (the reference overlaps its DataLoader / pin_memory copies the same way).  The next
        `native_step(batch)` with this very batch object picks the result up instead of tokenising synchronously."""
        if self._pf_stream is None:
            self._pf_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        key = batch["pc_fts"]                        # the tensor object itself (kept alive: an id() could be reused after a free)
        with torch.cuda.stream(self._pf_stream):
            dev_batch = self.m.prepare_batch(dict(batch))
            fts = dev_batch["pc_fts"]
            npoints = [int(v) for v in dev_batch["npoints_in_batch"]]
            pend = self._launch_tokens(fts, npoints, need_pairs)
        for t in pend.tensors + [v for v in dev_batch.values() if isinstance(v, torch.Tensor)]:
            t.record_stream(main)                    # allocated on the side stream, consumed on the main one
        self._prefetched = (key, dev_batch, pend, need_pairs)

    def take_prefetched(self, batch):
        """-> (device batch, pending tokens) if `prefetch` was called with this batch object, else None"""
        pf, self._prefetched = self._prefetched, None
        if pf is not None and pf[0] is batch["pc_fts"]:
            return pf[1], pf[2], pf[3]
        return None

    def native_step(self, batch, compute_loss=True, with_grad=True, compute_final_action=False, dropout_seed=None, pending=None,
                    slice_events=None):
        """forward (+ loss + backward into the gradient sinks) through `lotus_policy_step`.  Returns (final, losses(4,)).
        `pending`: tokens of this batch already in flight (`prefetch`)."""
        from . import native
        m = self.m
        if self._native is None:
            if not native.supported(m):
                raise NotImplementedError("native executor covers the production topology only")
            self._native = native.NativeStep(m)
        pt = m.ptv3_model
        fts = batch["pc_fts"]
        if not fts.is_cuda:
            raise RuntimeError("lotus engine needs the model on a CUDA device (there is no CPU fallback)")
        npoints = [int(v) for v in batch["npoints_in_batch"]]
        S, R = pt.num_stages, len(pt.order)
        if pending is None:
            pending = self._launch_tokens(fts, npoints, bool(with_grad and compute_loss))
        else:
            torch.cuda.current_stream().wait_event(pending.event)     # side-stream copies + tokeniser -> visible to this stream
        stages = tokenize_finish(pending)
        self.last_stages = stages
        tok_off = [0]
        for l in batch["txt_lens"]:
            tok_off.append(tok_off[-1] + int(l))
        tok_dev = torch.tensor(tok_off, dtype=torch.int32).to(fts.device, non_blocking=True)
        self.seed_counter += 1
        seed = dropout_seed if dropout_seed is not None else int(torch.initial_seed()) * 1000003 + self.seed_counter
        return self._native.run(batch, stages, tok_dev, tok_off, m.training, self.precision == "bf16", compute_loss,
                                not compute_final_action, with_grad and compute_loss, seed, slice_events=slice_events)

    def _graphed_forward(self, batch):
        """inference through a captured CUDA graph of the bucket this batch falls in (points rounded up to 512, tokens to 16); returns
        the (B, 8) float64 actions on the host, or None (the caller runs the eager path) when a stage outgrew the bucket's bounds"""
        from . import native
        pt = self.m.ptv3_model
        npts = [int(v) for v in batch["npoints_in_batch"]]
        lens = [int(v) for v in batch["txt_lens"]]
        B = len(npts)
        key = (B, native._round_up(sum(npts), 512), native._round_up(sum(lens), 16), native._round_up(max(npts), 512), native._round_up(max(lens), 16))
        if key[3] > 16384 or key[4] > 128:
            return None                      # beyond the shared-memory sort / token-list limits of the static path
        perms = self._perms(pt.num_stages, len(pt.order), pt.shuffle_orders)
        g = self._graphs.get(key)
        if g is None or g.stale():
            g = native.GraphedForward(self.m, *key)
            g.capture(batch, perms)
            self._graphs[key] = g
            self.graph_stats["captures"] += 1
        out = g(batch, perms)
        if out is None:
            self.graph_stats["fallbacks"] += 1
        else:
            self.graph_stats["replays"] += 1
        return out

    # ------------------------------------------------------------------ context (prepare_ptv3_batch)
    def _pose_embedding(self, poses):
        """RobotPoseEmbedding.forward (models/base.py:61-78); the quaternion -> euler step runs in scipy on the host
        exactly as in the reference"""
        from scipy.spatial.transform import Rotation as R
        pe = self.m.pose_embedding
        poses = poses.float()
        e = torch.from_numpy(R.from_quat(poses[:, 3:7].detach().cpu().numpy()).as_euler("xyz")).float().to(poses.device)
        rot_in = torch.cat([torch.sin(e), torch.cos(e)], -1).contiguous()
        x = ops.linear(poses[:, :3].contiguous(), pe.pos_embedding.weight, pe.pos_embedding.bias)
        x = ops.add(x, ops.linear(rot_in, pe.rot_embedding.weight, pe.rot_embedding.bias))
        x = ops.add(x, pe.open_embedding.weight[poses[:, -1].long()])
        return ops.layer_norm(x, pe.layer_norm.weight, pe.layer_norm.bias, pe.layer_norm.eps)

    def _context(self, batch, bf):
        """-> (ctx, token offsets) for the CA variants (packed tokens; pose / step-id appended per cloud:
        simple_policy_ptv3.py:413-429) or (per-cloud vector, None) for AdaNorm / Concat (:194-223, :449-456)"""
        m = self.m
        ac = m.config.action_config
        txt = batch["txt_embeds"].float()
        lens = [int(l) for l in batch["txt_lens"]]
        ctx = ops.linear(ops.to_bf16(txt) if bf else txt, m.txt_fc.weight, m.txt_fc.bias)
        pose = self._pose_embedding(batch["ee_poses"]) if ac.use_ee_pose else None
        step = m.stepid_embedding.weight[batch["step_ids"]] if (ac.use_step_id and hasattr(m, "stepid_embedding")) else None
        if m.variant == "ca":
            if pose is not None or step is not None:
                toks = list(torch.split(ctx, lens))
                for extra in (pose, step):
                    if extra is not None:
                        e = ops.cast(extra, torch.bfloat16) if bf else extra
                        toks = [torch.cat([c, e[i:i + 1]], 0) for i, c in enumerate(toks)]
                        lens = [l + 1 for l in lens]
                ctx = torch.cat(toks, 0)
            tok_off = [0]
            for l in lens:
                tok_off.append(tok_off[-1] + l)
            return ctx, tok_off
        if ac.txt_reduce == "attn" and m.variant != "concat":
            w = ops.linear(txt, m.txt_attn_fc.weight, m.txt_attn_fc.bias)
            ctx = torch.stack([(torch.softmax(wi, 0) * ci).sum(0) for wi, ci in zip(torch.split(w, lens), torch.split(ctx, lens))], 0)
        elif any(l != 1 for l in lens):
            raise ValueError("AdaNorm / Concat conditioning takes one (pre-reduced) text vector per keystep unless txt_reduce='attn'")
        if pose is not None:
            ctx = ops.add(ctx, pose)
        if step is not None:
            ctx = ops.add(ctx, step)
        return ctx, None

    # ------------------------------------------------------------------ policy
    def forward(self, batch, compute_loss=False, **kwargs):
        m = self.m
        cfg = m.config
        ac = cfg.action_config
        planner = getattr(m, "planner", False)
        if ac.pos_pred_type != "heatmap_disc" or ac.rot_pred_type != "euler_disc" or ac.reduce != "max":
            raise NotImplementedError("engine: the shipped head (heatmap_disc / euler_disc / max) is the wired one")
        fts = batch["pc_fts"]
        if not fts.is_cuda:
            raise RuntimeError("lotus engine needs the model on a CUDA device (there is no CPU fallback)")
        if (self.use_native and self.use_graph and not torch.is_grad_enabled() and not compute_loss and not m.training
                and self.precision == "bf16" and kwargs.get("compute_final_action", True)):
            out = self._graphed_forward(batch)
            if out is not None:
                return out
        if self.use_native and not torch.is_grad_enabled():
            final, lv = self.native_step(batch, compute_loss=compute_loss, with_grad=False,
                                         compute_final_action=kwargs.get("compute_final_action", True))
            if not compute_loss:
                return final
            lv = lv.clone()
            return final, {"pos": lv[0], "rot": lv[1], "open": lv[2], "total": lv[3]}
        fts = fts.float().contiguous()
        npoints = [int(v) for v in batch["npoints_in_batch"]]
        B = len(npoints)
        train = m.training
        self.seed_counter += 1
        seed = kwargs.get("dropout_seed", int(torch.initial_seed()) * 1000003 + self.seed_counter)
        drop = ops.DropCtx(seed, enabled=train)
        if self.record:
            self.last = {}

        bf = self.precision == "bf16"
        if bf:
            widths = list(m.ptv3_model.enc_channels) + list(m.ptv3_model.dec_channels)
            if any(c % 64 for c in widths):
                raise ValueError(f"tensor-core (bf16) mode needs channel widths that are multiples of 64, got {widths}; "
                                 "use precision='fp32' for this configuration")
            if m.variant != "ca":
                raise ValueError("tensor-core (bf16) mode covers the cross-attention variants; AdaNorm / Concat run in precision='fp32'")
            self._stage_weights()
        ctx, tok_off = self._context(batch, bf)
        feat, pd = fts, None
        if planner:                                   # motion_planner_ptv3.py:445-446
            feat = torch.cat([fts, m.pc_label_embedding.weight[batch["pc_labels"]]], -1).contiguous()
        if m.variant == "concat":                     # simple_policy_ptv3.py:458-462
            rep = torch.repeat_interleave(ctx, torch.tensor(npoints, device=fts.device), 0)
            feat, ctx = torch.cat([feat, rep], -1).contiguous(), None
        elif m.variant == "adanorm":
            pd = {"sctx": ops.act(ctx.contiguous(), "silu"), "B": B}

        x, stages = self.backbone(feat, fts[:, :3], npoints, ctx, tok_off, train, drop, pd)
        self._rec("dec_feat", x)
        s0 = stages[0]
        head = m.act_proj_head
        NB = 2 * head.pos_bins
        final_action = kwargs.get("compute_final_action", True)
        lc = cfg.loss_config
        if planner:
            return self._planner_tail(batch, x, fts, s0, B, NB, drop, compute_loss, final_action, lc)

        # action head (simple_policy_ptv3.py:113-120,149-152)
        hm, am = head.heatmap_mlp, head.action_mlp
        # (bf16 mode: the first layers run on tensor cores and emit fp32; the 90/217-wide logits stay fp32 SIMT)
        h = ops.act(ops.linear(x, hm[0].weight, hm[0].bias, out_fp32=True), "leaky", drop.site(head.dropout))
        heat = ops.linear(h, hm[3].weight, hm[3].bias)                      # (N, 3*2*pos_bins): [n, c*NB + b]
        pc = ops.segment_max(x, s0.cloud_off, None, B)                     # per-cloud max over points
        h2 = ops.act(ops.linear(pc, am[0].weight, am[0].bias, out_fp32=True), "leaky", drop.site(head.dropout))
        action = ops.linear(h2, am[3].weight, am[3].bias)                  # (B, 72*3 + 1)
        self._rec("heat", heat)
        self._rec("action", action)

        gt = batch.get("gt_actions")
        with torch.no_grad():
            final = ops.decode_actions(heat, fts, s0.cloud_off, B, NB, ac.pos_bin_size, action, head.euler_bins,
                                       head.euler_resolution, None if final_action else gt.float())
        if not compute_loss:
            return final
        if "disc_pos_probs_packed" in batch:
            prob = batch["disc_pos_probs_packed"].to(fts.device, non_blocking=True).float()
        elif "disc_pos_probs" in batch:
            prob = torch.cat([p.reshape(-1) for p in batch["disc_pos_probs"]]).to(fts.device, non_blocking=True).float()
        else:       # device-side batch producer (utils/action_position_utils.py:7-46)
            prob = ops.disc_pos_prob(fts, s0.cloud_off, B, s0.max_cloud, gt.float().contiguous(), ac.pos_bins, ac.pos_bin_size,
                                     self.pos_heatmap_type)
        lv = ops.policy_loss(heat, action, prob, s0.cloud_off, gt.float().contiguous(), B, NB, head.euler_bins,
                             lc.pos_weight, lc.rot_weight)
        return final, {"pos": lv[0], "rot": lv[1], "open": lv[2], "total": lv[3]}

    def _planner_tail(self, batch, x, fts, s0, B, NB, drop, compute_loss, final_action, lc):
        """trajectory head, losses and decode of MotionPlannerPTV3* (motion_planner_ptv3.py:77-148, 250-397).
        The heads read [point feature | step embedding]; the first layers are split accordingly so the per-point GEMM
        runs once instead of once per trajectory step:  W [x | e_t] + b = W_x x + (W_e e_t + b)."""
        m = self.m
        ac = m.config.action_config
        head = m.act_proj_head
        T, C = head.max_traj_len, head.hidden_size
        assert head.traj_embedding is not None, "traj_embed_size 0 is unused by the shipped planner"
        hm, am = head.heatmap_mlp, head.action_mlp
        x = ops.cast(x, torch.float32)
        N = x.shape[0]
        te = head.traj_embedding.weight[:T]
        W0 = hm[0].weight
        h0 = ops.linear(x, W0[:, :C].contiguous(), None)                                   # (N, hidden)
        tb = ops.linear(te.contiguous(), W0[:, C:].contiguous(), hm[0].bias)               # (T, hidden)
        h = (h0.unsqueeze(1) + tb.unsqueeze(0)).reshape(N * T, -1)
        h = ops.act(h, "leaky", drop.site(head.dropout))
        heat = ops.linear(h, hm[3].weight, hm[3].bias).view(N, T * 3 * NB)                 # [n, (t*3 + c)*NB + b]
        pc = ops.segment_max(x, s0.cloud_off, None, B)                                     # max over points of [x | e_t] = [max x | e_t]
        pct = torch.cat([pc.unsqueeze(1).expand(B, T, C), te.unsqueeze(0).expand(B, T, -1)], -1).reshape(B * T, -1).contiguous()
        h2 = ops.act(ops.linear(pct, am[0].weight, am[0].bias), "leaky", drop.site(head.dropout))
        action = ops.linear(h2, am[3].weight, am[3].bias)                                  # (B*T, 72*3 + 2)
        self._rec("heat", heat)
        self._rec("action", action)
        gt = batch.get("gt_trajs")
        gt2 = gt.float().reshape(B * T, -1).contiguous() if gt is not None else None
        with torch.no_grad():
            final = ops.decode_trajs(heat, fts, s0.cloud_off, B, T, NB, ac.pos_bin_size, action, head.euler_bins,
                                     head.euler_resolution, None if final_action else gt2)
        if not compute_loss:
            return final
        prob = torch.cat([p.reshape(-1) for p in batch["gt_trajs_disc_pos_probs"]]).to(fts.device, non_blocking=True).float()
        stop = batch["gt_trajs_stop"].to(fts.device).float().reshape(-1).contiguous()
        mask = batch["traj_masks"].to(fts.device).float().reshape(-1).contiguous()
        lv = ops.traj_loss(heat, action, prob, s0.cloud_off, gt2, stop, mask, B, T, NB, head.euler_bins, lc.pos_weight, lc.rot_weight)
        return final, {"pos": lv[0], "rot": lv[1], "open": lv[2], "stop": lv[3], "total": lv[4]}
