"""Execution engine of the drop-in policy: walks the parameter tree of `policy.py` / `modules.py` and issues
the C-ABI kernels (through `ops.py`) in the order of the reference forward
(reference: models/simple_policy_ptv3.py:225-306,403-431; models/PointTransformerV3/model_ca.py:383-412;
model.py:659-680,713-790,817-828,859-861).

No CPU / PyTorch-math fallback exists: a missing library or a non-CUDA device raises.
"""
import torch

from . import ops
from .lib import load
from .tokenizer import tokenize, CrossGeo


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
        self.use_native = False        # one-call C++ executor (csrc/engine_native.cu) for the production topology
        self._native = None
        self._tc_weights = None
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

    # ------------------------------------------------------------------ blocks
    def _block(self, blk, x, stg, drop, conv_in=None):
        a = blk.attn
        c = ops.subm_conv(x if conv_in is None else conv_in, blk.cpe[0].weight, blk.cpe[0].bias, stg.nbr3, stg.stride_alloc, rule=stg)
        c = ops.linear(c, blk.cpe[1].weight, blk.cpe[1].bias)
        x = ops.layer_norm(c, blk.cpe[2].weight, blk.cpe[2].bias, blk.cpe[2].eps, residual=x)
        h = ops.layer_norm(x, blk.norm1[0].weight, blk.norm1[0].bias, blk.norm1[0].eps)
        qkv = ops.linear(h, a.qkv.weight, a.qkv.bias)
        plan = stg.plans[a.order_index]
        o = ops.patch_attention(qkv, a.q_norm, a.k_norm, a.rpe, plan, a.num_heads, a.qk_norm, drop.site(a.attn_drop),
                                grid=stg.grid)
        x = ops.linear(o, a.proj.weight, a.proj.bias, residual=x, drop=drop.site(a.proj_drop_p))
        h = ops.layer_norm(x, blk.norm2[0].weight, blk.norm2[0].bias, blk.norm2[0].eps)
        mlp = blk.mlp[0]
        u = ops.linear(h, mlp.fc1.weight, mlp.fc1.bias)
        g = ops.act(u, "gelu", drop.site(blk.proj_drop))
        x = ops.linear(g, mlp.fc2.weight, mlp.fc2.bias, residual=x, drop=drop.site(blk.proj_drop))
        return x

    def _ca_block(self, blk, x, ctx, geo, drop):
        a = blk.attn
        h = ops.layer_norm(x, blk.norm1[0].weight, blk.norm1[0].bias, blk.norm1[0].eps)
        q = ops.linear(h, a.q.weight, a.q.bias)
        kv = ops.linear(ctx, a.kv.weight, a.kv.bias)
        o = ops.cross_attention(q, kv, a.q_norm, a.k_norm, geo, a.num_heads, a.qk_norm, drop.site(a.attn_drop))
        x = ops.linear(o, a.proj.weight, a.proj.bias, residual=x, drop=drop.site(a.proj_drop_p))
        h = ops.layer_norm(x, blk.norm2[0].weight, blk.norm2[0].bias, blk.norm2[0].eps)
        mlp = blk.mlp[0]
        u = ops.linear(h, mlp.fc1.weight, mlp.fc1.bias)
        g = ops.act(u, "gelu", drop.site(blk.proj_drop))
        x = ops.linear(g, mlp.fc2.weight, mlp.fc2.bias, residual=x, drop=drop.site(blk.proj_drop))
        return x

    # ------------------------------------------------------------------ backbone
    def backbone(self, feat, coord, npoints, ctx, tok_off_host, train, drop):
        """PointTransformerV3CA.forward(return_dec_layers=True): returns (last decoder feat, stages)."""
        pt = self.m.ptv3_model
        if not pt.with_ca or pt.pdnorm_bn or pt.pdnorm_ln:
            raise NotImplementedError("engine: only the CA backbone with plain BN/LN (the production model) is wired; "
                                      "AdaNorm/Concat variants construct and checkpoint but do not run yet")
        S = pt.num_stages
        R = len(pt.order)
        rows = [sorted({i % R for i in range(pt.enc_depths[s])} |
                       ({i % R for i in range(pt.dec_depths[s])} if s < S - 1 else set())) for s in range(S)]
        psz = [pt.enc_patch_size[s] for s in range(S)]
        for s in range(S - 1):
            assert pt.dec_patch_size[s] == psz[s], "enc/dec patch sizes of a stage must match (shared patch plan)"
        stages = tokenize(coord, npoints, pt.order, pt.stride, psz, rows, self._perms(S, R, pt.shuffle_orders),
                          self.m.config.action_config.voxel_size, self.div_mode)
        self.last_stages = stages
        dev = feat.device
        B = len(npoints)
        tok_off = torch.tensor(tok_off_host, dtype=torch.int32).to(dev, non_blocking=True)
        max_tok = max(b - a for a, b in zip(tok_off_host[:-1], tok_off_host[1:]))
        geos = [CrossGeo(stg.cloud_off, tok_off, B, stg.max_cloud, max_tok) for stg in stages]

        stem = pt.embedding.stem
        x = ops.subm_conv(feat, stem.conv.weight, None, stages[0].nbr5, stages[0].stride_alloc)     # Cin=7: fp32 SIMT in both modes
        x = ops.batch_norm_act(x, stem.norm, train, gelu=True, out_bf16=(self.precision == "bf16"))
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
                x = ops.batch_norm_act(pooled, down.norm[0], train, gelu=True)
                self._rec(f"enc{s}.down", x)
            for i in range(pt.enc_depths[s]):
                x = self._block(getattr(enc, f"block{i}"), x, stg, drop)
                self._rec(f"enc{s}.block{i}", x)
                x = self._ca_block(getattr(enc, f"ca_block{i}"), x, ctx, geos[s], drop)
                self._rec(f"enc{s}.ca_block{i}", x)
        for s in reversed(range(S - 1)):
            dec = getattr(pt.dec, f"dec{s}")
            stg, child = stages[s], stages[s + 1]
            up = dec.up
            a = ops.batch_norm_act(ops.linear(x, up.proj[0].weight, up.proj[0].bias), up.proj[1], train, gelu=True)
            b = ops.batch_norm_act(ops.linear(skips[s], up.proj_skip[0].weight, up.proj_skip[0].bias), up.proj_skip[1],
                                   train, gelu=True)
            x = ops.unpool_add(b, a, child.cluster, child.seg_ptr, child.member)
            self._rec(f"dec{s}.up", x)
            conv_in = b          # sparse_conv_feat is NOT refreshed by the add (model.py:824 vs :235-238)
            for i in range(pt.dec_depths[s]):
                x = self._block(getattr(dec, f"block{i}"), x, stg, drop, conv_in=conv_in)
                conv_in = None
                self._rec(f"dec{s}.block{i}", x)
                x = self._ca_block(getattr(dec, f"ca_block{i}"), x, ctx, geos[s], drop)
                self._rec(f"dec{s}.ca_block{i}", x)
        return x, stages

    # ------------------------------------------------------------------ native one-call step
    def native_step(self, batch, compute_loss=True, with_grad=True, compute_final_action=False, dropout_seed=None):
        """forward (+ loss + backward into the gradient sinks) through `lotus_policy_step`.  Returns (final, losses(4,))."""
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
        rows = [[0]] * S
        stages = tokenize(fts[:, :3], npoints, pt.order, pt.stride, list(pt.enc_patch_size), rows, self._perms(S, R, pt.shuffle_orders),
                          m.config.action_config.voxel_size, self.div_mode)
        self.last_stages = stages
        tok_off = [0]
        for l in batch["txt_lens"]:
            tok_off.append(tok_off[-1] + int(l))
        tok_dev = torch.tensor(tok_off, dtype=torch.int32).to(fts.device, non_blocking=True)
        self.seed_counter += 1
        seed = dropout_seed if dropout_seed is not None else int(torch.initial_seed()) * 1000003 + self.seed_counter
        return self._native.run(batch, stages, tok_dev, tok_off, m.training, self.precision == "bf16", compute_loss,
                                not compute_final_action, with_grad and compute_loss, seed)

    # ------------------------------------------------------------------ policy
    def forward(self, batch, compute_loss=False, **kwargs):
        m = self.m
        cfg = m.config
        ac = cfg.action_config
        if m.variant != "ca":
            raise NotImplementedError("engine: SimplePolicyPTV3CA is the wired variant")
        if ac.use_ee_pose or ac.use_step_id:
            raise NotImplementedError("engine: pose / step-id context tokens are disabled in production and not wired yet")
        if ac.pos_pred_type != "heatmap_disc" or ac.rot_pred_type != "euler_disc" or ac.reduce != "max":
            raise NotImplementedError("engine: the production head (heatmap_disc / euler_disc / max) is the wired one")
        fts = batch["pc_fts"]
        if not fts.is_cuda:
            raise RuntimeError("lotus engine needs the model on a CUDA device (there is no CPU fallback)")
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
            self._stage_weights()
        # context tokens (simple_policy_ptv3.py:413-429)
        txt = batch["txt_embeds"].float()
        ctx = ops.linear(ops.to_bf16(txt) if bf else txt, m.txt_fc.weight, m.txt_fc.bias)
        tok_off = [0]
        for l in batch["txt_lens"]:
            tok_off.append(tok_off[-1] + int(l))

        x, stages = self.backbone(fts, fts[:, :3], npoints, ctx, tok_off, train, drop)
        self._rec("dec_feat", x)
        s0 = stages[0]

        # action head (simple_policy_ptv3.py:113-120,149-152)
        head = m.act_proj_head
        hm, am = head.heatmap_mlp, head.action_mlp
        # (bf16 mode: the first layers run on tensor cores and emit fp32; the 90/217-wide logits stay fp32 SIMT)
        h = ops.act(ops.linear(x, hm[0].weight, hm[0].bias, out_fp32=True), "leaky", drop.site(head.dropout))
        heat = ops.linear(h, hm[3].weight, hm[3].bias)                      # (N, 3*2*pos_bins): [n, c*NB + b]
        pc = ops.segment_max(x, s0.cloud_off, None, B)                     # per-cloud max over points
        h2 = ops.act(ops.linear(pc, am[0].weight, am[0].bias, out_fp32=True), "leaky", drop.site(head.dropout))
        action = ops.linear(h2, am[3].weight, am[3].bias)                  # (B, 72*3 + 1)
        self._rec("heat", heat)
        self._rec("action", action)

        NB = 2 * head.pos_bins
        gt = batch.get("gt_actions")
        final_action = kwargs.get("compute_final_action", True)
        with torch.no_grad():
            final = ops.decode_actions(heat, fts, s0.cloud_off, B, NB, ac.pos_bin_size, action, head.euler_bins,
                                       head.euler_resolution, None if final_action else gt.float())
        if not compute_loss:
            return final
        if "disc_pos_probs_packed" in batch:
            prob = batch["disc_pos_probs_packed"]
        else:
            prob = torch.cat([p.reshape(-1) for p in batch["disc_pos_probs"]])
        prob = prob.to(fts.device, non_blocking=True).float()
        lc = cfg.loss_config
        lv = ops.policy_loss(heat, action, prob, s0.cloud_off, gt.float().contiguous(), B, NB, head.euler_bins,
                             lc.pos_weight, lc.rot_weight)
        return final, {"pos": lv[0], "rot": lv[1], "open": lv[2], "total": lv[3]}
