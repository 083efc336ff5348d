"""B200-native training step for the drop-in policy.

What the reference does per step (train/train_simple_policy.py:202-243): forward, `backward()` with DDP
bucketed all-reduces, `clip_grad_norm_`, a per-tensor Python AdamW loop over 421 tensors, `zero_grad`.
Here: every parameter lives in ONE flat fp32 buffer and every gradient in another; backward kernels accumulate
straight into the flat gradient buffer (ops.GRAD_SINK); data parallelism is ONE NCCL all-reduce over that buffer
(the only collective — keysteps shard by batch); clip coefficient and the reference's AdamW
arithmetic (train/optim/adamw.py:53-112) are fused into one kernel over the flat buffers.

BatchNorm statistics stay per-rank: identical to the reference at world_size 1, and it keeps the
gradient all-reduce the only collective.
"""
import math
import weakref

import torch
import torch.distributed as dist

from . import ops
from .lib import load
from .ops import _p, _st


def warmup_cosine(step, warmup_step, tot_step, num_cycles=0.5):
    """train/optim/sched.py:44-55"""
    if step < warmup_step:
        return step / warmup_step
    progress = float(step - warmup_step) / float(max(1, tot_step - warmup_step))
    return 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress))


def get_lr_sched_decay_rate(global_step, warmup_steps, num_train_steps):
    """train/optim/sched.py:95-113 (cosine)"""
    return max(warmup_cosine(global_step, warmup_steps, num_train_steps), 1e-5)


def _drop_sinks(keys, grad_base):
    for k in keys:
        g = ops.GRAD_SINK.get(k)
        if g is not None and g.untyped_storage().data_ptr() == grad_base:
            del ops.GRAD_SINK[k]


def _completion_group(name, deepest):
    """which slice of the backward completes this parameter's gradient (csrc/engine_native.cu records one event per slice):
    0 head + decoder, 1 the deepest encoder stage (44 % of the bytes in production), 2 the stage above it, 3 everything that
    finishes with the step (shallow encoder stages, stem, text projection)"""
    if name.startswith("act_proj_head") or ".dec." in name:
        return 0
    if f".enc.enc{deepest}." in name:
        return 1
    if deepest >= 1 and f".enc.enc{deepest - 1}." in name:
        return 2
    return 3


N_GRAD_SLICES = 4



def nccl_env_defaults():
    """Call BEFORE torch.distributed.init_process_group("nccl").  The gradient all-reduce runs on a communication stream underneath the
    backward; the compute kernels are persistent (one CTA per SM holding ~200 KB of shared memory) and chained by programmatic dependent
    launch, so every SM an NCCL channel occupies is an SM a GEMM tile waits for.  With NCCL's default channel count the 2-GPU step measured
    18.6 ms; capped at 12 channels 15.2 ms (2: 19.1, 4: 16.2, 8: 15.2, 12: 15.2; one GPU: 14.3).  A value already in the environment wins."""
    import os
    os.environ.setdefault("NCCL_MAX_NCHANNELS", "12")

class FlatAdamW:
    """The reference optimizer (optim/misc.py:13-55 grouping: names containing 'bias' are not decayed; optim/adamw.py)
    over flat buffers.  `state_dict()` / `load_state_dict()` use the reference's per-parameter layout
    (`{'state': {i: {'step','exp_avg','exp_avg_sq'}}, 'param_groups': [...]}`) so `train_state_latest.pt` round-trips."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.05, max_grad_norm=10.0):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        assert named and all(p.dtype == torch.float32 for _, p in named), "fp32 master parameters expected"
        # the flat layout and the all-reduce are device-agnostic (gloo tests drive them on CPU tensors);
        # `step()` launches the fused CUDA kernels and refuses to run anywhere else
        dev = named[0][1].device
        # group order of the reference: decayed tensors first, then the 'bias' tensors
        decay = [(n, p) for n, p in named if "bias" not in n]
        no_decay = [(n, p) for n, p in named if "bias" in n]
        self.names = [n for n, _ in decay + no_decay]
        self.params = [p for _, p in decay + no_decay]
        self.group_sizes = (len(decay), len(no_decay))
        sizes = [p.numel() for p in self.params]
        slots = [(n + 31) // 32 * 32 for n in sizes]        # 128-byte aligned slots (TMA reduce-add targets need 16-byte bases)
        total = sum(slots)
        self.flat_param = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        # The FLAT layout follows the order in which the backward completes gradients (head + decoder first, then encoder stage 4, 3,
        # and the shallow stages / stem / text projection last), so the one logical gradient all-reduce of the path can be issued as a
        # few contiguous slices behind events recorded inside the backward and overlap it (`all_reduce_slices`).  `self.params` keeps
        # the reference's group order (decayed tensors, then biases) for the optimizer state dict; `self.offsets[i]` locates params[i].
        import re
        stage_ids = [int(m.group(1)) for n in self.names for m in [re.search(r"\.enc\.enc(\d+)\.", n)] if m]
        deepest = max(stage_ids) if stage_ids else 0
        grp = [_completion_group(n, deepest) for n in self.names]
        layout = sorted(range(len(self.params)), key=lambda i: (grp[i], i))
        self.offsets = [0] * len(self.params)
        ends, off = [], 0
        starts = [None] * N_GRAD_SLICES
        # gradient sinks are process-global (keyed by parameter address): drop only what THIS model registered before
        # (re-building an optimizer for the same model), never another live model's entries
        for p in self.params:
            ops.GRAD_SINK.pop(p.data_ptr(), None)
        self._sink_keys = []
        for pos, i in enumerate(layout):
            p, n, slot = self.params[i], sizes[i], slots[i]
            if starts[grp[i]] is None:
                starts[grp[i]] = off
            self.offsets[i] = off
            view = self.flat_param[off: off + n].view_as(p)
            view.copy_(p.data)
            p.data = view                                   # parameters now alias the flat buffer
            g = self.flat_grad[off: off + n].view_as(p)
            p.grad = g                                      # reference-style access (`p.grad`) keeps working
            ops.GRAD_SINK[p.data_ptr()] = g
            self._sink_keys.append(p.data_ptr())
            off += slot
            ends.append(off)
        self.layout = layout
        nxt = total
        for g in reversed(range(N_GRAD_SLICES)):            # an empty group starts where the next one does
            if starts[g] is None:
                starts[g] = nxt
            nxt = starts[g]
        self.slice_bounds = starts + [total]                # slice k = flat[slice_bounds[k] : slice_bounds[k + 1]], k in completion order
        # the entries die with this optimizer (its flat buffer is freed then; a later allocation may reuse the addresses)
        weakref.finalize(self, _drop_sinks, list(self._sink_keys), self.flat_grad.data_ptr())
        self.seg_end = torch.tensor(ends, dtype=torch.int64, device=dev)
        wd_of = [weight_decay] * len(decay) + [0.0] * len(no_decay)
        self.seg_wd = torch.tensor([wd_of[i] for i in layout], dtype=torch.float32, device=dev)     # segments are in LAYOUT order
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.step_count = 0
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def zero_grad(self):
        self.flat_grad.zero_()

    def broadcast_parameters(self, model, src=0):
        """What DDP's wrap does once (reference train/utils/distributed.py:196-212 `wrap_model` -> DistributedDataParallel broadcasts
        rank 0's parameters and buffers at construction): the per-rank seeds (train_simple_policy.py:64-67) make every rank
        initialise different weights; after this call all replicas start from rank `src`'s.  One broadcast of the flat parameter
        buffer plus the BatchNorm buffers; set-up only, never inside a step."""
        if self.world <= 1:
            return
        dist.broadcast(self.flat_param, src=src)
        for b in model.buffers():
            dist.broadcast(b, src=src)
        ops.PARAM_EPOCH[0] += 1

    def all_reduce(self):
        """the one collective of the path: sum of the flat gradient over ranks (mean applied inside the AdamW kernel)"""
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)

    def slices(self):
        return [self.flat_grad[a:b] for a, b in zip(self.slice_bounds[:-1], self.slice_bounds[1:])]

    def all_reduce_slices(self, events=None):
        """The same collective, issued as contiguous slices in backward-completion order.  `events[k]` (torch.cuda.Event recorded by the
        native backward when slice k's gradients are complete; the last slice needs none) lets slice k start on a communication stream
        while the backward is still producing slices k+1..: only the last, small slice stays exposed.  Without events (CPU / gloo
        tests, per-op engine) the slices are reduced back to back -- same result as `all_reduce`."""
        if self.world <= 1:
            return
        if events is None or not self.flat_grad.is_cuda:
            for sl in self.slices():
                if sl.numel():
                    dist.all_reduce(sl, op=dist.ReduceOp.SUM)
            return
        if getattr(self, "_comm_stream", None) is None:
            self._comm_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        works = []
        with torch.cuda.stream(self._comm_stream):
            for k, sl in enumerate(self.slices()):
                if k < len(events) and events[k] is not None:
                    self._comm_stream.wait_event(events[k])
                else:
                    self._comm_stream.wait_stream(main)        # the tail slice: everything enqueued so far
                if sl.numel():
                    works.append(dist.all_reduce(sl, op=dist.ReduceOp.SUM, async_op=True))
        for w in works:
            w.wait()                                           # the main stream waits for the collectives
        main.wait_stream(self._comm_stream)

    def step(self, lr=None, grad_scale=1.0):
        if not self.flat_param.is_cuda:
            raise RuntimeError("FlatAdamW.step runs the fused CUDA optimizer kernels: move the model to CUDA first "
                               "(there is no CPU optimizer path)")
        L = load()
        lr = self.lr if lr is None else lr
        self.step_count += 1
        ops.PARAM_EPOCH[0] += 1
        st = _st()
        n = self.flat_param.numel()
        scale = grad_scale / self.world
        self.last_scale = scale
        if self.max_grad_norm and self.max_grad_norm > 0:
            L.lotus_grad_sumsq_f32(_p(self.flat_grad), n, _p(self.sumsq), st)
            sumsq = _p(self.sumsq)
        else:
            sumsq = None
        L.lotus_adamw_step_f32(_p(self.flat_param), _p(self.flat_grad), _p(self.exp_avg), _p(self.exp_avg_sq), n,
                               _p(self.seg_end), _p(self.seg_wd), len(self.params), lr, self.betas[0], self.betas[1],
                               self.eps, self.step_count, 1 | 2, sumsq, float(self.max_grad_norm or 0.0), scale, st)   # 1: bias correction, 2: slots are 4-aligned

    def grad_norm(self):
        """global gradient norm of the last step (after the all-reduce mean) — one D2H read"""
        return math.sqrt(float(self.sumsq.item())) * getattr(self, "last_scale", 1.0 / self.world)

    # ---- reference checkpoint layout (train/utils/save.py:38-45)
    def state_dict(self):
        state = {}
        for i, p in enumerate(self.params):
            n, off = p.numel(), self.offsets[i]
            state[i] = {"step": self.step_count, "exp_avg": self.exp_avg[off: off + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off: off + n].view_as(p).clone()}
        nd = self.group_sizes[0]
        groups = [
            {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay, "correct_bias": True,
             "params": list(range(nd))},
            {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0.0, "correct_bias": True,
             "params": list(range(nd, len(self.params)))},
        ]
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        for i, p in enumerate(self.params):
            n, off = p.numel(), self.offsets[i]
            if i in sd["state"]:
                self.exp_avg[off: off + n].copy_(sd["state"][i]["exp_avg"].reshape(-1))
                self.exp_avg_sq[off: off + n].copy_(sd["state"][i]["exp_avg_sq"].reshape(-1))
                self.step_count = int(sd["state"][i]["step"])


class TrainStep:
    """one optimizer step of `train_simple_policy.py:202-243` on the calling rank"""

    def __init__(self, model, learning_rate=1e-4, betas=(0.9, 0.98), weight_decay=0.05, grad_norm=10.0, warmup_steps=5000,
                 num_train_steps=150000):
        self.model = model
        self.opt = FlatAdamW(model, learning_rate, betas, 1e-6, weight_decay, grad_norm)
        self.learning_rate, self.warmup_steps, self.num_train_steps = learning_rate, warmup_steps, num_train_steps
        self.global_step = 0
        self.lr_fn = None             # optional schedule override: global_step -> lr (train_loop.py)
        self.micro_step = 0           # micro-batches accumulated since the last optimizer step (carried across epochs)
        import os
        # world > 1, native engine: reduce gradient slices behind the backward (LOTUS_NO_AR_OVERLAP=1: one slice after the other at the end)
        self.overlap_allreduce = os.environ.get("LOTUS_NO_AR_OVERLAP") is None
        self._events, self._slice_events = None, None
        self.opt.broadcast_parameters(model)      # world > 1: replicas start from rank 0's weights, as under DDP

    def __call__(self, batch, next_batch=None):
        losses = self.accumulate(batch, first=True, next_batch=next_batch)
        self.apply()
        return losses

    def accumulate(self, batch, first=True, next_batch=None):
        """forward + backward of one (micro-)batch; gradients ACCUMULATE in the flat buffer (zeroed when `first`).
        `next_batch` (native engine): the batch of the NEXT call -- its host->device copy and integer tokeniser are launched on a
        side stream now, so they overlap this step instead of opening the next one with an idle GPU."""
        if first is None:                                # running counter: zero only after an optimizer step (reference: zero_grad
            first = self.micro_step == 0                 # follows optimizer.step(), so a partial window carries over an epoch boundary)
        if first:
            self.opt.zero_grad()
            self.micro_step = 0
        self.micro_step += 1
        eng = self.model.engine()
        if eng.use_native:
            pre = eng.take_prefetched(batch)
            pending = None
            if pre is not None and pre[2]:
                batch, pending = pre[0], pre[1]
            else:
                batch = self.model.prepare_batch(batch)
            self._slice_events = None
            if self.opt.world > 1 and self.overlap_allreduce:
                if self._events is None:
                    self._events = [torch.cuda.Event() for _ in range(N_GRAD_SLICES - 1)]
                self._slice_events = self._events
            _, lv = eng.native_step(batch, compute_loss=True, with_grad=True, compute_final_action=False, pending=pending,
                                    slice_events=self._slice_events)
            if next_batch is not None:                   # after this step is enqueued: the GPU is busy while the host prepares the next
                eng.prefetch(next_batch, need_pairs=True)
            losses = {"pos": lv[0], "rot": lv[1], "open": lv[2], "total": lv[3]}
        else:
            _, losses = self.model(batch, compute_loss=True, compute_final_action=False)
            losses["total"].backward()
        return losses

    def apply(self, accumulation=1):
        """all-reduce, clip, AdamW with the scheduled learning rate (train_simple_policy.py:214-232); the 1/accumulation of the
        reference's loss scaling (:206-207) is folded into the optimizer kernel's gradient scale"""
        # (with gradient accumulation the early slices are only final after the LAST micro-batch: events of that one are used)
        self.opt.all_reduce_slices(self._slice_events)
        self._slice_events = None
        self.micro_step = 0
        self.global_step += 1
        if self.lr_fn is not None:
            lr = self.lr_fn(self.global_step)
        else:
            lr = self.learning_rate * get_lr_sched_decay_rate(self.global_step, self.warmup_steps, self.num_train_steps)
        self.opt.step(lr, grad_scale=1.0 / accumulation)
