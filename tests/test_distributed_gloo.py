"""World-size-2 checks of the data-parallel step logic on CPU tensors (gloo).

The path has ONE collective: a sum all-reduce of the flat fp32 gradient buffer, with the 1/world mean and
the clip coefficient folded into the optimizer.  These tests drive the trainer's own flat layout + all-reduce over gloo
and check the result against the oracle's restatement of the reference optimizer (train/optim/adamw.py:53-112) applied
to the mean gradient of the two ranks — i.e. what DDP + clip_grad_norm_ + AdamW do in train_simple_policy.py:202-243.
The CUDA optimizer kernel itself is compared with the same oracle in tests/test_gpu_parity.py.
"""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_grads(rank, n):
    g = torch.Generator().manual_seed(1000 + rank)
    return torch.randn(n, generator=g) * (0.5 + rank)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import importlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfgm = importlib.import_module("robot-3dlotus_b200.config")
        policy = importlib.import_module("robot-3dlotus_b200.policy")
        trainer = importlib.import_module("robot-3dlotus_b200.trainer")
        testing = importlib.import_module("robot-3dlotus_b200.testing")
        from oracle import lotus_oracle as O

        model = policy.build_parameter_shell(cfgm.small_model_config())
        testing.deterministic_reinit(model, 77)                 # identical replicas, as after DDP's initial broadcast
        before = {n: p.detach().clone() for n, p in model.named_parameters()}
        opt = trainer.FlatAdamW(model, lr=3e-4, weight_decay=0.05, max_grad_norm=1.0)
        assert opt.world == world
        # layout: parameters alias the flat buffer, p.grad aliases the flat gradient, slots are 128-byte aligned
        for n, p in model.named_parameters():
            assert torch.equal(p.detach(), before[n])
            assert p.grad is not None and p.grad.shape == p.shape
            assert p.data_ptr() % 128 == opt.flat_param.data_ptr() % 128
        nd = opt.group_sizes[0]
        assert all("bias" not in n for n in opt.names[:nd]) and all("bias" in n for n in opt.names[nd:])

        # per-rank gradients written through the p.grad views (what the backward kernels do through GRAD_SINK)
        opt.zero_grad()
        off = 0
        for p in opt.params:
            p.grad.copy_(_rank_grads(rank, 1 << 20)[off % 4096: off % 4096 + p.numel()].view_as(p))
            off += p.numel()
        mine = [p.grad.clone() for p in opt.params]
        opt.all_reduce()                                          # the one collective of the path
        # expected: sum over ranks, recomputed locally from the seeds
        off, tot = 0, []
        for p in opt.params:
            s = sum(_rank_grads(r, 1 << 20)[off % 4096: off % 4096 + p.numel()] for r in range(world)).view_as(p)
            tot.append(s)
            off += p.numel()
        for p, s in zip(opt.params, tot):
            assert torch.allclose(p.grad, s, rtol=0, atol=1e-6)
        # slot padding stays zero (the grad-norm kernel sums the whole flat buffer)
        used = torch.zeros_like(opt.flat_grad, dtype=torch.bool)
        for p in opt.params:
            o = (p.grad.data_ptr() - opt.flat_grad.data_ptr()) // 4
            used[o: o + p.numel()] = True
        assert float(opt.flat_grad[~used].abs().sum()) == 0.0

        # optimizer on the mean gradient == oracle restatement of the reference step on the mean gradient
        mean = [s / world for s in tot]
        params = [p.detach().clone() for p in opt.params]
        m = [torch.zeros_like(p) for p in params]
        v = [torch.zeros_like(p) for p in params]
        wds = [0.05] * nd + [0.0] * (len(params) - nd)
        norm = O.adamw_step(params, mean, m, v, 1, 3e-4, wds, betas=(0.9, 0.98), eps=1e-6, max_grad_norm=1.0)
        # the fused kernel's arithmetic, restated on the flat buffers: coef = scale * min(1, max_norm/(|g_sum|*scale+1e-6))
        scale = 1.0 / world
        gn = float(torch.sqrt((opt.flat_grad.double() ** 2).sum())) * scale
        assert abs(gn - norm) <= 1e-5 * norm
        coef = scale * min(1.0, 1.0 / (gn + 1e-6))
        for name, p_ref, p, g in zip(opt.names, params, opt.params, tot):
            gi = g * coef
            mi = 0.1 * gi
            vi = 0.02 * gi * gi
            ss = 3e-4 * (1 - 0.98) ** 0.5 / (1 - 0.9)
            out = p.detach() - ss * mi / (vi.sqrt() + 1e-6)
            if "bias" not in name:
                out = out - 3e-4 * 0.05 * out
            assert torch.allclose(out, p_ref, rtol=1e-5, atol=1e-7)
        # every rank holds the same reduced gradient -> replicas stay in lock-step
        digest = torch.tensor([float(opt.flat_grad.double().sum()), float(opt.flat_grad.double().abs().sum())], dtype=torch.float64)
        both = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(both, digest)
        assert torch.equal(both[0], both[1])
        # step() itself is CUDA-only and says so
        try:
            opt.step()
            raise AssertionError("step() on CPU must refuse to run")
        except RuntimeError as e:
            assert "CUDA" in str(e)
        # the rank's own contribution differed (the test would be vacuous otherwise)
        assert not torch.allclose(mine[0], tot[0])
        ret[rank] = "ok"
    except BaseException as e:                                   # surface the failure in the parent
        ret[rank] = f"{type(e).__name__}: {e}"
        raise
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def _shard_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import importlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        syn = importlib.import_module("robot-3dlotus_b200.synthetic")
        # weak scaling: every rank draws its own keysteps (bench.py seeds by rank); shards must differ and the
        # whole-job count is world x per-rank
        b = syn.make_batch(seed=100 + rank, sizes=[300, 420, 350, 500])
        n = torch.tensor([len(b["npoints_in_batch"])], dtype=torch.int64)
        dist.all_reduce(n)
        assert int(n) == 4 * world
        digest = torch.tensor([float(b["pc_fts"].double().sum())], dtype=torch.float64)
        both = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(both, digest)
        assert not torch.equal(both[0], both[1])
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == 10.0 + world - 1
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_keystep_sharding_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_shard_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def _bcast_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import importlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfgm = importlib.import_module("robot-3dlotus_b200.config")
        policy = importlib.import_module("robot-3dlotus_b200.policy")
        trainer = importlib.import_module("robot-3dlotus_b200.trainer")
        # per-rank seeds BEFORE the model is built, exactly like train_loop.train / train_simple_policy.py:64-67: the replicas
        # initialise differently and only the set-up broadcast (DDP's wrap in the reference) makes them equal
        torch.manual_seed(2024 + rank)
        model = policy.build_parameter_shell(cfgm.small_model_config())
        # make the BatchNorm buffers rank-dependent too (they are broadcast as well)
        for b in model.buffers():
            if b.dtype.is_floating_point:
                b.add_(float(rank))
        d0 = torch.tensor([float(sum(p.double().abs().sum() for p in model.parameters()))], dtype=torch.float64)
        both = [torch.zeros_like(d0) for _ in range(world)]
        dist.all_gather(both, d0)
        assert not torch.equal(both[0], both[1]), "test is vacuous: replicas already equal before the broadcast"
        step = trainer.TrainStep(model)                          # no deterministic_reinit, no manual broadcast
        digest = torch.tensor([float(step.opt.flat_param.double().sum()), float(step.opt.flat_param.double().abs().sum()),
                               float(sum(b.double().sum() for b in model.buffers()))], dtype=torch.float64)
        both = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(both, digest)
        assert torch.equal(both[0], both[1]), (both[0], both[1])
        # parameters still alias the flat buffer after the broadcast
        for p in model.parameters():
            o = (p.data_ptr() - step.opt.flat_param.data_ptr()) // 4
            assert torch.equal(p.detach().reshape(-1), step.opt.flat_param[o: o + p.numel()])
        ret[rank] = "ok"
    except BaseException as e:
        ret[rank] = f"{type(e).__name__}: {e}"
        raise
    finally:
        dist.destroy_process_group()


def test_replicas_start_equal_world2():
    """without DDP nothing broadcast rank 0's initial weights; TrainStep now does it at construction."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bcast_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_grad_sinks_are_scoped_per_optimizer():
    """a second optimizer in the process must not drop the first model's gradient sinks."""
    sys.path.insert(0, ROOT)
    import importlib
    cfgm = importlib.import_module("robot-3dlotus_b200.config")
    policy = importlib.import_module("robot-3dlotus_b200.policy")
    trainer = importlib.import_module("robot-3dlotus_b200.trainer")
    ops = importlib.import_module("robot-3dlotus_b200.ops")
    m1 = policy.build_parameter_shell(cfgm.small_model_config())
    o1 = trainer.FlatAdamW(m1)
    keys1 = [p.data_ptr() for p in m1.parameters()]
    m2 = policy.build_parameter_shell(cfgm.small_model_config())
    o2 = trainer.FlatAdamW(m2)
    assert all(k in ops.GRAD_SINK for k in keys1)
    for p in m1.parameters():
        assert ops.GRAD_SINK[p.data_ptr()].data_ptr() == p.grad.data_ptr()
    keys2 = [p.data_ptr() for p in m2.parameters()]
    del o2, m2
    import gc
    gc.collect()
    assert not any(k in ops.GRAD_SINK for k in keys2) and all(k in ops.GRAD_SINK for k in keys1)
    assert o1 is not None


def _slice_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import importlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfgm = importlib.import_module("robot-3dlotus_b200.config")
        policy = importlib.import_module("robot-3dlotus_b200.policy")
        trainer = importlib.import_module("robot-3dlotus_b200.trainer")
        testing = importlib.import_module("robot-3dlotus_b200.testing")
        model = policy.build_parameter_shell(cfgm.small_model_config())
        testing.deterministic_reinit(model, 5)
        opt = trainer.FlatAdamW(model)
        # slices: contiguous, in backward-completion order, covering the whole flat buffer exactly once
        b = opt.slice_bounds
        assert len(b) == trainer.N_GRAD_SLICES + 1 and b[0] == 0 and b[-1] == opt.flat_grad.numel() and all(x <= y for x, y in zip(b[:-1], b[1:]))
        name_of = dict(zip([p.data_ptr() for p in opt.params], opt.names))
        for i, (n, p) in enumerate(zip(opt.names, opt.params)):
            k = max(j for j in range(trainer.N_GRAD_SLICES) if b[j] <= opt.offsets[i])
            want = 0 if (n.startswith("act_proj_head") or ".dec." in n) else 1 if ".enc.enc4." in n else 2 if ".enc.enc3." in n else 3
            assert k == want, (n, k, want)
            assert p.grad.data_ptr() == opt.flat_grad.data_ptr() + 4 * opt.offsets[i]
        g = torch.Generator().manual_seed(10 + rank)
        opt.flat_grad.copy_(torch.randn(opt.flat_grad.numel(), generator=g))
        mine = opt.flat_grad.clone()
        ref = mine.clone()
        dist.all_reduce(ref)                                        # the single-collective form
        opt.all_reduce_slices(None)                                 # the sliced form (no events on CPU)
        assert torch.equal(opt.flat_grad, ref) and not torch.equal(mine, ref)
        # optimizer state dict keeps the reference's per-parameter order whatever the flat layout
        opt.exp_avg.copy_(torch.arange(opt.exp_avg.numel(), dtype=torch.float32))
        sd = opt.state_dict()
        for i, p in enumerate(opt.params):
            assert torch.equal(sd["state"][i]["exp_avg"].reshape(-1), opt.exp_avg[opt.offsets[i]: opt.offsets[i] + p.numel()])
        opt.exp_avg.zero_()
        opt.load_state_dict(sd)
        assert float(opt.exp_avg[opt.offsets[3]]) == float(opt.offsets[3])
        ret[rank] = "ok"
    except BaseException as e:
        ret[rank] = f"{type(e).__name__}: {e}"
        raise
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_in_completion_order_slices_world2():
    """SURVEY §8e / VERDICT r1 #5: still ONE logical collective per step, issued as four contiguous slices in the order the backward
    completes them (head + decoder, encoder stage 4, stage 3, rest) so that it can overlap the backward on the GPU."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_slice_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_nccl_channel_cap_default(monkeypatch):
    """the launchers cap NCCL's channel count before creating the communicator (the all-reduce shares the SMs with persistent compute
    kernels); a value already in the environment wins"""
    import importlib
    trainer = importlib.import_module("robot-3dlotus_b200.trainer")
    monkeypatch.delenv("NCCL_MAX_NCHANNELS", raising=False)
    trainer.nccl_env_defaults()
    assert os.environ["NCCL_MAX_NCHANNELS"] == "12"
    monkeypatch.setenv("NCCL_MAX_NCHANNELS", "4")
    trainer.nccl_env_defaults()
    assert os.environ["NCCL_MAX_NCHANNELS"] == "4"
