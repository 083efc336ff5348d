"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the reference-minted golden
vectors.  Integer / index outputs bit-exact; fp32 activations and gradients within 1e-4 relative (north star).
Run on the B200 box:  python -m pytest tests -m gpu
"""
import ctypes

import numpy as np
import pytest
import torch

from conftest import pkg
from oracle import lotus_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def make_model(g):
    cfg = pkg("config").ConfigNode(g["config"])
    model = pkg("policy").build_parameter_shell(cfg)
    pkg("testing").deterministic_reinit(model, g["reinit_seed"])
    return cfg, model



def _bf16_norm_tol(name, grad):
    """relative tolerance on a gradient NORM of the bf16 tensor-core mode against the fp32 reference.  Weights: 10 % (measured spread
    over 95 repeats of the production step, tools/flaky_probe.py: <= 3.5 %).  One-dimensional parameters (biases, norm scales) are column
    sums over up to 97 k rows with heavy cancellation, and the step is not bit-reproducible (fp32 atomics feed bf16 roundings): the same
    95 repeats put `enc0.block0.cpe.1.bias` anywhere between 2 % and 8 %, so they get 25 %."""
    return 0.25 if grad.dim() <= 1 else 0.10

def lattice_clouds(sizes, side, seed):
    """unique voxels on a side^3 lattice per cloud; coordinates sit mid-voxel and an anchor point at the origin
    (first point of cloud 0) pins the batch minimum, so the model's quantisation returns the lattice exactly."""
    rng = np.random.default_rng(seed)
    grids = []
    for i, n in enumerate(sizes):
        v = rng.permutation(side ** 3 - 1)[: n - (1 if i == 0 else 0)] + 1          # voxel 0 is reserved for the anchor
        g = np.stack([v // (side * side), (v // side) % side, v % side], 1)
        if i == 0:
            g = np.concatenate([np.zeros((1, 3), dtype=g.dtype), g])
        grids.append(g)
    grid = torch.from_numpy(np.concatenate(grids)).int()
    coord = grid.float() * 0.01 + 0.005
    coord[0] = 0.0
    assert torch.equal(O.grid_coords(coord), grid)
    return coord, grid


# ------------------------------------------------------------------------------------------------ tokeniser
def _tokenize_gold(g, perms):
    tk = pkg("tokenizer")
    b = g["batch"]
    coord = b["pc_fts"][:, :3].to(DEV)
    rows = [[0, 1, 2, 3]] * 5
    return tk.tokenize(coord, b["npoints_in_batch"], ["z", "z-trans", "hilbert", "hilbert-trans"], [2, 2, 2, 2],
                       [128] * 5, rows, [p.tolist() for p in perms])


def test_tokeniser_matches_reference_golden(gold_small):
    g = gold_small
    stages = _tokenize_gold(g, g["perms"])
    names = ["embedding"] + [f"enc{s}.down" for s in range(1, 5)]
    for s, (stg, name) in enumerate(zip(stages, names)):
        ref = g["inter"][name]
        n = ref["grid_coord"].shape[0]
        assert stg.n == n and stg.depth == ref["serialized_depth"], (s, stg.n, n)
        assert torch.equal(stg.grid.cpu(), ref["grid_coord"]), s
        assert torch.equal(stg.batch.cpu().long(), ref["batch"]), s
        assert torch.equal(stg.code[:, :n].cpu(), ref["serialized_code"]), s
        assert torch.equal(stg.order[:, :n].cpu().long(), ref["serialized_order"]), s
        assert torch.equal(stg.inverse[:, :n].cpu().long(), ref["serialized_inverse"]), s
        assert torch.equal(torch.tensor(stg.cloud_off_host[1:]), ref["offset"]), s
        if s > 0:
            assert torch.equal(stg.cluster.cpu().long(), ref["pooling_inverse"]), s
        blk = g["inter"][f"enc{s}.block0"]
        pl = stg.plans[0]
        assert torch.equal(pl.pad.cpu().long(), blk["pad"]) and torch.equal(pl.unpad.cpu().long(), blk["unpad"]), s
        assert torch.equal(pl.cu_seqlens.cpu(), blk["cu_seqlens_key"]), s
        # compositions the attention kernels consume
        assert torch.equal(pl.gather_idx.cpu().long(), ref["serialized_order"][0][blk["pad"]]), s
        assert torch.equal(pl.slot_of_point.cpu().long(), blk["unpad"][ref["serialized_inverse"][0]]), s


def test_tokeniser_invariants_full_size():
    """BASELINE config-2 sizes (32 keysteps, <=4096 points): size-independent properties + oracle equality."""
    syn, tk = pkg("synthetic"), pkg("tokenizer")
    b = syn.make_batch(seed=2024, batch_size=32, with_targets=False)
    coord = b["pc_fts"][:, :3]
    perms = [torch.randperm(4, generator=torch.Generator().manual_seed(s)).tolist() for s in range(5)]
    stages = tk.tokenize(coord.to(DEV), b["npoints_in_batch"], ["z", "z-trans", "hilbert", "hilbert-trans"], [2, 2, 2, 2],
                         [128] * 5, [[0, 1, 2, 3]] * 5, perms)
    batch_ids = O.offset2batch(b["offset"])
    grid = O.grid_coords(coord)
    ser = O.serialize(grid, batch_ids, perm=torch.tensor(perms[0]))
    pt = dict(grid=grid, batch=batch_ids, offset=b["offset"], **ser)
    for s, stg in enumerate(stages):
        n = stg.n
        assert n == pt["grid"].shape[0]
        assert torch.equal(stg.grid.cpu(), pt["grid"]) and torch.equal(stg.code[:, :n].cpu(), pt["code"])
        order, inverse = stg.order[:, :n].cpu().long(), stg.inverse[:, :n].cpu().long()
        assert torch.equal(order, pt["order"]) and torch.equal(inverse, pt["inverse"])
        ar = torch.arange(n)
        for r in range(4):
            assert torch.equal(inverse[r][order[r]], ar)                       # inverse o order = id
            c = stg.code[r, :n].cpu()[order[r]]
            assert bool((c[1:] > c[:-1]).all())                                # strictly sorted (unique voxels)
        pad, unpad, cu = O.patch_plan(pt["offset"], 128)
        pl = stg.plans[2]
        assert torch.equal(pl.pad.cpu().long(), pad) and torch.equal(pl.unpad.cpu().long(), unpad)
        assert torch.equal(pl.cu_seqlens.cpu(), cu)
        assert torch.equal(pl.pad.cpu().long()[pl.unpad.cpu().long()], ar)     # unpad o pad = id on real slots
        nbr = O.neighbor_table(pt["batch"], pt["grid"], 3)
        assert torch.equal(stg.nbr3[:, :n].cpu().long().t(), nbr)
        if s == 0:
            assert torch.equal(stg.nbr5[:, :n].cpu().long().t(), O.neighbor_table(pt["batch"], pt["grid"], 5))
        if s < 4:
            st = O.pool_structure(pt, torch.tensor(perms[s + 1]))
            nxt = stages[s + 1]
            assert torch.equal(nxt.cluster.cpu().long(), st["cluster"])
            # CSR consistency: members of cluster m are exactly the points with cluster id m
            seg, mem = nxt.seg_ptr.cpu().long(), nxt.member[:n].cpu().long()
            assert torch.equal(st["cluster"][mem], torch.repeat_interleave(torch.arange(nxt.n), seg[1:] - seg[:-1]))
            pt = dict(grid=st["grid"], batch=st["batch"], offset=st["offset"], code=st["code"], order=st["order"],
                      inverse=st["inverse"], depth=st["depth"])


def test_tokeniser_duplicate_voxels():
    """VERDICT r1 weak #3 / SURVEY D.4: the reference's augmentation (z-rotation + U(0, 2 mm) noise BEFORE re-gridding,
    simple_policy_dataset.py:158-181) puts two points into one voxel now and then.  `torch.argsort` leaves the tie order of equal codes
    unspecified and spconv keeps one of the duplicate voxels; this path pins both: ties sort by original index (== torch.argsort(stable=True),
    the oracle) and a neighbour entry is the LOWEST-index point of the neighbouring voxel.  Checked bit-exact at every stage."""
    tk = pkg("tokenizer")
    sizes = [700, 300, 129, 257]
    coord, grid = lattice_clouds(sizes, 20, 31)
    rng = np.random.default_rng(9)
    # duplicate ~12 % of the points of every cloud (same coordinates -> same voxel), some of them three times, keeping clouds contiguous
    pieces, new_sizes, s0 = [], [], 0
    for n in sizes:
        idx = np.arange(s0, s0 + n)
        dup = rng.choice(idx, size=max(2, n // 8), replace=True)
        allidx = np.concatenate([idx, dup])
        if s0 == 0:
            allidx = np.concatenate([allidx[:1], rng.permutation(allidx[1:])])      # the anchor point stays first
        else:
            allidx = rng.permutation(allidx)
        pieces.append(allidx)
        new_sizes.append(len(allidx))
        s0 += n
    sel = torch.from_numpy(np.concatenate(pieces))
    coord, grid = coord[sel], grid[sel]
    assert torch.equal(O.grid_coords(coord), grid)
    off = torch.tensor(np.cumsum(new_sizes))
    perms = [torch.randperm(4, generator=torch.Generator().manual_seed(40 + s)).tolist() for s in range(3)]
    stages = tk.tokenize(coord.to(DEV), new_sizes, ["z", "z-trans", "hilbert", "hilbert-trans"], [2, 2], [128] * 3, [[0, 1, 2, 3]] * 3, perms)
    batch_ids = O.offset2batch(off)
    ser = O.serialize(grid, batch_ids, perm=torch.tensor(perms[0]))
    pt = dict(grid=grid, batch=batch_ids, offset=off, **ser)
    assert int((torch.unique(ser["code"][0], return_counts=True)[1] > 1).sum()) > 50        # the fixture really has duplicates
    for s, stg in enumerate(stages):
        n = stg.n
        assert n == pt["grid"].shape[0]
        assert torch.equal(stg.grid.cpu(), pt["grid"]) and torch.equal(stg.code[:, :n].cpu(), pt["code"])
        order, inverse = stg.order[:, :n].cpu().long(), stg.inverse[:, :n].cpu().long()
        assert torch.equal(order, pt["order"]) and torch.equal(inverse, pt["inverse"])     # stable tie order
        pad, unpad, cu = O.patch_plan(pt["offset"], 128)
        pl = stg.plans[1]
        assert torch.equal(pl.pad.cpu().long(), pad) and torch.equal(pl.unpad.cpu().long(), unpad) and torch.equal(pl.cu_seqlens.cpu(), cu)
        # neighbour table: oracle entry mapped to the lowest-index point of that voxel
        key = (pt["batch"].long() << 48) | (pt["grid"][:, 0].long() << 32) | (pt["grid"][:, 1].long() << 16) | pt["grid"][:, 2].long()
        uniq, invk = torch.unique(key, return_inverse=True)
        canon = torch.full((uniq.numel(),), n, dtype=torch.long).scatter_reduce(0, invk, torch.arange(n), "amin")[invk]
        nbr = O.neighbor_table(pt["batch"], pt["grid"], 3)
        exp = torch.where(nbr >= 0, canon[nbr.clamp(min=0)], nbr)
        assert torch.equal(stg.nbr3[:, :n].cpu().long().t(), exp)
        if s < 2:
            st = O.pool_structure(pt, torch.tensor(perms[s + 1]))
            nxt = stages[s + 1]
            assert torch.equal(nxt.cluster.cpu().long(), st["cluster"])
            seg, mem = nxt.seg_ptr.cpu().long(), nxt.member[:n].cpu().long()
            # CSR: same segment boundaries, and the members of cluster m are exactly the points with cluster id m (the order INSIDE a
            # cluster differs -- code order here, index order in the reference -- and nothing downstream depends on it: max / mean are
            # order-free and every member of a cluster shares grid >> 1, batch and all four codes >> 3)
            assert torch.equal(seg, st["idx_ptr"])
            assert torch.equal(st["cluster"][mem], torch.repeat_interleave(torch.arange(nxt.n), seg[1:] - seg[:-1]))
            assert torch.equal(torch.sort(mem)[0], torch.arange(n))
            pt = dict(grid=st["grid"], batch=st["batch"], offset=st["offset"], code=st["code"], order=st["order"],
                      inverse=st["inverse"], depth=st["depth"])


@pytest.mark.parametrize("depth", [1, 3, 7, 9, 12, 16])
def test_codes_golden_all_depths(units, depth):
    """curve codes at every depth the reference supports (known-answer vectors from the reference)."""
    L = pkg("lib").load()
    c = units["codes"][depth]
    n = c["grid"].shape[0]
    grid = c["grid"].to(DEV).int().contiguous()
    bt = c["batch"].to(DEV).int().contiguous()
    # one pseudo-cloud per distinct batch id is not needed for the code kernel: sort a single segment
    meta = torch.tensor([n, depth, 0, 0, 0, 0, 0, 0], dtype=torch.int32, device=DEV)
    off = torch.tensor([0, n], dtype=torch.int32, device=DEV)
    code = torch.empty((4, n), dtype=torch.int64, device=DEV)
    order = torch.empty((4, n), dtype=torch.int32, device=DEV)
    inverse = torch.empty_like(order)
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    cv = (ctypes.c_int32 * 4)(0, 1, 2, 3)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    L.lotus_serialize(P(grid), P(bt), P(off), 1, n, n, P(meta), cv, 4, n, P(code), P(order), P(inverse), P(err), None)
    torch.cuda.synchronize()
    for r, name in enumerate(["z", "z-trans", "hilbert", "hilbert-trans"]):
        assert torch.equal(code[r].cpu(), c[name]), (depth, name)


def test_empty_and_tiny_clouds():
    """ragged input: clouds of 1 and 2 points next to a padded one (reference patch-plan edge cases)."""
    tk = pkg("tokenizer")
    sizes = [1, 300, 2, 129]
    coord, _ = lattice_clouds(sizes, 40, 0)
    stages = tk.tokenize(coord.to(DEV), sizes, ["z", "z-trans", "hilbert", "hilbert-trans"], [2, 2, 2, 2], [128] * 5,
                         [[0, 1, 2, 3]] * 5, None)          # all four serialization rows are checked below
    off = torch.tensor(np.cumsum(sizes))
    pad, unpad, cu = O.patch_plan(off, 128)
    pl = stages[0].plans[0]
    assert torch.equal(pl.pad.cpu().long(), pad) and torch.equal(pl.cu_seqlens.cpu(), cu)
    grid = O.grid_coords(coord)
    ser = O.serialize(grid, O.offset2batch(off))
    assert torch.equal(stages[0].order[:, : sum(sizes)].cpu().long(), ser["order"])


# ------------------------------------------------------------------------------------------------ FPS / kNN / RPE
def test_fps_bit_exact(units):
    L = pkg("lib").load()
    f = units["fps"]
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    pts = f["points"].to(DEV).contiguous()
    n = pts.shape[0]
    off = torch.tensor([0, n], dtype=torch.int32, device=DEV)
    start = torch.tensor([f["start"]], dtype=torch.int32, device=DEV)
    out = torch.empty((1, f["npoint"]), dtype=torch.int32, device=DEV)
    L.lotus_fps(P(pts), pts.shape[1], P(off), 1, n, f["npoint"], P(start), P(out), None)
    assert torch.equal(f["points"][out[0].cpu().long()], f["sampled"])          # golden from the reference helper
    # batched, larger, against the oracle restatement
    rng = np.random.default_rng(1)
    sizes = [4096, 1000, 37]
    xyz = rng.standard_normal((sum(sizes), 3)).astype(np.float32)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    starts = [5, 999, 0]
    out = torch.empty((3, 32), dtype=torch.int32, device=DEV)
    x = torch.from_numpy(xyz).to(DEV)
    t_off = torch.tensor(offs, dtype=torch.int32, device=DEV)          # keep the device buffers alive across the call
    t_start = torch.tensor(starts, dtype=torch.int32, device=DEV)
    L.lotus_fps(P(x), 3, P(t_off), 3, max(sizes), 32, P(t_start), P(out), None)
    for b in range(3):
        ref = O.farthest_point_sample_indices(xyz[offs[b]:offs[b + 1]], 32, starts[b]) + offs[b]
        assert np.array_equal(out[b].cpu().numpy(), ref), b


def test_knn_and_ball_query_bit_exact():
    L = pkg("lib").load()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    rng = np.random.default_rng(2)
    qs, rs = [300, 50], [2000, 700]
    q = (rng.integers(0, 60, size=(sum(qs), 3)) * 0.01).astype(np.float32)      # grid points -> many exact ties
    r = (rng.integers(0, 60, size=(sum(rs), 3)) * 0.01).astype(np.float32)
    qo, ro = np.concatenate([[0], np.cumsum(qs)]), np.concatenate([[0], np.cumsum(rs)])
    k = 16
    idx = torch.empty((sum(qs), k), dtype=torch.int32, device=DEV)
    dist = torch.empty((sum(qs), k), dtype=torch.float32, device=DEV)
    tq, tr = torch.from_numpy(q).to(DEV), torch.from_numpy(r).to(DEV)
    tqo, tro = torch.tensor(qo, dtype=torch.int32, device=DEV), torch.tensor(ro, dtype=torch.int32, device=DEV)
    L.lotus_knn(P(tq), P(tqo), P(tr), P(tro), 2, max(qs), k, P(idx), P(dist), None)
    ri, rd = O.knn_indices(q, r, k, qo[1:], ro[1:])
    assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(dist.cpu().numpy(), rd)
    ns = 8
    bq = torch.empty((sum(qs), ns), dtype=torch.int32, device=DEV)
    L.lotus_ball_query(P(tq), P(tqo), P(tr), P(tro), 2, max(qs), 0.05, ns, P(bq), None)
    for b in range(2):
        ref = O.ball_query_indices(q[qo[b]:qo[b + 1]], r[ro[b]:ro[b + 1]], 0.05, ns)
        ref = np.where(ref >= 0, ref + ro[b], -1)
        assert np.array_equal(bq[qo[b]:qo[b + 1]].cpu().numpy(), ref), b


def test_rpe_bias_golden(units):
    L = pkg("lib").load()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    r = units["rpe"]
    # build grid coords whose pairwise differences reproduce rel[p, i, j] = g[i] - g[j] for a 16-point patch
    rng = np.random.default_rng(3)
    K, H = 16, r["table"].shape[1]
    g = torch.from_numpy(rng.integers(0, 60, size=(2 * K, 3))).int()
    rel_ = (g.view(2, K, 1, 3) - g.view(2, 1, K, 3)).long()
    want = O.rpe_bias(r["table"], rel_, r["pos_bnd"])
    out = torch.empty((2, H, K, K), device=DEV)
    gi = torch.arange(2 * K, dtype=torch.int32, device=DEV)
    t_table, t_g = r["table"].to(DEV), g.to(DEV)
    L.lotus_rpe_bias(P(t_table), P(t_g), P(gi), 2, K, H, r["pos_bnd"], P(out), None)
    assert torch.allclose(out.cpu(), want, atol=1e-7)
    assert torch.allclose(O.rpe_bias(r["table"], r["rel"], r["pos_bnd"]), r["bias"], atol=1e-7)   # oracle == reference


# ------------------------------------------------------------------------------------------------ float kernels vs torch fp32
def _check_grads(fn_cuda, fn_ref, inputs, tol=TOL):
    """inputs: dict name -> cpu tensor (float ones require grad).  Compares outputs and all input grads."""
    cu = {k: (v.to(DEV).requires_grad_(v.dtype.is_floating_point)) for k, v in inputs.items()}
    rf = {k: (v.clone().requires_grad_(v.dtype.is_floating_point)) for k, v in inputs.items()}
    yc, yr = fn_cuda(**cu), fn_ref(**rf)
    assert rel(yc, yr) < tol, ("fwd", rel(yc, yr))
    go = torch.randn(yr.shape, generator=torch.Generator().manual_seed(5))
    yc.backward(go.to(DEV))
    yr.backward(go)
    # some gradients are mathematically zero (k_norm.bias under softmax): compare those against the scale of
    # the largest gradient of the same call instead of against their own rounding noise
    scale = max(rf[k].grad.abs().max().item() for k in inputs if rf[k].grad is not None)
    for k in inputs:
        if rf[k].grad is not None:
            gc = cu[k].grad if cu[k].grad is not None else torch.zeros_like(cu[k])     # unused input -> zero grad
            if rf[k].grad.abs().max().item() < 1e-4 * scale:
                assert gc.abs().max().item() < 1e-4 * scale, (k, "expected ~0", gc.abs().max().item(), scale)
            else:
                assert rel(gc, rf[k].grad) < tol, (k, rel(gc, rf[k].grad))


def test_linear_layernorm_act_kernels():
    ops = pkg("ops")
    g = torch.Generator().manual_seed(0)
    for (M, K, N) in [(777, 96, 130), (5, 7, 3), (4096, 64, 256)]:
        _check_grads(lambda x, w, b, r: ops.linear(x, w, b, residual=r),
                     lambda x, w, b, r: r + torch.nn.functional.linear(x, w, b),
                     dict(x=torch.randn(M, K, generator=g), w=torch.randn(N, K, generator=g) * 0.1,
                          b=torch.randn(N, generator=g), r=torch.randn(M, N, generator=g)))
    for C in (24, 32, 96, 768):
        _check_grads(lambda x, w, b, r: ops.layer_norm(x, w, b, 1e-6, residual=r),
                     lambda x, w, b, r: r + torch.nn.functional.layer_norm(x, (x.shape[1],), w, b, 1e-6),
                     dict(x=torch.randn(333, C, generator=g) * 3 + 1, w=torch.randn(C, generator=g),
                          b=torch.randn(C, generator=g), r=torch.randn(333, C, generator=g)))
    for kind, f in (("gelu", torch.nn.functional.gelu), ("leaky", lambda t: torch.nn.functional.leaky_relu(t, 0.02))):
        _check_grads(lambda x: ops.act(x, kind), lambda x: f(x), dict(x=torch.randn(1000, 33, generator=g) * 2))


def test_batchnorm_gelu_kernel():
    ops = pkg("ops")
    g = torch.Generator().manual_seed(1)
    for rows, C in [(5000, 64), (38, 96), (2, 8)]:
        x = torch.randn(rows, C, generator=g) * 2 + 0.5
        bn_c = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01)
        with torch.no_grad():
            bn_c.weight.copy_(torch.randn(C, generator=g)); bn_c.bias.copy_(torch.randn(C, generator=g))
            bn_c.running_mean.copy_(torch.randn(C, generator=g)); bn_c.running_var.copy_(torch.rand(C, generator=g) + 0.5)
        import copy
        bn_g = copy.deepcopy(bn_c).to(DEV)
        for train in (True, False):
            bn_c.train(train); bn_g.train(train)
            xc = x.clone().requires_grad_(True); xg = x.to(DEV).requires_grad_(True)
            yc = torch.nn.functional.gelu(bn_c(xc))
            yg = ops.batch_norm_act(xg, bn_g, train, gelu=True)
            assert rel(yg, yc) < TOL
            go = torch.randn(rows, C, generator=g)
            for p in list(bn_c.parameters()) + list(bn_g.parameters()):
                p.grad = None
            yc.backward(go); yg.backward(go.to(DEV))
            assert rel(xg.grad, xc.grad) < TOL and rel(bn_g.weight.grad, bn_c.weight.grad) < TOL
            assert rel(bn_g.bias.grad, bn_c.bias.grad) < TOL
            assert rel(bn_g.running_mean, bn_c.running_mean) < 1e-5 and rel(bn_g.running_var, bn_c.running_var) < 1e-5
            assert int(bn_g.num_batches_tracked) == int(bn_c.num_batches_tracked)


def test_subm_conv_kernel():
    ops, tk = pkg("ops"), pkg("tokenizer")
    sizes = [500, 260]
    coord, grid = lattice_clouds(sizes, 20, 4)
    stages = tk.tokenize(coord.to(DEV), sizes, ["z"], [2], [128, 128], [[0], [0]], None)
    batch_ids = O.offset2batch(torch.tensor(np.cumsum(sizes)))
    g = torch.Generator().manual_seed(2)
    # (5, 7, 64) / (5, 6, 128) / (5, 8, 64): the stem shapes -> the sparse SIMT kernels (k_conv_fwd_narrow / k_conv_wgrad_narrow2)
    for k, cin, cout, nbr_dev in [(3, 32, 48, stages[0].nbr3), (5, 7, 16, stages[0].nbr5), (5, 7, 64, stages[0].nbr5),
                                  (5, 6, 128, stages[0].nbr5), (5, 8, 64, stages[0].nbr5)]:
        nbr = O.neighbor_table(batch_ids, O.grid_coords(coord), k)
        _check_grads(lambda x, w, b: ops.subm_conv(x, w, b, nbr_dev, stages[0].stride_alloc),
                     lambda x, w, b: O.subm_conv(x, nbr, w, b),
                     dict(x=torch.randn(sum(sizes), cin, generator=g), w=torch.randn(cout, k, k, k, cin, generator=g) * 0.1,
                          b=torch.randn(cout, generator=g)))


@pytest.mark.parametrize("H,D,qk_norm,rpe", [(2, 16, True, False), (4, 24, True, False), (2, 32, False, False), (1, 32, True, True)])
def test_patch_attention_kernel(H, D, qk_norm, rpe):
    ops, tk, mods = pkg("ops"), pkg("tokenizer"), pkg("modules")
    sizes = [256, 128] if rpe else [300, 90, 129, 1]          # RPE (non-flash branch) needs full patches
    coord, grid = lattice_clouds(sizes, 24, 6)
    stg = tk.tokenize(coord.to(DEV), sizes, ["hilbert"], [2], [128, 128], [[0], [0]], None)[0]
    off = torch.tensor(np.cumsum(sizes))
    ser = O.serialize(O.grid_coords(coord), O.offset2batch(off), ["hilbert"])
    plan = O.patch_plan(off, 128)
    C = H * D
    g = torch.Generator().manual_seed(3)
    attn = mods.SerializedAttention(C, H, 128, qk_norm=qk_norm, enable_rpe=rpe, enable_flash=not rpe)
    inputs = dict(x=torch.randn(sum(sizes), C, generator=g), wqkv=torch.randn(3 * C, C, generator=g) * C ** -0.5,
                  qg=torch.randn(D, generator=g), qb=torch.randn(D, generator=g),
                  kg=torch.randn(D, generator=g), kb=torch.randn(D, generator=g), table=torch.randn(93, H, generator=g) * 0.5)

    def cuda_fn(x, wqkv, qg, qb, kg, kb, table):
        qkv = ops.linear(x, wqkv)
        return ops.PatchAttnFn.apply(qkv, qg if qk_norm else None, qb if qk_norm else None, kg if qk_norm else None,
                                     kb if qk_norm else None, table if rpe else None, stg.plans[0], H, qk_norm, ops.NO_DROP,
                                     stg.grid, 15)

    def ref_fn(x, wqkv, qg, qb, kg, kb, table):
        P = {"qkv.weight": wqkv, "q_norm.weight": qg, "q_norm.bias": qb, "k_norm.weight": kg, "k_norm.bias": kb,
             "proj.weight": torch.eye(C), "proj.bias": torch.zeros(C)}
        y = O.patch_attention(x, P, "", H, ser["order"][0], ser["inverse"][0], plan, 128, qk_norm,
                              table if rpe else None, O.grid_coords(coord))
        return y + 0 * (qg.sum() + qb.sum() + kg.sum() + kb.sum() + table.sum())
    _check_grads(cuda_fn, ref_fn, inputs)


@pytest.mark.parametrize("H,D", [(2, 16), (4, 24), (2, 32)])
def test_cross_attention_kernel(H, D):
    ops, tk = pkg("ops"), pkg("tokenizer")
    sizes, toks = [300, 90, 1], [12, 30, 8]
    C = H * D
    g = torch.Generator().manual_seed(4)
    geo = tk.CrossGeo(torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32, device=DEV),
                      torch.tensor([0] + list(np.cumsum(toks)), dtype=torch.int32, device=DEV), 3, max(sizes), max(toks))
    inputs = dict(q=torch.randn(sum(sizes), C, generator=g), kv=torch.randn(sum(toks), 2 * C, generator=g),
                  qg=torch.randn(D, generator=g), qb=torch.randn(D, generator=g), kg=torch.randn(D, generator=g),
                  kb=torch.randn(D, generator=g))

    def ref_fn(q, kv, qg, qb, kg, kb):
        P = {"q.weight": torch.eye(C), "q.bias": torch.zeros(C), "kv.weight": torch.eye(2 * C), "kv.bias": torch.zeros(2 * C),
             "q_norm.weight": qg, "q_norm.bias": qb, "k_norm.weight": kg, "k_norm.bias": kb,
             "proj.weight": torch.eye(C), "proj.bias": torch.zeros(C)}
        return O.cross_attention(q, kv, P, "", H, torch.tensor(np.cumsum(sizes)), torch.tensor(np.cumsum(toks)), True)
    _check_grads(lambda q, kv, qg, qb, kg, kb: ops.CrossAttnFn.apply(q, kv, qg, qb, kg, kb, geo, H, True, ops.NO_DROP),
                 ref_fn, inputs)


def test_pool_unpool_kernels():
    ops = pkg("ops")
    g = torch.Generator().manual_seed(5)
    n, M, C = 1000, 300, 48
    cluster = torch.randint(0, M, (n,), generator=g)
    cluster[:M] = torch.arange(M)                                     # no empty cluster
    member = torch.sort(cluster, stable=True)[1]
    counts = torch.bincount(cluster, minlength=M)
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(counts, 0)])
    dm, dp, dc = member.int().to(DEV), ptr.int().to(DEV), cluster.int().to(DEV)
    _check_grads(lambda x: ops.segment_max(x, dp, dm, M), lambda x: O.segment_max(x[member], ptr),
                 dict(x=torch.randn(n, C, generator=g)))
    x = torch.randn(n, 3, generator=g)
    assert rel(ops.segment_mean(x.to(DEV), dp, dm, M), O.segment_mean(x[member], ptr)) < 1e-6
    _check_grads(lambda skip, child: ops.unpool_add(skip, child, dc, dp, dm), lambda skip, child: skip + child[cluster],
                 dict(skip=torch.randn(n, C, generator=g), child=torch.randn(M, C, generator=g)))


def test_loss_and_decode_kernels(units):
    ops, syn = pkg("ops"), pkg("synthetic")
    b = syn.make_batch(seed=3, sizes=[200, 77, 130])
    npts, B = b["npoints_in_batch"], 3
    N = sum(npts)
    g = torch.Generator().manual_seed(6)
    heat = torch.randn(N, 90, generator=g) * 3
    action = torch.randn(B, 217, generator=g) * 2
    off = torch.tensor([0] + list(np.cumsum(npts)), dtype=torch.int32, device=DEV)
    prob = torch.cat([p.reshape(-1) for p in b["disc_pos_probs"]]).to(DEV)
    gt = b["gt_actions"]
    for weights in ([0, 0, 0, 1.0], [0.3, 1.7, -0.5, 2.0]):
        hc, ac = heat.clone().requires_grad_(True), action.clone().requires_grad_(True)
        hg, ag = heat.to(DEV).requires_grad_(True), action.to(DEV).requires_grad_(True)
        lv = ops.policy_loss(hg, ag, prob, off, gt.to(DEV), B, 30, 72, 1.5, 0.7)
        xt = hc.view(N, 3, 30).permute(1, 0, 2)
        ref = O.policy_loss(xt, ac[:, :216].reshape(B, 72, 3), ac[:, -1], gt, b["disc_pos_probs"], npts, 1.5, 0.7)
        refv = torch.stack([ref["pos"], ref["rot"], ref["open"], ref["total"]])
        assert rel(lv, refv) < 1e-5
        w = torch.tensor(weights)
        (lv * w.to(DEV)).sum().backward()
        (refv * w).sum().backward()
        assert rel(hg.grad, hc.grad) < TOL and rel(ag.grad, ac.grad) < TOL
    # decode: final actions vs the oracle restatement of the reference decode (float64 output)
    fts = b["pc_fts"].to(DEV)
    out = ops.decode_actions(heat.to(DEV), fts, off, B, 30, 0.01, action.to(DEV), 72, 5, None)
    xt = heat.view(N, 3, 30).permute(1, 0, 2)
    pos = O.best_pos_from_logits(xt, b["pc_fts"][:, :3], npts, 15, 0.01)
    quat = O.euler_bins_to_quat(torch.argmax(action[:, :216].reshape(B, 72, 3), 1).numpy(), 5)
    assert out.dtype == torch.float64
    assert torch.equal(out[:, :3].cpu(), pos.double())
    assert np.allclose(out[:, 3:7].cpu().numpy(), quat, atol=1e-12)
    assert torch.equal(out[:, 7].cpu(), action[:, -1].double())
    out2 = ops.decode_actions(heat.to(DEV), fts, off, B, 30, 0.01, action.to(DEV), 72, 5, gt.to(DEV))
    assert torch.equal(out2[:, :3].cpu(), gt[:, :3].double())


def test_dropout_mask_consistency():
    """dropout is a counter-based mask: forward and backward must see the same mask, keep-rate ~ 1-p."""
    ops = pkg("ops")
    x = torch.ones(4000, 64, device=DEV, requires_grad=True)
    w = torch.eye(64, device=DEV)
    y = ops.linear(x, w, None, drop=(0.25, 1234, 7))
    keep = (y != 0)
    assert abs(keep.float().mean().item() - 0.75) < 0.01
    assert torch.allclose(y[keep], torch.full_like(y[keep], 1 / 0.75))
    y.sum().backward()
    assert torch.equal(x.grad != 0, keep)
    u = torch.randn(1000, 32, device=DEV, requires_grad=True)
    a = ops.act(u, "gelu", (0.5, 99, 3))
    a.sum().backward()
    dropped = (a == 0) & (torch.nn.functional.gelu(u).abs() > 1e-6)
    assert abs(dropped.float().mean().item() - 0.5) < 0.03 and bool((u.grad[dropped] == 0).all())
    assert bool((u.grad[~dropped & (u.detach().abs() > 1e-3)] != 0).all())


# ------------------------------------------------------------------------------------------------ full policy
def _run_policy(g, record=True):
    cfg, model = make_model(g)
    model = model.to(DEV).train()
    eng = model.engine()
    eng.forced_perms = g["perms"]
    eng.record = record
    batch = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in g["batch"].items()}
    final, losses = model(batch, compute_loss=True, compute_final_action=False)
    return cfg, model, eng, final, losses


@pytest.mark.parametrize("which", ["small", "prod", "big"])
def test_policy_forward_backward_matches_reference(which, gold_small, gold_prod, request):
    # "big": two keysteps of 4096 / 3000 points at production width (BASELINE config-2 per-keystep sizes)
    g = gold_small if which == "small" else (gold_prod if which == "prod" else request.getfixturevalue("gold_big"))
    cfg, model, eng, final, losses = _run_policy(g)
    for k in ("pos", "rot", "open", "total"):
        assert abs(losses[k].item() - g["losses"][k].item()) <= TOL * abs(g["losses"][k].item()), k
    for name, ref in g["inter"].items():
        got = eng.last[name]
        if which == "small":
            assert rel(got, ref["feat"]) < TOL, name
        else:
            assert tuple(got.shape) == ref["shape"]
            assert rel(got[:8], ref["feat_head"]) < TOL, name
            if "feat_tail" in ref:
                assert rel(got[-8:], ref["feat_tail"]) < TOL, name
            assert abs(got.double().abs().sum().item() - ref["feat_abs_sum"]) < TOL * ref["feat_abs_sum"], name
    ref_act = g["train_actions"].double()
    assert final.dtype == torch.float64 and torch.equal(final[:, :3].cpu(), ref_act[:, :3])      # pred_pos := gt
    assert torch.allclose(final[:, 3:7].cpu(), ref_act[:, 3:7], atol=1e-9) and rel(final[:, 7], ref_act[:, 7]) < TOL
    losses["total"].backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert set(grads) == set(g["grad_norms"])
    floor = 1e-6 * max(g["grad_norms"].values())
    errs = []
    for k, n in g["grad_norms"].items():
        gn = grads[k].norm().item()
        if n < floor:
            assert gn < 10 * floor, (k, gn, n)
        else:
            errs.append((abs(gn - n) / n, k))
    errs.sort()
    # typical gradient within 1e-4; the tail is bounded by max-pool argmax flips (see tests/test_oracle_golden.py)
    assert errs[len(errs) // 2][0] < 2e-5 and errs[-1][0] < (TOL if which == "small" else 2e-2), errs[-5:]
    for k, ref in g["grads"].items():
        if ref.norm().item() > floor:
            assert rel(grads[k], ref) < (TOL if which == "small" else 2e-2), k
    sd = model.state_dict()
    for k, ref in g["bn_after"].items():
        if "num_batches" in k:
            assert int(sd[k]) == int(ref)
        else:
            assert rel(sd[k], ref) < 1e-5, k


def test_policy_eval_final_action_matches_reference(gold_small):
    g = gold_small
    cfg, model = make_model(g)
    sd = model.state_dict()
    sd.update(g["bn_after"])
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    model.engine().forced_perms = g["perms_eval"]
    batch = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in g["batch"].items()}
    with torch.no_grad():
        final = model(batch)
    ref = g["final_actions_eval"]
    assert final.dtype == torch.float64 and final.shape == ref.shape
    assert torch.allclose(final[:, :3].cpu(), ref[:, :3], atol=1e-6)
    assert torch.allclose(final[:, 3:7].cpu(), ref[:, 3:7], atol=1e-9)
    assert rel(final[:, 7], ref[:, 7]) < TOL


def test_policy_vs_oracle_with_dropout_off_random_batch():
    """a fresh seeded batch (not a fixture): CUDA path vs the oracle restatement end to end."""
    cfgm, syn, testing = pkg("config"), pkg("synthetic"), pkg("testing")
    cfg = cfgm.small_model_config(attn_drop=0.0, proj_drop=0.0)
    cfg.action_config.dropout = 0.0
    model = pkg("policy").build_parameter_shell(cfg)
    testing.deterministic_reinit(model, 77)
    batch = syn.make_batch(seed=5, sizes=[513, 128, 40, 1000, 2])
    perms = torch.stack([torch.randperm(4, generator=torch.Generator().manual_seed(s)) for s in range(5)])
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        ref = O.policy_forward(sd, cfg, batch, perms=perms, train=True)
    model = model.to(DEV).train()
    model.engine().forced_perms = perms
    final, losses = model({k: v for k, v in batch.items()}, compute_loss=True, compute_final_action=False)
    for k in ("pos", "rot", "open", "total"):
        assert abs(losses[k].item() - ref["losses"][k].item()) <= TOL * abs(ref["losses"][k].item()), k


def test_checkpoint_roundtrip_strict(gold_small):
    cfg, model = make_model(gold_small)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    m2 = pkg("policy").build_parameter_shell(pkg("config").ConfigNode(gold_small["config"]))
    missing = m2.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys


# ------------------------------------------------------------------------------------------------ bf16 tensor-core mode
def test_policy_bf16_mode_tracks_fp32_reference(gold_prod):
    """tensor-core mode (bf16 activations, tcgen05 GEMMs) at production width (channel widths must be multiples of 64):
    losses within 2% of the fp32 reference golden values and gradients pointing the same way."""
    for g in (gold_prod,):
        cfg, model = make_model(g)
        model = model.to(DEV).train()
        eng = model.engine()
        eng.forced_perms = g["perms"]
        eng.precision = "bf16"
        batch = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in g["batch"].items()}
        final, losses = model(batch, compute_loss=True, compute_final_action=False)
        for k in ("pos", "rot", "open", "total"):
            assert abs(losses[k].item() - g["losses"][k].item()) <= 2e-2 * abs(g["losses"][k].item()), (k, losses[k].item())
        losses["total"].backward()
        grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
        assert set(grads) == set(g["grad_norms"])
        big = sorted(g["grad_norms"].items(), key=lambda kv: -kv[1])[:40]
        for k, n in big:
            gn = grads[k].norm().item()
            assert abs(gn - n) <= _bf16_norm_tol(k, grads[k]) * n, (k, gn, n)
        for k, ref in g["grads"].items():
            if ref.norm().item() > 1e-3 * big[0][1]:
                a, b = grads[k].flatten().double().cpu(), ref.flatten().double()
                cos = (a @ b / (a.norm() * b.norm())).item()
                # bf16 rounding through ~60 layers + atomics order: the deepest gradient (stem) wanders 0.92-0.97 run to run;
                # fp32 validation mode is the 1e-4 parity path
                assert cos > 0.88, (k, cos)


# ------------------------------------------------------------------------------------------------ native one-call executor
@pytest.mark.parametrize("which", ["small", "prod", "big"])
def test_native_step_matches_reference_fp32(which, gold_small, gold_prod, request):
    """lotus_policy_step (C++ executor: forward + loss + backward in one call) in fp32 validation mode against the
    reference golden vectors: losses 1e-4, every parameter gradient as tight as the per-op path."""
    g = gold_small if which == "small" else (gold_prod if which == "prod" else request.getfixturevalue("gold_big"))
    cfg, model = make_model(g)
    model = model.to(DEV).train()
    eng = model.engine()
    eng.forced_perms = g["perms"]
    batch = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in g["batch"].items()}
    final, lv = eng.native_step(batch, compute_loss=True, with_grad=True, compute_final_action=False)
    names = ["pos", "rot", "open", "total"]
    for i, k in enumerate(names):
        assert abs(lv[i].item() - g["losses"][k].item()) <= TOL * abs(g["losses"][k].item()), (k, lv[i].item())
    ref_act = g["train_actions"].double()
    assert torch.equal(final[:, :3].cpu(), ref_act[:, :3]) and torch.allclose(final[:, 3:7].cpu(), ref_act[:, 3:7], atol=1e-9)
    grads = {k: p.grad for k, p in model.named_parameters()}
    floor = 1e-6 * max(g["grad_norms"].values())
    errs = []
    for k, n in g["grad_norms"].items():
        gn = grads[k].norm().item()
        if n < floor:
            assert gn < 10 * floor, (k, gn, n)
        else:
            errs.append((abs(gn - n) / n, k))
    errs.sort()
    assert errs[len(errs) // 2][0] < 2e-5 and errs[-1][0] < (TOL if which == "small" else 2e-2), errs[-5:]
    for k, ref in g["grads"].items():
        if ref.norm().item() > floor:
            assert rel(grads[k], ref) < (TOL if which == "small" else 2e-2), k
    sd = model.state_dict()
    for k, ref in g["bn_after"].items():
        if "num_batches" in k:
            assert int(sd[k]) == int(ref)
        else:
            assert rel(sd[k], ref) < 1e-5, k


def test_native_step_bf16_and_eval(gold_prod, gold_small):
    g = gold_prod
    cfg, model = make_model(g)
    model = model.to(DEV).train()
    eng = model.engine()
    eng.forced_perms = g["perms"]
    eng.precision = "bf16"
    batch = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in g["batch"].items()}
    final, lv = eng.native_step(batch, compute_loss=True, with_grad=True, compute_final_action=False)
    for i, k in enumerate(["pos", "rot", "open", "total"]):
        assert abs(lv[i].item() - g["losses"][k].item()) <= 2e-2 * abs(g["losses"][k].item()), (k, lv[i].item())
    grads = {k: p.grad for k, p in model.named_parameters()}
    big = sorted(g["grad_norms"].items(), key=lambda kv: -kv[1])[:40]
    for k, n in big:
        assert abs(grads[k].norm().item() - n) <= _bf16_norm_tol(k, grads[k]) * n, (k, grads[k].norm().item(), n)
    # eval / inference through the nn.Module API with the native executor (fp32): final actions == reference
    g = gold_small
    cfg, model = make_model(g)
    sd = model.state_dict()
    sd.update(g["bn_after"])
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    model.engine().forced_perms = g["perms_eval"]
    model.engine().use_native = True
    with torch.no_grad():
        final = model({k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in g["batch"].items()})
    ref = g["final_actions_eval"]
    assert torch.allclose(final[:, :3].cpu(), ref[:, :3], atol=1e-6) and torch.allclose(final[:, 3:7].cpu(), ref[:, 3:7], atol=1e-9)
    assert rel(final[:, 7], ref[:, 7]) < TOL


def test_flat_adamw_matches_reference_optimizer():
    """three steps of the fused flat-buffer clip+AdamW against the oracle's restatement of train/optim/adamw.py:53-112
    with torch's clip_grad_norm_ (train_simple_policy.py:237-243), decayed / bias groups as optim/misc.py:13-55"""
    cfg = pkg("config").small_model_config()
    model = pkg("policy").build_parameter_shell(cfg)
    pkg("testing").deterministic_reinit(model, 5)
    model = model.to(DEV)
    opt = pkg("trainer").FlatAdamW(model, lr=2e-3, weight_decay=0.05, max_grad_norm=0.5)
    nd = opt.group_sizes[0]
    ref_p = [p.detach().cpu().clone() for p in opt.params]
    ref_m = [torch.zeros_like(p) for p in ref_p]
    ref_v = [torch.zeros_like(p) for p in ref_p]
    wds = [0.05] * nd + [0.0] * (len(ref_p) - nd)
    gen = torch.Generator().manual_seed(3)
    for step in range(1, 4):
        opt.zero_grad()
        amp = [0.02, 5.0, 0.3][step - 1]                     # below, far above and near the clip threshold
        grads = [torch.randn(p.shape, generator=gen) * amp / np.sqrt(p.numel()) for p in ref_p]
        for p, g in zip(opt.params, grads):
            p.grad.copy_(g)
        lr = 2e-3 * step / 3
        opt.step(lr)
        norm = O.adamw_step(ref_p, grads, ref_m, ref_v, step, lr, wds, betas=(0.9, 0.98), eps=1e-6, max_grad_norm=0.5)
        assert abs(opt.grad_norm() - norm) <= 1e-5 * norm
        for name, p, r in zip(opt.names, opt.params, ref_p):
            assert torch.allclose(p.detach().cpu(), r, rtol=2e-5, atol=1e-7), name
    sd = opt.state_dict()
    assert sd["state"][0]["step"] == 3 and len(sd["param_groups"]) == 2
    for i in (0, nd, len(ref_p) - 1):
        assert torch.allclose(sd["state"][i]["exp_avg"].cpu(), ref_m[i], rtol=2e-5, atol=1e-9)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"].cpu(), ref_v[i], rtol=2e-5, atol=1e-12)


def _clone_batch(b):
    out = {}
    for k, v in b.items():
        if isinstance(v, torch.Tensor):
            out[k] = v.clone()
        elif isinstance(v, list) and v and isinstance(v[0], torch.Tensor):
            out[k] = [t.clone() for t in v]
        else:
            out[k] = list(v) if isinstance(v, list) else v
    return out


@pytest.mark.parametrize("kind", ["adanorm", "concat", "ca_tokens", "planner_ca", "planner_adanorm"])
def test_variant_models_match_reference(kind, gold_variants):
    """§8a rows 15 / 19 / 20: AdaNorm (PDNorm), Concat, pose + step-id context tokens and the two motion planners through
    the CUDA path against the reference-minted fixtures (training step, then the eval / final-action pass)"""
    g = gold_variants[kind]
    cfg = pkg("config").ConfigNode(g["config"])
    model = pkg("policy").build_parameter_shell(cfg)
    assert type(model).__name__ == g["model_class"]
    pkg("testing").deterministic_reinit(model, g["reinit_seed"])
    model = model.to(DEV).train()
    eng = model.engine()
    eng.forced_perms, eng.record = g["perms"], True
    final, losses = model(_clone_batch(g["batch"]), compute_loss=True, compute_final_action=False)
    assert set(losses) == set(g["losses"])
    for k, ref in g["losses"].items():
        assert abs(losses[k].item() - ref.item()) <= TOL * abs(ref.item()), k
    f = eng.last["dec_feat"]
    assert tuple(f.shape) == g["dec_feat_shape"]
    assert rel(f[:16], g["dec_feat_head"]) < TOL
    assert abs(f.double().abs().sum().item() - g["dec_feat_abs_sum"]) < TOL * g["dec_feat_abs_sum"]
    ref_act = g["train_actions"].double()
    assert final.dtype == torch.float64 and final.shape == ref_act.shape
    assert torch.equal(final[..., :3].cpu(), ref_act[..., :3])                          # pred_pos := gt
    assert torch.allclose(final[..., 3:7].cpu(), ref_act[..., 3:7], atol=1e-9)
    assert rel(final[..., 7:], ref_act[..., 7:]) < TOL
    losses["total"].backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert {k for k, p in model.named_parameters() if p.grad is None} == set(g["params_without_grad"])
    floor = 1e-6 * max(g["grad_norms"].values())
    errs = []
    for k, n in g["grad_norms"].items():
        gn = grads[k].norm().item()
        if n < floor:
            assert gn < 10 * floor, (k, gn, n)
        else:
            errs.append((abs(gn - n) / n, k))
    errs.sort()
    assert errs[len(errs) // 2][0] < 2e-5 and errs[-1][0] < TOL, errs[-5:]
    for k, ref in g["grads"].items():
        if ref.norm().item() > floor:
            assert rel(grads[k], ref) < TOL, k
    sd = model.state_dict()
    for k, ref in g["bn_after"].items():
        assert rel(sd[k], ref) < 1e-5, k
    # eval pass (BatchNorm running statistics as left by the training step, like the fixture) + final-action decode
    model.eval()
    eng.forced_perms = g["perms_eval"]
    with torch.no_grad():
        ev = model(_clone_batch(g["batch"]))
    ref = g["final_actions_eval"]
    assert ev.dtype == torch.float64 and ev.shape == ref.shape
    assert torch.allclose(ev[..., :3].cpu(), ref[..., :3], atol=1e-6)
    assert torch.allclose(ev[..., 3:7].cpu(), ref[..., 3:7], atol=1e-9)
    assert rel(ev[..., 7:], ref[..., 7:]) < TOL


def test_planner_bf16_mode_runs_and_tracks_fp32():
    """MotionPlannerPTV3CA at production width in tensor-core mode: losses within bf16 tolerance of the fp32 path"""
    cfgm, syn, testing = pkg("config"), pkg("synthetic"), pkg("testing")
    cfg = cfgm.production_model_config(attn_drop=0.0, proj_drop=0.0, in_channels=4)
    cfg.model_class = "MotionPlannerPTV3CA"
    cfg.action_config.update(dropout=0.0, max_traj_len=5, pc_label_channels=64, traj_embed_size=64, txt_reduce="attn")
    model = pkg("policy").build_parameter_shell(cfg)
    testing.deterministic_reinit(model, 31)
    model = model.to(DEV).train()
    batch = syn.make_planner_batch(seed=8, sizes=[600, 300])
    perms = torch.stack([torch.randperm(4, generator=torch.Generator().manual_seed(s)) for s in range(5)])
    eng = model.engine()
    eng.forced_perms = perms
    out = {}
    for prec in ("fp32", "bf16"):
        eng.precision = prec
        for p in model.parameters():
            p.grad = None
        final, losses = model(_clone_batch(batch), compute_loss=True, compute_final_action=False)
        losses["total"].backward()
        out[prec] = ({k: v.item() for k, v in losses.items()}, model.act_proj_head.traj_embedding.weight.grad.clone())
        assert final.shape == (2, 5, 9)
    for k in out["fp32"][0]:
        assert abs(out["bf16"][0][k] - out["fp32"][0][k]) < 3e-2 * abs(out["fp32"][0][k]) + 1e-3, (k, out)
    a, b = out["fp32"][1].flatten().double(), out["bf16"][1].flatten().double()
    assert torch.dot(a, b) / (a.norm() * b.norm()) > 0.9


def test_disc_pos_prob_producer_matches_reference(units):
    """device-side soft position labels (batch_producer.cu) vs the reference-minted golden vector and vs the numpy
    restatement on a ragged batch: a degenerate axis (ground truth > 1 cm from every candidate -> one-hot at the first
    closest candidate), a one-point cloud, both heatmap types.  Tolerance 1e-6 relative (fp64 arithmetic, fp32 storage)."""
    ops, syn = pkg("ops"), pkg("synthetic")
    p = units["pos_codec"]
    xyz, gt = p["xyz"], p["gt"]
    off = torch.tensor([0, xyz.shape[0]], dtype=torch.int32, device=DEV)
    gta = torch.zeros(1, 7)
    gta[0, :3] = gt
    out = ops.disc_pos_prob(xyz.to(DEV).contiguous(), off, 1, xyz.shape[0], gta.to(DEV), 15, 0.01)
    ref = p["prob"].float().reshape(-1)
    assert torch.allclose(out.cpu(), ref, rtol=1e-6, atol=0), (out.cpu() - ref).abs().max()
    # ragged batch through the packed 7-channel feature tensor
    rng = np.random.default_rng(5)
    sizes = [257, 1, 1000, 40]
    clouds = [rng.uniform(-0.3, 0.3, size=(n, 3)).astype(np.float32) for n in sizes]
    gts = np.stack([c[rng.integers(0, len(c))] + rng.normal(0, 0.003, 3).astype(np.float32) for c in clouds]).astype(np.float32)
    gts[2, 1] = 5.0                                            # axis with no candidate within 1 cm -> fallback
    gts[3] = clouds[3][7]                                      # exact hit: distance 0 -> weight 1 / 1e-4
    fts = torch.from_numpy(np.concatenate([np.concatenate([c, rng.uniform(-1, 1, size=(len(c), 4)).astype(np.float32)], 1) for c in clouds]))
    gta = torch.zeros(len(sizes), 7)
    gta[:, :3] = torch.from_numpy(gts)
    off = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32, device=DEV)
    for kind in ("dist", "plain"):
        out = ops.disc_pos_prob(fts.to(DEV), off, len(sizes), max(sizes), gta.to(DEV), 15, 0.01, kind).cpu()
        o = 0
        for c, g in zip(clouds, gts):
            if kind == "dist":
                ref = syn.disc_gt_pos_prob(c, g, 0.01, 15)
            else:                                              # action_position_utils.py:19-30
                shift = np.arange(-15, 15) * 0.01
                d = np.abs(g[None, :, None] - (c[:, :, None] + shift[None, None, :]))
                d = np.transpose(d, (1, 0, 2)).reshape(3, -1)
                ref = np.zeros_like(d, dtype=np.float32)
                ref[d < 0.01] = 1
                for i in range(3):
                    if ref[i].sum() == 0:
                        ref[i, np.argmin(d[i])] = 1
                ref = ref / ref.sum(-1, keepdims=True)
            n = ref.size
            got = out[o: o + n].reshape(3, -1)
            assert torch.allclose(got, torch.from_numpy(ref.astype(np.float32)), rtol=1e-6, atol=0), kind
            assert torch.allclose(got.sum(-1), torch.ones(3), atol=1e-5)
            o += n
        assert o == out.numel()


def test_native_step_with_device_side_labels(gold_prod):
    """a training step fed only (pc_fts, txt, gt_actions) -- soft labels produced on the device -- gives the same losses as
    the same step fed the host-computed `disc_pos_probs` of the collate"""
    cfgm, syn, policy = pkg("config"), pkg("synthetic"), pkg("policy")
    cfg = cfgm.production_model_config(attn_drop=0.0, proj_drop=0.0)
    cfg.action_config.dropout = 0.0
    torch.manual_seed(3)
    model = policy.build_parameter_shell(cfg).to(DEV).train()
    eng = model.engine()
    eng.precision, eng.use_native = "bf16", True
    eng.forced_perms = torch.stack([torch.arange(4)] * 5)
    b = syn.make_batch(seed=11, sizes=[700, 300, 1200])
    full = model.prepare_batch(dict(b))
    _, l1 = eng.native_step(full, compute_loss=True, with_grad=False)
    l1 = l1.clone()
    slim = {k: v for k, v in b.items() if k != "disc_pos_probs"}
    slim = model.prepare_batch(slim)
    _, l2 = eng.native_step(slim, compute_loss=True, with_grad=False)
    assert abs(l1[0].item() - l2[0].item()) <= 1e-5 * abs(l1[0].item()), (l1, l2)          # position CE: the label tensors agree
    # (rot / open / total only see run-to-run bf16 noise: BatchNorm statistics are reduced with fp32 atomics)
    assert torch.allclose(l1, l2, rtol=2e-3, atol=1e-4), (l1, l2)


def test_native_step_accepts_float64_labels_and_strided_inputs():
    """the reference collate emits `disc_pos_probs` as float64 (numpy arithmetic with a float64 shift,
    utils/action_position_utils.py:7-46) and INTEGRATION.md tells users to feed collate batches as given.  The native step takes raw
    `const float*`: inputs are coerced (dtype, device, contiguity) before any pointer is taken.  Same losses as the float32 batch."""
    cfgm, syn, policy = pkg("config"), pkg("synthetic"), pkg("policy")
    cfg = cfgm.production_model_config(attn_drop=0.0, proj_drop=0.0)
    cfg.action_config.dropout = 0.0
    torch.manual_seed(3)
    model = policy.build_parameter_shell(cfg).to(DEV).train()
    eng = model.engine()
    eng.precision, eng.use_native = "bf16", True
    eng.forced_perms = torch.stack([torch.arange(4)] * 5)
    b = syn.make_batch(seed=13, sizes=[500, 300])
    _, l1 = eng.native_step(model.prepare_batch(dict(b)), compute_loss=True, with_grad=False)
    l1 = l1.clone()
    b64 = dict(b)
    b64["disc_pos_probs"] = [p.double() for p in b["disc_pos_probs"]]
    b64["gt_actions"] = b["gt_actions"].double()
    wide = torch.zeros(b["txt_embeds"].shape[0], 2 * b["txt_embeds"].shape[1])
    wide[:, ::2] = b["txt_embeds"]
    b64["txt_embeds"] = wide[:, ::2]                                  # non-contiguous view
    assert not b64["txt_embeds"].is_contiguous()
    _, l2 = eng.native_step(model.prepare_batch(b64), compute_loss=True, with_grad=False)
    assert abs(l1[0].item() - l2[0].item()) <= 1e-5 * abs(l1[0].item()), (l1, l2)
    assert torch.allclose(l1, l2, rtol=2e-3, atol=1e-4), (l1, l2)


def test_pipelined_input_prefetch_matches_synchronous_steps():
    """TrainStep with `next_batch` (H2D copy + tokeniser of the next batch on a side stream during the current step) walks the same
    trajectory as the synchronous steps: same losses over 4 optimizer steps from pinned host batches of different sizes"""
    cfgm, syn, policy, trainer = pkg("config"), pkg("synthetic"), pkg("policy"), pkg("trainer")
    batches = [syn.make_batch(seed=70 + i, sizes=sz) for i, sz in enumerate([[600, 300], [900], [200, 250, 400], [1000, 128]])]
    host = []
    for b in batches:
        b = {k: v for k, v in b.items() if k != "disc_pos_probs"}
        host.append({k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in b.items()})
    traj = {}
    for mode in ("sync", "pipelined"):
        cfg = cfgm.production_model_config(attn_drop=0.0, proj_drop=0.0)
        cfg.action_config.dropout = 0.0
        torch.manual_seed(5)
        model = policy.build_parameter_shell(cfg).to(DEV).train()
        eng = model.engine()
        eng.precision, eng.use_native = "bf16", True
        eng.forced_perms = torch.stack([torch.arange(4)] * 5)
        step = trainer.TrainStep(model, learning_rate=1e-4, warmup_steps=1)
        out = []
        for i in range(4):
            nb = host[i + 1] if (mode == "pipelined" and i + 1 < 4) else None
            losses = step(dict(host[i]), next_batch=nb)
            out.append(torch.stack([losses[k] for k in ("pos", "rot", "open", "total")]).clone())
        torch.cuda.synchronize()
        traj[mode] = torch.stack(out).cpu()
    assert torch.isfinite(traj["pipelined"]).all()
    assert torch.allclose(traj["sync"], traj["pipelined"], rtol=5e-3, atol=1e-3), (traj["sync"], traj["pipelined"])


def test_augmentation_producer_matches_reference_stream():
    """SURVEY §8f row 2 / VERDICT r1 #9: z-rotation, U(0, 2 mm) jitter, subsample gather, centring and feature assembly on the device,
    fed the reference stream's own random draws (indices, angle, noise): features equal the numpy result to one float32 ulp (the
    float64 sums behind the centroid are formed in a different order), all four fixture keysteps in ONE packed call"""
    from conftest import load_golden
    ops = pkg("ops")
    cases = load_golden("producer.pt")
    for use_color in (True, False):
        cs = [c for c in cases if c["use_color"] == use_color]
        raw_xyz = torch.cat([c["raw_xyz"] for c in cs]).to(DEV)
        raw_rgb = torch.cat([c["raw_rgb"] for c in cs]).to(DEV) if use_color else None
        roff = np.cumsum([0] + [c["raw_xyz"].shape[0] for c in cs])
        sel = torch.cat([c["point_idxs"] + int(roff[i]) for i, c in enumerate(cs)]).int().to(DEV)
        out_off = torch.tensor(np.cumsum([0] + [c["point_idxs"].numel() for c in cs]), dtype=torch.int32, device=DEV)
        angle = torch.tensor([c["angle"] for c in cs], dtype=torch.float64)
        noise = torch.cat([c["noise"] for c in cs])
        gt_in = torch.stack([c["gt_pos_in"] for c in cs])
        fts, cen, gt = ops.augment_keysteps(raw_xyz, raw_rgb, sel, out_off, angle, noise=noise, table_height=cs[0]["table_height"], gt_pos=gt_in)
        ref = torch.cat([c["pc_ft"] for c in cs])
        assert fts.shape == ref.shape
        assert torch.allclose(cen.cpu(), torch.stack([c["centroid"] for c in cs]), rtol=0, atol=1e-12), (cen.cpu(), [c["centroid"] for c in cs])
        assert torch.allclose(fts.cpu(), ref, rtol=0, atol=2e-7), (fts.cpu() - ref).abs().max()
        assert (fts.cpu() == ref).float().mean().item() > 0.999
        assert torch.allclose(cen.cpu(), torch.stack([c["centroid"] for c in cs]), rtol=0, atol=1e-12)
        assert torch.allclose(gt.cpu(), torch.stack([c["gt_pos_out"] for c in cs]), rtol=0, atol=2e-7)


def test_subsample_producer_draws_uniform_subsets():
    """device-side subsampling (simple_policy_dataset.py:281-294): exactly count[b] distinct indices inside cloud b, different per
    seed, every point roughly equally likely; then the device-generated jitter stays inside [0, 2 mm) and the clouds come out centred"""
    ops = pkg("ops")
    sizes, counts = [5000, 3000, 130, 9000], [4096, 2900, 125, 4096]
    raw_off = torch.tensor(np.cumsum([0] + sizes), dtype=torch.int32, device=DEV)
    hits = torch.zeros(sum(sizes))
    first = None
    for seed in range(40):
        sel, out_off = ops.subsample_keysteps(raw_off, counts, seed)
        s = sel.cpu().long()
        o = 0
        for b, (n, k) in enumerate(zip(sizes, counts)):
            part = s[o: o + k]
            o += k
            assert part.numel() == k and part.unique().numel() == k
            assert int(part.min()) >= int(raw_off[b]) and int(part.max()) < int(raw_off[b + 1])
        hits[s] += 1
        if first is None:
            first = s.clone()
        elif seed == 1:
            assert not torch.equal(first, s)
    # inclusion frequency ~ k / n per cloud
    o = 0
    for n, k in zip(sizes, counts):
        f = hits[o: o + n] / 40
        assert abs(f.mean().item() - k / n) < 1e-6 and f.std().item() < 0.12
        o += n
    g = torch.Generator().manual_seed(0)
    raw_xyz = (torch.rand(sum(sizes), 3, generator=g) + torch.tensor([0.0, 0.0, 0.76])).to(DEV)
    sel, out_off = ops.subsample_keysteps(raw_off, counts, 7)
    angle = torch.zeros(4, dtype=torch.float64)
    f0, c0, _ = ops.augment_keysteps(raw_xyz, None, sel, out_off, angle, noise=None, noise_amp=0.0, center=False)
    f1, c1, _ = ops.augment_keysteps(raw_xyz, None, sel, out_off, angle, noise=None, noise_amp=0.002, seed=11, center=False)
    d = (f1[:, :3] - f0[:, :3]).cpu()
    assert float(d.min()) >= -1e-7 and float(d.max()) < 0.002 + 1e-6 and 0.0008 < float(d.mean()) < 0.0012
    assert torch.equal(f0[:, 3], f1[:, 3])                                 # height is taken before the augmentation
    f2, c2, _ = ops.augment_keysteps(raw_xyz, None, sel, out_off, angle, noise_amp=0.002, seed=11, center=True)
    oc = out_off.cpu().tolist()
    for b in range(4):
        assert f2[oc[b]: oc[b + 1], :3].double().mean(0).abs().max().item() < 1e-6


def test_voxel_downsample_front_end():
    """SURVEY §8f row 4 (second half): workspace crop + 1 cm voxel down-sample of the merged camera clouds on the device against the
    numpy restatement of open3d's algorithm (oracle; parity unpinned: open3d is absent offline): same voxels in the same (first
    appearance) order, first-point colours and indices bit-equal, voxel means equal to float64 rounding (the sums are atomics).
    Edge cases: points outside the crop, many points per voxel, an empty crop."""
    ops = pkg("ops")
    g = np.random.default_rng(5)
    n = 4 * 128 * 128
    xyz = np.concatenate([g.uniform([-0.6, -0.8, 0.6], [0.9, 0.8, 1.8], size=(n // 2, 3)),
                          g.normal([0.2, 0.0, 0.9], 0.03, size=(n // 2, 3))]).astype(np.float32)       # half of it piles into a few hundred voxels
    xyz = xyz[g.permutation(n)]
    rgb = g.integers(0, 256, size=(n, 3)).astype(np.uint8)
    lo, hi = [-0.5, -0.6, 0.7505], [0.75, 0.6, 1.6]                    # workspace box with the table height folded into z_lo
    m_ref, c_ref, f_ref = O.voxel_downsample(xyz, rgb, lo, hi, 0.01)
    m, c, f = ops.voxel_downsample(torch.from_numpy(xyz).to(DEV), torch.from_numpy(rgb).to(DEV), lo, hi, 0.01)
    assert m.shape[0] == m_ref.shape[0] and 1000 < m.shape[0] < n // 2
    assert torch.equal(f.cpu().long(), torch.from_numpy(f_ref)) and torch.equal(c.cpu(), torch.from_numpy(c_ref))
    assert np.allclose(m.cpu().numpy(), m_ref, rtol=0, atol=1e-12)
    # every output point lies inside the crop and inside its own voxel
    assert bool(((m > torch.tensor(lo, device=DEV)) & (m < torch.tensor(hi, device=DEV))).all())
    # empty crop
    m0, c0, f0 = ops.voxel_downsample(torch.from_numpy(xyz).to(DEV), torch.from_numpy(rgb).to(DEV), [5, 5, 5], [6, 6, 6], 0.01)
    assert m0.shape[0] == 0 and f0.numel() == 0
    # then the rest of process_point_clouds (:224-251: height, centring, colour scaling) through the producer kernel on the voxel means
    out_off = torch.tensor([0, m.shape[0]], dtype=torch.int32, device=DEV)
    fts, cen, _ = ops.augment_keysteps(m.float(), c, None, out_off, torch.zeros(1, dtype=torch.float64), noise_amp=0.0, table_height=0.7505)
    assert fts.shape == (m.shape[0], 7) and abs(float(fts[:, :3].double().mean())) < 1e-6 and float(fts[:, 6].min()) > 0


def test_graph_replay_matches_eager_inference():
    """SURVEY §8f row 1 / VERDICT r1 #2: the whole inference forward (tokeniser + lotus_policy_step) captured as ONE CUDA graph per size
    bucket and replayed over static buffers gives the eager path's actions for DIFFERENT clouds of the bucket (bit-equal: the real rows go
    through the same kernels; rows between the actual sizes and the bucket bounds are inert), re-draws the shuffled orders per call, and
    falls back to the eager path when a cloud outgrows the bucket's stage bounds"""
    cfgm, syn, policy = pkg("config"), pkg("synthetic"), pkg("policy")
    torch.manual_seed(5)
    model = policy.build_parameter_shell(cfgm.production_model_config()).to(DEV).eval()
    eng = model.engine()
    eng.precision, eng.use_native = "bf16", True
    for B, sizes_list in ((1, [[3000], [2800], [2600]]), (3, [[900, 400, 700], [1000, 300, 650]])):
        eager, graphed = [], []
        for use_graph in (False, True):
            eng.use_graph = use_graph
            for i, sizes in enumerate(sizes_list):
                b = syn.make_batch(seed=300 + i, sizes=sizes, with_targets=False)
                eng.forced_perms = torch.stack([torch.randperm(4, generator=torch.Generator().manual_seed(10 * i + s)) for s in range(5)])
                with torch.no_grad():
                    out = model(dict(b), compute_loss=False, compute_final_action=True)
                (graphed if use_graph else eager).append(out.cpu().clone())
        for a, g in zip(eager, graphed):
            assert a.shape == g.shape == (B, 8) and torch.equal(a, g), (a, g)
    # (one capture per (batch, point bucket, token bucket) key: the five calls above fall into at most five buckets, every call replays)
    assert 2 <= eng.graph_stats["captures"] <= 5 and eng.graph_stats["replays"] == 5 and eng.graph_stats["fallbacks"] == 0
    # a cloud made of isolated voxels does not shrink under grid pooling: stage 1 outgrows its bound -> eager fallback, same answer
    g = np.random.default_rng(0)
    vox = g.permutation(40 * 40 * 20)[:1500]
    grid = np.stack([vox // 800, (vox // 20) % 40, vox % 20], 1) * 3                  # every voxel 3 cells from the next
    xyz = torch.from_numpy(grid.astype(np.float32) * 0.01 + 0.005)
    b = syn.make_batch(seed=1, sizes=[1500], with_targets=False)
    b["pc_fts"] = torch.cat([xyz - xyz.mean(0), b["pc_fts"][:, 3:]], 1)
    eng.forced_perms = None
    with torch.no_grad():
        eng.use_graph = True
        out_g = model(dict(b), compute_loss=False, compute_final_action=True).cpu()
    assert eng.graph_stats["fallbacks"] == 1 and torch.isfinite(out_g).all()
