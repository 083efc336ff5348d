"""micro-benchmark of the fused patch-attention kernels by stage shape (dev tool): hot-loop CUDA-event timing"""
import ctypes, importlib, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("robot-3dlotus_b200.lib").load()
tk = importlib.import_module("robot-3dlotus_b200.tokenizer")
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
PO = lambda t, off: ctypes.c_void_p(t.data_ptr() + off)
dev = "cuda"
drop = float(os.environ.get("DROP", "0.1"))
cases = [(32, 3000, 2, 32), (32, 3000, 4, 32), (32, 1400, 4, 32), (32, 380, 8, 32), (32, 100, 16, 32), (32, 30, 32, 24)]
only = os.environ.get("CASE")
if only is not None:
    cases = [cases[int(only)]]
reps = int(os.environ.get("REPS", "20"))
for (B, npts, H, D) in cases:
    sizes = [npts] * B
    n = sum(sizes)
    rng = np.random.default_rng(0)
    # unique voxels per cloud on a lattice
    side = int(np.ceil(npts ** (1 / 3))) + 2
    coords = []
    for b in range(B):
        v = rng.permutation(side ** 3)[:npts]
        coords.append(np.stack([v // (side * side), (v // side) % side, v % side], 1))
    coord = torch.from_numpy(np.concatenate(coords)).float() * 0.01 + 0.005
    stg = tk.tokenize(coord.to(dev), sizes, ["z"], [2], [128, 128], [[0], [0]], None)[0]
    pl = stg.plans[0]
    C = H * D
    qkv = (torch.randn(n, 3 * C, device=dev) * 1.5).bfloat16()
    gq, bq, gk, bk = [torch.randn(D, device=dev) * 0.1 + (1 if i % 2 == 0 else 0) for i in range(4)]
    out = torch.empty(n, C, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(pl.n_pad * H, device=dev)
    dout = torch.randn(n, C, device=dev).bfloat16()
    dqkv = torch.empty_like(qkv)
    alias = torch.empty(pl.n_pad, 2, C, device=dev, dtype=torch.bfloat16)
    dg = [torch.zeros(D, device=dev) for _ in range(4)]
    es = 2
    def fwd():
        L.lotus_patch_attn_fused_fwd_bf16(P(qkv), PO(qkv, C * es), PO(qkv, 2 * C * es), 3 * C, 3 * C, 3 * C, P(gq), P(bq), P(gk), P(bk), P(pl.gather_idx),
                                          P(pl.slot_of_point), P(pl.cu_seqlens), pl.n_patches, pl.max_len, H, D, D ** -0.5, None, None, 0, P(out), C, P(lse),
                                          drop, 1234, 3, None)
    def bwd():
        L.lotus_patch_attn_fused_bwd_bf16(P(qkv), PO(qkv, C * es), PO(qkv, 2 * C * es), 3 * C, 3 * C, 3 * C, P(gq), P(bq), P(gk), P(bk), P(pl.gather_idx),
                                          P(pl.slot_of_point), P(pl.cu_seqlens), pl.n_patches, pl.n_pad, pl.max_len, H, D, D ** -0.5, None, None, 0, None,
                                          P(out), P(dout), C, P(lse), P(dqkv), PO(dqkv, C * es), PO(dqkv, 2 * C * es), 3 * C, 3 * C, 3 * C,
                                          P(dg[0]), P(dg[1]), P(dg[2]), P(dg[3]), P(alias), drop, 1234, 3, None)
    res = {}
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / reps * 1e3
    if os.environ.get("TS"):
        nct = pl.n_patches * ((H + 1) // 2)
        for name, fn in (("fwd", fwd), ("bwd", bwd)):
            buf = torch.zeros(nct * 8, dtype=torch.int64, device=dev)
            L.lotus_attn_debug_buffer(P(buf)); fn(); torch.cuda.synchronize(); L.lotus_attn_debug_buffer(None)
            ts = buf.view(nct, 8).double()
            d = (ts[:, 1:6] - ts[:, 0:5])
            ok = ts[:, 0] > 0
            print("   ", name, "load split: start->pt known %.0f | ->affine barrier %.0f | ->tiles staged %.0f" % (
                float((ts[ok, 6] - ts[ok, 0]).mean()), float((ts[ok, 7] - ts[ok, 6]).mean()), float((ts[ok, 1] - ts[ok, 7]).mean())))
            print("   ", name, "phase cycles (mean over CTAs): load %.0f | mmaA %.0f | elementwise %.0f | mmaB %.0f | epilogue %.0f | total %.0f" % (
                *[float(d[ok, i].mean()) for i in range(5)], float((ts[ok, 5] - ts[ok, 0]).mean())))
    byf = 2.0 * 4 * pl.n_pad * C; byb = 2.0 * 8 * pl.n_pad * C
    print(f"B={B} n={n} patches={pl.n_patches} H={H} D={D} drop={drop}: fwd {res['fwd']:7.1f} us ({byf / res['fwd'] / 1e3:6.0f} GB/s)  bwd {res['bwd']:7.1f} us ({byb / res['bwd'] / 1e3:6.0f} GB/s)")
