"""micro-benchmark of the sparse-conv tensor-core kernels at the production stage shapes (dev tool)"""
import ctypes, importlib, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = lambda n: importlib.import_module("robot-3dlotus_b200." + n)
L = pkg("lib").load(); syn, tk = pkg("synthetic"), pkg("tokenizer")
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
dev = "cuda"
b = syn.make_batch(seed=2024, batch_size=32, with_targets=False)
sizes = [int(v) for v in b["npoints_in_batch"]]
xyz = b["pc_fts"][:, :3].contiguous().to(dev)
stages = tk.tokenize(xyz, sizes, ["z", "z-trans", "hilbert", "hilbert-trans"], [2, 2, 2, 2], [128] * 5, [[0]] * 5, None, need_pairs=True)
chans = [(0, 64), (0, 128), (1, 128), (2, 256), (3, 512), (4, 768)]
for s, C in chans:
    stg = stages[s]; n = stg.n
    x = torch.randn(n, C, device=dev).bfloat16(); w = (torch.randn(C, 27, C, device=dev) * 0.02).bfloat16()
    wt = w.permute(2, 1, 0).contiguous(); bias = torch.zeros(C, device=dev)
    y = torch.empty_like(x); dx = torch.empty_like(x); dw = torch.zeros(C, 27, C, device=dev)
    fns = {"fwd": lambda: L.lotus_subm_conv_fwd_bf16(P(x), P(w), P(bias), P(y), P(stg.nbr3), stg.stride_alloc, n, C, C, 27, 0, None),
           "dgrad": lambda: L.lotus_subm_conv_dgrad_bf16(P(x), P(wt), None, P(dx), P(stg.nbr3), stg.stride_alloc, n, C, C, 27, 0, None),
           "wgrad": lambda: L.lotus_subm_conv_wgrad_pairs_bf16(P(x), P(x), P(dw), P(stg.pairs_out), P(stg.pairs_in), P(stg.pair_cnt), stg.stride_alloc, n, C, C, 27, None)}
    res = {}
    only = os.environ.get("FNS")
    if only: fns = {k: v for k, v in fns.items() if k in only.split(",")}
    for name, fn in fns.items():
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 10 * 1e3
    pairs = int((stg.nbr3[:, :n] >= 0).sum())
    print(f"stage {s} n={n} C={C} pairs/pt={pairs / n:.1f}: " + "  ".join(f"{k} {v:7.1f}us" for k, v in res.items()) + (f"   useful {2.0 * pairs * C * C / res['fwd'] / 1e6:6.1f} TF/s" if "fwd" in res else ""))
