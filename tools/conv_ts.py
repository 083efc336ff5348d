"""per-CTA phase timestamps of the tensor-core sparse conv and a few linear shapes on a real tokenised batch (dev tool)"""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = lambda n: importlib.import_module("robot-3dlotus_b200." + n)
L = pkg("lib").load()
syn, tk = pkg("synthetic"), pkg("tokenizer")
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
b = syn.make_batch(seed=2024, batch_size=32)
fts = b["pc_fts"].cuda()
stages = tk.tokenize(fts[:, :3], b["npoints_in_batch"], ["z", "z-trans", "hilbert", "hilbert-trans"], [2, 2, 2, 2], [128] * 5, [[0]] * 5, None)
def stamp(fn, ncta, label, nkb):
    ts = torch.zeros(ncta * 8, dtype=torch.int64, device="cuda")
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    L.lotus_tc_debug_buffer(P(ts)); fn(); torch.cuda.synchronize(); L.lotus_tc_debug_buffer(None)
    t = ts.view(ncta, 8).cpu(); rel = (t - t[:, :1]).float()
    med = [int(v) for v in rel.median(0).values.tolist()]
    print(label, "event us %.1f" % (e0.elapsed_time(e1) * 1e3), "span us %.1f" % ((t[:, 7].max() - t[:, 0].min()).item() / 1e3), "CTAs", ncta,
          "median ns [setup, first_full, last_mma, epi_start, epi_smem, stored, end]:", med[1:], "mainloop ns/kb %.0f" % ((med[3] - med[2]) / max(nkb - 1, 1)))
for si, C in [(0, 64), (1, 128), (2, 256), (3, 512), (4, 768)]:
    stg = stages[si]; n = stg.n
    x = torch.randn(n, C, device="cuda").bfloat16(); w = torch.randn(C, 27 * C, device="cuda").bfloat16(); y = torch.empty(n, C, device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(C, device="cuda")
    bn = 64 if C <= 64 else 128
    ncta = ((n + 127) // 128) * ((C + bn - 1) // bn)
    stamp(lambda: L.lotus_subm_conv_fwd_bf16(P(x), P(w), P(bias), P(y), P(stg.nbr3), stg.stride_alloc, n, C, C, 27, 0, None), ncta, f"conv s{si} n={n} C={C}", 27 * C // 64)
n = stages[0].n
for (N, K, act) in [(256, 64, 1), (64, 256, 0), (192, 64, 0), (64, 64, 0)]:
    x = torch.randn(n, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16(); y = torch.empty(n, N, device="cuda", dtype=torch.bfloat16)
    u = torch.empty(n, N, device="cuda", dtype=torch.bfloat16); bias = torch.zeros(N, device="cuda")
    bn = 64 if N <= 64 else 128
    ncta = ((n + 127) // 128) * ((N + bn - 1) // bn)
    stamp(lambda: L.lotus_linear_fwd_bf16(P(x), P(w), P(bias), None, P(y), P(u) if act else None, n, N, K, act, 0, 0.1 if act else 0.0, 5, 3, None), ncta, f"linear M={n} N={N} K={K} act={act}", K // 64)
