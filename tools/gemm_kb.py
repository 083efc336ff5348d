"""steady-state ns per k-block of the TMA-fed tensor-core GEMM for long reductions (dev tool)"""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("robot-3dlotus_b200.lib").load()
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
for (M, N, K) in [(976, 768, 3072), (3210, 512, 2048), (128, 128, 4096), (128, 64, 4096), (18944, 128, 4096), (18944, 64, 4096)]:
    x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
    y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); b = torch.randn(N, device="cuda")
    bn = 64 if N <= 64 else 128
    ncta = ((M + 127) // 128) * ((N + bn - 1) // bn)
    ts = torch.zeros(ncta * 8, dtype=torch.int64, device="cuda")
    f = lambda: L.lotus_linear_fwd_bf16(P(x), P(w), P(b), None, P(y), None, M, N, K, 0, 0, 0.0, 0, 0, None)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    L.lotus_tc_debug_buffer(P(ts)); f(); torch.cuda.synchronize(); L.lotus_tc_debug_buffer(None)
    t = ts.view(ncta, 8).cpu(); rel = (t - t[:, :1]).float(); med = [int(v) for v in rel.median(0).values.tolist()]
    nkb = K // 64
    print((M, N, K), "CTAs", ncta, "event us %.1f" % (e0.elapsed_time(e1) * 1e3), "ns/kb %.0f" % ((med[3] - med[2]) / (nkb - 1)), "TFLOP/s %.0f" % (2.0 * M * N * K / (e0.elapsed_time(e1) * 1e-3) / 1e12))
