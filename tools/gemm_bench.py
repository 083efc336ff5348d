"""micro-benchmark of the tcgen05 linear kernels by shape (dev tool): hot-loop CUDA-event timing"""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("robot-3dlotus_b200.lib").load()
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
dev = "cuda"
shapes = [(768, 256, 256), (96652, 128, 512), (96652, 512, 128), (96652, 128, 128), (45262, 128, 128), (12140, 256, 256),
          (12140, 1024, 256), (3210, 512, 512), (1100, 768, 3072), (1100, 3072, 768), (96652, 192, 64), (96652, 64, 64)]
for (M, N, K) in shapes:
    x = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16); b = torch.randn(N, device=dev)
    dy = torch.randn(M, N, device=dev).bfloat16(); wt = w.t().contiguous(); dx = torch.empty_like(x)
    dw = torch.zeros(N, K, device=dev)
    db = torch.zeros(N, device=dev); pre = torch.empty_like(y); u = torch.randn(M, K, device=dev).bfloat16(); resd = torch.randn(M, N, device=dev).bfloat16()
    res = {}
    for name, fn in (("fwd", lambda: L.lotus_linear_fwd_bf16(P(x), P(w), P(b), None, P(y), None, M, N, K, 0, 0, 0.0, 0, 0, None)),
                     ("dgrad", lambda: L.lotus_linear_dgrad_bf16(P(dy), P(wt), None, P(dx), M, N, K, 0, None)),
                     ("fwd_gelu2", lambda: L.lotus_linear_fwd_bf16(P(x), P(w), P(b), None, P(y), P(pre), M, N, K, 1, 0, 0.1, 1, 1, None)),
                     ("fwd_res", lambda: L.lotus_linear_fwd_bf16(P(x), P(w), P(b), P(resd), P(y), None, M, N, K, 0, 0, 0.1, 1, 1, None)),
                     ("dgrad_b", lambda: L.lotus_linear_dgrad_bias_bf16(P(dy), P(wt), None, P(dx), P(db), M, N, K, 0, None)),
                     ("dgrad_act_b", lambda: L.lotus_linear_dgrad_act_bias_bf16(P(dy), P(wt), P(u), P(dx), P(db), M, N, K, 1, 0.1, 1, 1, None)),
                     ("wgrad", lambda: L.lotus_linear_wgrad_bf16(P(dy), P(x), P(dw), M, N, K, None)),
                     ("cublas", lambda: torch.matmul(x, w.t()))):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 20 * 1e3
    by = 2.0 * (M * K + N * K + M * N)
    print(f"M={M:6d} N={N:5d} K={K:5d}  " + "  ".join(f"{k} {v:7.1f}us" for k, v in res.items()) +
          f"   floor(hbm) {by / 6.5e6:6.1f}us  fwd {by / res['fwd'] / 1e3:6.0f} GB/s {2.0 * M * N * K / res['fwd'] / 1e6:6.1f} TF/s")
