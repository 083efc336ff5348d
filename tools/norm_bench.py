"""micro-benchmark of the LayerNorm / BatchNorm kernels at the stage shapes of the headline step (dev tool): hot-loop CUDA-event timing"""
import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("robot-3dlotus_b200.lib").load()
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
dev = "cuda"
reps = int(os.environ.get("REPS", "50"))
side = torch.cuda.Stream()
ST = ctypes.c_void_p(side.cuda_stream)
def timeit(fn):
    """device time per call: `reps` calls captured into one CUDA graph (the Python / ctypes call rate is ~9 us, above these kernels)"""
    with torch.cuda.stream(side):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        g.replay()
        e1.record(side); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for rows, C in [(96652, 64), (45262, 128), (12140, 256), (3210, 512), (1100, 512)]:
    x = torch.randn(rows, C, device=dev).bfloat16(); y = torch.empty_like(x); dy = torch.randn_like(x); dx = torch.empty_like(x)
    g = torch.randn(C, device=dev); b = torch.randn(C, device=dev); dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
    f = lambda: L.lotus_layernorm_fwd_bf16(P(x), 1, C, 0, P(y), 1, C, 0, None, P(g), P(b), P(mean), P(rstd), rows, C, 1e-5, ST)
    bw = lambda: L.lotus_layernorm_bwd_bf16(P(dy), 0, 1, C, 0, P(x), 1, C, 0, P(dx), 1, C, 0, P(g), P(mean), P(rstd), P(dg), P(db), rows, C, 0, ST)
    tf, tb = timeit(f), timeit(bw)
    print(f"LN rows={rows:6d} C={C:4d}: fwd {tf:6.1f} us ({rows * C * 4 / tf / 1e3:5.0f} GB/s)  bwd {tb:6.1f} us ({rows * C * 6 / tb / 1e3:5.0f} GB/s)")
for rows, C in [(96652, 64), (45262, 128), (12140, 256), (3210, 512)]:
    x = torch.randn(rows, C, device=dev).bfloat16(); y = torch.empty_like(x); dy = torch.randn_like(x); dx = torch.empty_like(x)
    g = torch.randn(C, device=dev); b = torch.randn(C, device=dev); dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev); sm = torch.empty(C, device=dev); sr = torch.empty(C, device=dev)
    scr = torch.zeros(4 * C, device=dev)
    f = lambda: L.lotus_batchnorm_fwd_bf16(P(x), 0, P(y), rows, C, P(g), P(b), P(rm), P(rv), P(sm), P(sr), P(scr), 1, 1e-3, 0.01, 1, ST)
    bw = lambda: L.lotus_batchnorm_bwd_bf16(P(dy), P(x), 0, rows, C, P(g), P(b), P(sm), P(sr), P(dx), P(dg), P(db), 1, 1, ST)
    tf, tb = timeit(f), timeit(bw)
    print(f"BN rows={rows:6d} C={C:4d}: fwd {tf:6.1f} us ({rows * C * 6 / tf / 1e3:5.0f} GB/s)  bwd {tb:6.1f} us ({rows * C * 10 / tb / 1e3:5.0f} GB/s)")
