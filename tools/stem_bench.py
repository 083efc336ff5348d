"""micro-benchmark of the fp32 stem convolution (5x5x5, Cin 7 -> 64) on bench-shaped surface clouds (dev tool)"""
import ctypes, importlib, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("robot-3dlotus_b200.lib").load()
tk = importlib.import_module("robot-3dlotus_b200.tokenizer")
syn = importlib.import_module("robot-3dlotus_b200.synthetic")
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
dev = "cuda"
B = int(os.environ.get("B", "32"))
batch = syn.make_batch(seed=2024, batch_size=B)
fts = batch["pc_fts"].to(dev).float().contiguous()
sizes = [int(v) for v in batch["npoints_in_batch"]]
n = fts.shape[0]
stg = tk.tokenize(fts[:, :3].contiguous(), sizes, ["z"], [2], [128, 128], [[0], [0]], None)[0]
nbr, stride = stg.nbr5, stg.stride_alloc
present = float((nbr[:, :n] >= 0).float().mean())
Cin, Cout, K3 = fts.shape[1], 64, 125
w = torch.randn(Cout, K3, Cin, device=dev) * 0.1
bias = torch.randn(Cout, device=dev)
y = torch.empty(n, Cout, device=dev); dy = torch.randn(n, Cout, device=dev); dw = torch.zeros_like(w)
side = torch.cuda.Stream(); ST = ctypes.c_void_p(side.cuda_stream)
def timeit(fn, reps=10):
    with torch.cuda.stream(side):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(reps): fn()
        e1.record(side); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
tf = timeit(lambda: L.lotus_subm_conv_fwd_f32(P(fts), P(w), P(bias), P(y), P(nbr), stride, n, Cin, Cout, K3, ST))
tw = timeit(lambda: L.lotus_subm_conv_wgrad_f32(P(dy), P(fts), P(dw), P(nbr), stride, n, Cin, Cout, K3, ST))
pairs = present * n * K3
print(f"stem conv n={n} Cin={Cin} Cout={Cout}: tap density {present:.3f} ({pairs / 1e6:.2f} M pairs)  fwd {tf:7.1f} us ({pairs * Cin * Cout * 2 / tf / 1e6:5.2f} TFLOP/s)  wgrad {tw:7.1f} us")
