import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("robot-3dlotus_b200.lib").load()
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
M, N, K = [int(v) for v in sys.argv[1:4]]
x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); b = torch.randn(N, device="cuda")
for _ in range(3):
    L.lotus_linear_fwd_bf16(P(x), P(w), P(b), None, P(y), None, M, N, K, 0, 0, 0.0, 0, 0, None)
torch.cuda.synchronize()
