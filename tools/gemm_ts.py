import ctypes, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = importlib.import_module("robot-3dlotus_b200.lib").load()
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
for (M, N, K) in [(128, 128, 128), (128, 128, 512), (96652, 512, 128)]:
    x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
    y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); b = torch.randn(N, device="cuda")
    ncta = ((M + 127) // 128) * ((N + 127) // 128)
    ts = torch.zeros(ncta * 8, dtype=torch.int64, device="cuda")
    for _ in range(3):
        L.lotus_linear_fwd_bf16(P(x), P(w), P(b), None, P(y), None, M, N, K, 0, 0, 0.0, 0, 0, None)
    torch.cuda.synchronize()
    L.lotus_tc_debug_buffer(P(ts))
    L.lotus_linear_fwd_bf16(P(x), P(w), P(b), None, P(y), None, M, N, K, 0, 0, 0.0, 0, 0, None)
    torch.cuda.synchronize()
    L.lotus_tc_debug_buffer(None)
    t = ts.view(ncta, 8).cpu()
    rel = (t - t[:, :1]).float()
    print((M, N, K), "median ns since CTA start per stamp:", [int(v) for v in rel.median(0).values.tolist()],
          " kernel span us:", (t[:, 7].max() - t[:, 0].min()).item() / 1e3, " CTA life median us", rel[:, 7].median().item() / 1e3)
