"""per-call GEMM timings by shape (dev tool): python-engine bf16 step with CUDA events around every C call"""
import importlib, sys, os, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = lambda n: importlib.import_module("robot-3dlotus_b200." + n)
cfgm, syn, policy, trainer, lib = pkg("config"), pkg("synthetic"), pkg("policy"), pkg("trainer"), pkg("lib")
dev = torch.device("cuda", 0)
model = policy.build_parameter_shell(cfgm.production_model_config()).to(dev).train()
eng = model.engine(); eng.precision = "bf16"
step = trainer.TrainStep(model)
b = bench.to_device(bench.pack_batch(syn.make_batch(seed=2024, batch_size=32), pin=False), dev)
for _ in range(3): step(dict(b))
L = lib.load()
prof = bench.CallProfiler(L)
with prof: step(dict(b))
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for name, args, e0, e1 in prof.records:
    ms = e0.elapsed_time(e1)
    if name in bench.GEMM_ARGS:
        i, j, k, es = bench.GEMM_ARGS[name]
        key = (name.replace("lotus_linear_", ""), int(args[i]), int(args[j]), int(args[k]))
    elif name in bench.CONV_ARGS:
        i = bench.CONV_ARGS[name]
        key = (name.replace("lotus_subm_conv_", "conv_"), int(args[i]), int(args[i + 1]), int(args[i + 2]))
    elif "attn" in name or "layernorm" in name:
        key = (name.replace("lotus_", ""),) + tuple(int(a) for a in args if isinstance(a, int) and 0 < a < 10**7)[:6]
    else:
        continue
    agg[key][0] += 1; agg[key][1] += ms
for key, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    extra = ""
    if key[0] in ("fwd_bf16", "dgrad_bf16", "wgrad_bf16"):
        M, N, K = key[1:]
        fl = 2.0 * M * N * K * n; by = 2.0 * (M * K + N * K + M * N) * n
        extra = f"  {fl / ms / 1e9:7.1f} TF/s  {by / ms / 1e6:7.1f} GB/s"
    print(f"{ms:7.3f} ms  x{n:<3d} {key}{extra}")
