"""host/GPU timing breakdown of one training step (dev tool)"""
import importlib, sys, time, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = lambda n: importlib.import_module("robot-3dlotus_b200." + n)
cfgm, syn, policy, trainer, tk, native = pkg("config"), pkg("synthetic"), pkg("policy"), pkg("trainer"), pkg("tokenizer"), pkg("native")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0)
model = policy.build_parameter_shell(cfgm.production_model_config()).to(dev).train()
eng = model.engine(); eng.precision = "bf16"; eng.use_native = True
step = trainer.TrainStep(model)
batches = [bench.to_device(bench.pack_batch(syn.make_batch(seed=2024 + i, batch_size=32), pin=False), dev) for i in range(4)]
for i in range(6):
    step(dict(batches[i % 4]))
torch.cuda.synchronize()
orig_tok = pkg("engine").tokenize
T = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        T.setdefault(name, []).append((1e3 * (t1 - t0), 1e3 * (t2 - t0)))
        return r
    return w
pkg("engine").tokenize = timed("tokenize", orig_tok)
eng._native.run = timed("native.run", eng._native.run)
step.opt.step = timed("opt.step", step.opt.step)
step.opt.zero_grad = timed("zero_grad", step.opt.zero_grad)
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step(dict(batches[i % 4]))
    torch.cuda.synchronize(); T.setdefault("step_total", []).append((0, 1e3 * (time.perf_counter() - t0)))
for k, v in T.items():
    host = sorted(x[0] for x in v)[len(v) // 2]; tot = sorted(x[1] for x in v)[len(v) // 2]
    print(f"{k:12s} host {host:8.2f} ms   host+gpu {tot:8.2f} ms")
