"""where does a pipelined e2e step spend host time? (dev tool)"""
import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = lambda n: importlib.import_module("robot-3dlotus_b200." + n)
cfgm, syn, policy, trainer = pkg("config"), pkg("synthetic"), pkg("policy"), pkg("trainer")
dev = torch.device("cuda", 0)
model = policy.build_parameter_shell(cfgm.production_model_config()).to(dev).train()
eng = model.engine(); eng.precision = "bf16"; eng.use_native = True
step = trainer.TrainStep(model)
host = [bench.pack_batch(syn.make_batch(seed=2024 + i, batch_size=32), pin=True, labels="device") for i in range(4)]
T = {}
def timed(obj, name):
    fn = getattr(obj, name)
    def w(*a, **k):
        t0 = time.perf_counter(); r = fn(*a, **k); T.setdefault(name, []).append(1e3 * (time.perf_counter() - t0)); return r
    setattr(obj, name, w)
timed(eng, "prefetch"); timed(eng, "native_step"); timed(step.opt, "step"); timed(step.opt, "zero_grad"); timed(model, "prepare_batch")
tk = pkg("tokenizer"); engm = pkg("engine")
for nm in ("tokenize_launch", "tokenize_finish"):
    f = getattr(engm, nm)
    def mk(f, nm):
        def w(*a, **k):
            t0 = time.perf_counter(); r = f(*a, **k); T.setdefault(nm, []).append(1e3 * (time.perf_counter() - t0)); return r
        return w
    setattr(engm, nm, mk(f, nm))
for mode in ("sync", "pipelined"):
    for k in T: T[k] = []
    for i in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        losses = step(dict(host[i % 4]), next_batch=host[(i + 1) % 4] if mode == "pipelined" else None)
        x = losses["total"].item()
        T.setdefault("step_total", []).append(1e3 * (time.perf_counter() - t0))
    print(mode, {k: round(sorted(v)[len(v) // 2], 2) for k, v in T.items() if v})
