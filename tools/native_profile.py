"""dump the per-call records of one profiled native bf16 training step (dev tool): name, args, device ms"""
import importlib, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = lambda n: importlib.import_module("robot-3dlotus_b200." + n)
cfgm, syn, policy, trainer, native = pkg("config"), pkg("synthetic"), pkg("policy"), pkg("trainer"), pkg("native")
dev = torch.device("cuda", 0)
model = policy.build_parameter_shell(cfgm.production_model_config()).to(dev).train()
eng = model.engine(); eng.precision = "bf16"; eng.use_native = True
step = trainer.TrainStep(model)
b = bench.to_device(bench.pack_batch(syn.make_batch(seed=2024, batch_size=32), pin=False, labels="device"), dev)
for _ in range(3):
    step(dict(b))
recs = []
for rep in range(3):
    recs.append(native.profile_step(lambda: step(dict(b))))
out = []
for i, (name, args, ms) in enumerate(recs[0]):
    best = min(r[i][2] for r in recs)
    out.append({"i": i, "name": name, "ints": [a for a in args if 0 <= a < (1 << 31)][:12], "ms": best})
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/native_profile.json", "w"))
print("calls", len(out), "sum ms", sum(o["ms"] for o in out))
