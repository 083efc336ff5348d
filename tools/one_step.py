"""run N native bf16 training steps (dev tool for ncu captures)"""
import importlib, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = lambda n: importlib.import_module("robot-3dlotus_b200." + n)
cfgm, syn, policy, trainer = pkg("config"), pkg("synthetic"), pkg("policy"), pkg("trainer")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
model = policy.build_parameter_shell(cfgm.production_model_config()).to(dev).train()
eng = model.engine(); eng.precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"; eng.use_native = True
step = trainer.TrainStep(model)
b = bench.to_device(bench.pack_batch(syn.make_batch(seed=2024, batch_size=32), pin=False), dev)
for _ in range(n):
    step(dict(b))
torch.cuda.synchronize()
print("done", float(step.model.engine()._native.losses[3]))
