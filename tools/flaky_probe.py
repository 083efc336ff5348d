"""dev: repeat the bf16 native step of tests/test_gpu_parity.py::test_native_step_bf16_and_eval and print the worst deviations per run"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_parity as tp
g = torch.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "policy_production.pt"), weights_only=False)
DEV = "cuda"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
worst = []
for it in range(n):
    cfg, model = tp.make_model(g)
    model = model.to(DEV).train()
    eng = model.engine()
    eng.forced_perms = g["perms"]
    eng.precision = "bf16"
    batch = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in g["batch"].items()}
    final, lv = eng.native_step(batch, compute_loss=True, with_grad=True, compute_final_action=False)
    dl = max(abs(lv[i].item() - g["losses"][k].item()) / abs(g["losses"][k].item()) for i, k in enumerate(["pos", "rot", "open", "total"]))
    grads = {k: p.grad for k, p in model.named_parameters()}
    big = sorted(g["grad_norms"].items(), key=lambda kv: -kv[1])[:40]
    devs = sorted(((abs(grads[k].norm().item() - nn_) / nn_, k) for k, nn_ in big), reverse=True)
    dg = devs[0]
    dw = max((d for d in devs if grads[d[1]].dim() > 1), default=(0.0, ""))
    worst.append((dl, dg[0], dw[0]))
    if it < 3 or dg[0] > 0.05:
        print(f"run {it}: loss dev {dl:.4f}  worst {dg[0]:.4f} ({dg[1]})  worst weight {dw[0]:.4f} ({dw[1]})", flush=True)
print("runs", n, "max loss dev", max(w[0] for w in worst), "max grad-norm dev (all)", max(w[1] for w in worst), "(weights only)", max(w[2] for w in worst))
