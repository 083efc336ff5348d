"""Helpers shared by tests / bench (no reference access)."""
import torch


@torch.no_grad()
def deterministic_reinit(model, seed):
    """Re-draw every entry of `model.state_dict()` in name order from one seeded CPU generator.

    Weights ~ N(0, 0.05) (large enough that every layer contributes visibly to the output), norm scales
    ~ 1 + N(0, 0.1), biases ~ N(0, 0.05), BN running_var in [0.5, 1.5].  The same procedure applied to
    the reference model (here) and to the CUDA-backed module (GPU box) yields identical weights without
    shipping a weight file.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = model.state_dict()
    for name in sorted(sd.keys()):
        t = sd[name]
        if name.endswith("num_batches_tracked"):
            t.zero_()
            continue
        r = torch.randn(t.shape, generator=g, dtype=torch.float32)
        if name.endswith("running_var"):
            v = 1.0 + 0.5 * torch.tanh(r)
        elif name.endswith("running_mean"):
            v = 0.05 * r
        elif t.ndim == 1 and name.endswith("weight"):      # LayerNorm / BatchNorm scale
            v = 1.0 + 0.1 * r
        elif name.endswith("bias"):
            v = 0.05 * r
        else:
            v = 0.05 * r
        t.copy_(v.to(t.dtype))
    return model
