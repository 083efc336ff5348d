"""Stand-in for timm==1.0.9 `timm.models.layers.{DropPath, trunc_normal_}` (test infrastructure)."""
import torch
import torch.nn as nn


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    # timm's trunc_normal_ has the same (absolute a/b) semantics as torch's.
    return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask
