"""CPU fp32 stand-in for spconv_cu120==2.3.6 `spconv.pytorch` (test infrastructure).

Restates the documented semantics of the pieces the reference touches
(`model.py:167-174,235-251,616-622,845-852`):

* `SparseConvTensor(features, indices[(b,x,y,z) int32], spatial_shape, batch_size)`,
  `.features`, `.indices`, `.replace_feature`.
* `SubMConv3d`: submanifold convolution (cross-correlation):
  out[p] = sum_{d in k^3} W[:, d0, d1, d2, :] @ x[q],  q = the active voxel of the same batch
  index at coord[p] + d - k//2 (absent neighbours contribute nothing); output sites == input
  sites.  Weight layout [out, k0, k1, k2, in] (spconv >= 2.2 KRSC), kernel dims follow the order
  of the index columns (here x, y, z).  `padding` has no effect on a submanifold conv.
"""
import math
import torch
import torch.nn as nn
from . import modules  # noqa: F401


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, **kw):
        self.features = features
        self.indices = indices
        self.spatial_shape = list(spatial_shape)
        self.batch_size = batch_size
        self._rulebooks = kw.get("_rulebooks", {})

    def replace_feature(self, feat):
        return SparseConvTensor(feat, self.indices, self.spatial_shape, self.batch_size,
                                _rulebooks=self._rulebooks)


def build_neighbor_table(indices, k):
    """nbr[p, d] = index of the active voxel at coord[p]+off(d) in the same batch, or -1."""
    idx = indices.long()
    r = k // 2
    b = idx[:, 0]
    c = idx[:, 1:] + r  # shift so neighbours never go negative
    dims = c.max(0).values + r + 1
    def lin(bb, cc):
        return ((bb * dims[0] + cc[:, 0]) * dims[1] + cc[:, 1]) * dims[2] + cc[:, 2]
    key = lin(b, c)
    skey, sidx = torch.sort(key)
    assert (skey[1:] != skey[:-1]).all(), "duplicate voxels: submanifold conv undefined"
    n = idx.shape[0]
    nbr = torch.full((n, k ** 3), -1, dtype=torch.long)
    d = 0
    for d0 in range(k):
        for d1 in range(k):
            for d2 in range(k):
                off = torch.tensor([d0 - r, d1 - r, d2 - r])
                qk = lin(b, c + off)
                pos = torch.searchsorted(skey, qk).clamp(max=n - 1)
                hit = skey[pos] == qk
                nbr[:, d] = torch.where(hit, sidx[pos], torch.full_like(pos, -1))
                d += 1
    return nbr


class SubMConv3d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None, **kw):
        super().__init__()
        assert stride == 1 and dilation == 1 and groups == 1
        self.in_channels, self.out_channels, self.k = in_channels, out_channels, kernel_size
        self.indice_key = indice_key
        k = kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, k, k, k, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        fan_in = self.in_channels * self.k ** 3
        bound = 1.0 / math.sqrt(fan_in)  # kaiming_uniform_(a=sqrt(5)) == U(-1/sqrt(fan_in), +)
        nn.init.uniform_(self.weight, -bound, bound)
        if self.bias is not None:
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: SparseConvTensor):
        key = (self.indice_key, self.k)
        if self.indice_key is None or key not in x._rulebooks:
            nbr = build_neighbor_table(x.indices, self.k)
            if self.indice_key is not None:
                x._rulebooks[key] = nbr
        else:
            nbr = x._rulebooks[key]
        feat = x.features
        w = self.weight.reshape(self.out_channels, self.k ** 3, self.in_channels)
        out = feat.new_zeros(feat.shape[0], self.out_channels)
        for d in range(self.k ** 3):
            q = nbr[:, d]
            m = q >= 0
            if m.any():
                out = out.index_add(0, m.nonzero()[:, 0], feat[q[m]] @ w[:, d].t())
        if self.bias is not None:
            out = out + self.bias
        return x.replace_feature(out)
