"""Stand-in for torch-scatter==2.1.2 `segment_csr` (test infrastructure).

CSR segment reduction: out[m] = reduce(src[indptr[m]:indptr[m+1]]).  Only the
reductions the reference calls (`model.py:760-765`: max, mean) plus sum/min.
"""
import torch


def segment_csr(src, indptr, out=None, reduce="sum"):
    assert out is None
    reduce = {"sum": "sum", "add": "sum", "mean": "mean", "max": "max", "min": "min"}[reduce]
    return torch.segment_reduce(src, reduce, offsets=indptr, axis=0)
