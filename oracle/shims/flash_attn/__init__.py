"""fp32 CPU stand-in for flash-attn==2.5.9.post1 varlen entry points (test infrastructure).

Restates the published semantics: per-sequence softmax(q k^T * scale) v over the
segments given by cu_seqlens.  No dropout (the oracle runs with dropout 0 / eval).
The reference casts its inputs with `.half()` before the call; `oracle/ref_import.py`
neutralises that cast so the oracle is "flash-path semantics in fp32"
(SURVEY Appendix D.1).
"""
import torch

__version__ = "2.5.9.post1-oracle-shim"


def _seg_attn(q, k, v, scale):
    # q (Lq,H,D) k,v (Lk,H,D)
    s = torch.einsum("qhd,khd->hqk", q, k) * scale
    p = torch.softmax(s, dim=-1)
    return torch.einsum("hqk,khd->qhd", p, v)


def flash_attn_varlen_qkvpacked_func(qkv, cu_seqlens, max_seqlen, dropout_p=0.0, softmax_scale=None, **kw):
    assert dropout_p == 0, "oracle shim: run with attn_drop=0 or eval()"
    H, D = qkv.shape[2], qkv.shape[3]
    scale = softmax_scale if softmax_scale is not None else D ** -0.5
    cu = cu_seqlens.tolist()
    outs = []
    for a, b in zip(cu[:-1], cu[1:]):
        assert b - a <= max_seqlen
        outs.append(_seg_attn(qkv[a:b, 0], qkv[a:b, 1], qkv[a:b, 2], scale))
    return torch.cat(outs, 0)


def flash_attn_varlen_kvpacked_func(q, kv, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                                    dropout_p=0.0, softmax_scale=None, **kw):
    assert dropout_p == 0, "oracle shim: run with attn_drop=0 or eval()"
    D = q.shape[-1]
    scale = softmax_scale if softmax_scale is not None else D ** -0.5
    cq, ck = cu_seqlens_q.tolist(), cu_seqlens_k.tolist()
    outs = []
    for i in range(len(cq) - 1):
        outs.append(_seg_attn(q[cq[i]:cq[i + 1]], kv[ck[i]:ck[i + 1], 0], kv[ck[i]:ck[i + 1], 1], scale))
    return torch.cat(outs, 0)
