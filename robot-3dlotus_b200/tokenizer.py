"""Integer tokeniser driver: one stream-ordered pass that builds EVERY stage's structure (grid coords,
curve codes, orders, patch plans, pooling clusters, sparse-conv neighbourhoods) from the input coordinates
with no host synchronisation, followed by ONE device->host copy of the data-dependent sizes.

Reference behaviour replaced (relative to /root/reference/genrobo3d/models/PointTransformerV3/):
Point.serialization model.py:83-138, Point.sparsify :140-176 (+ spconv indice generation),
SerializedAttention.get_padding_and_inverse :410-466, SerializedPooling.forward (integer half) :713-771.
The reference syncs the host at :102, :164-166, :173, :440 (per cloud) — SURVEY D.3.
"""
import ctypes
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from .lib import load
from .ops import _p, _st

CURVE_ID = {"z": 0, "z-trans": 1, "hilbert": 2, "hilbert-trans": 3}
META_STRIDE = 8


@dataclass
class PatchPlan:
    gather_idx: torch.Tensor      # (n_pad,) int32: order[pad]
    slot_of_point: torch.Tensor   # (n,) int32: unpad[inverse]
    cu_seqlens: torch.Tensor      # (n_patches+1,) int32
    pad: torch.Tensor
    unpad: torch.Tensor
    n_pad: int = 0
    n_patches: int = 0
    max_len: int = 128


@dataclass
class CrossGeo:
    pt_off: torch.Tensor
    tok_off: torch.Tensor
    B: int
    max_points: int
    max_tokens: int


@dataclass
class Stage:
    n: int = 0
    depth: int = 0
    stride_alloc: int = 0                      # row stride of code/order/inverse/nbr allocations
    grid: torch.Tensor = None                  # (n,3) int32
    batch: torch.Tensor = None                 # (n,) int32
    cloud_off: torch.Tensor = None             # (B+1,) int32 device
    cloud_off_host: List[int] = None
    code: torch.Tensor = None                  # (4, stride) int64
    order: torch.Tensor = None                 # (4, stride) int32
    inverse: torch.Tensor = None
    curves: List[int] = None
    nbr3: torch.Tensor = None                  # (27, stride) int32
    pairs_out: torch.Tensor = None             # (27, stride) int32 compact rulebook: output voxel of the j-th pair of tap d
    pairs_in: torch.Tensor = None              # (27, stride) int32 ... and its input voxel
    pair_cnt: torch.Tensor = None              # (27,) int32 pairs per tap (device)
    nbr5: Optional[torch.Tensor] = None        # (125, stride) int32, stage 0 only
    plans: Dict[int, PatchPlan] = field(default_factory=dict)
    # pooling from the parent stage
    cluster: Optional[torch.Tensor] = None     # (n_parent,) int32 : parent point -> this stage's point
    seg_ptr: Optional[torch.Tensor] = None     # (n+1,) int32 : CSR over the parent's code-sorted order (row 0)
    member: Optional[torch.Tensor] = None      # parent's order row 0
    head: Optional[torch.Tensor] = None

    @property
    def max_cloud(self):
        c = self.cloud_off_host
        return max(b - a for a, b in zip(c[:-1], c[1:]))


def _next_pow2(x):
    p = 1
    while p < x:
        p <<= 1
    return p


class PendingTokens:
    """kernels of one tokenisation in flight on `stream`; `finish()` waits for the size copy and slices the stage tables"""

    def __init__(self):
        self.tensors = []          # every device allocation of this tokenisation (for record_stream when it ran on a side stream)
        self.stages = None
        self.event = None
        self.host = None
        self.S = self.B = 0


def tokenize(coord, npoints: List[int], orders, strides, patch_sizes, rows_needed, perms=None, grid_size=0.01,
             div_mode=0, need_stem=True, need_pairs=True):
    """launch + finish in one go (the single host sync of a forward)"""
    return tokenize_finish(tokenize_launch(coord, npoints, orders, strides, patch_sizes, rows_needed, perms, grid_size, div_mode,
                                           need_stem, need_pairs))


def tokenize_launch(coord, npoints: List[int], orders, strides, patch_sizes, rows_needed, perms=None, grid_size=0.01,
                    div_mode=0, need_stem=True, need_pairs=True, static=None):
    """coord (N,3) fp32 cuda (may be a strided view of pc_fts); npoints: host list of per-cloud counts.
    perms: list of S row permutations (host lists) or None.  Returns [Stage_0 .. Stage_{S-1}].
    Only the serialization rows a stage uses are computed (rows 0..max(rows_needed[s]); row 0 also drives the pooling), and the
    compact pair lists (weight-gradient rulebook) only when `need_pairs` (training)."""
    L = load()
    dev = coord.device
    st = _st()
    S = len(strides) + 1
    B = len(npoints)
    N0 = int(sum(npoints))
    if static is not None:
        # shape-static form (CUDA-graph capture with size buckets): `coord` is a static buffer of `n_upper` rows, the ACTUAL cloud
        # offsets live in the device tensor static["off0"] (B+1 int32; the point count is read on the device), and the stage tables
        # are used over static["bounds"][s] rows; no host copy, no size read-back
        N0 = int(coord.shape[0])
    i32 = dict(device=dev, dtype=torch.int32)
    coordc = coord if coord.is_contiguous() else coord.contiguous()

    off0 = [0]
    for c in npoints:
        off0.append(off0[-1] + int(c))
    cloud_off_all = torch.empty((S, B + 1), **i32)
    if static is not None:
        cloud_off_all[0].copy_(static["off0"])
    else:
        cloud_off_all[0].copy_(torch.tensor(off0, dtype=torch.int32), non_blocking=True)
    meta = torch.zeros((S, META_STRIDE), **i32)
    scratch = torch.empty(4 * (B + 2) + 8, **i32)
    err = torch.zeros(1, **i32)
    max_cloud = max(npoints) if static is None else int(static["cloud_bounds"][0])
    hash_cap = _next_pow2(2 * N0)
    pair_scratch = torch.empty(27 * ((N0 + 1023) // 1024 + 1), **i32)
    hkeys = torch.empty(hash_cap, device=dev, dtype=torch.int64)
    hvals = torch.empty(hash_cap, **i32)

    stages: List[Stage] = []
    grid = torch.empty((N0, 3), **i32)
    batch = torch.empty(N0, **i32)
    if static is not None:
        L.lotus_grid_coord_n(_p(coordc), N0, _p(cloud_off_all[0][B:]), grid_size, div_mode, _p(grid), _p(meta[0]), _p(scratch), st)
    else:
        L.lotus_grid_coord(_p(coordc), N0, grid_size, div_mode, _p(grid), _p(meta[0]), _p(scratch), st)
    L.lotus_batch_ids(_p(cloud_off_all[0]), B, _p(batch), st)
    curves = [CURVE_ID[o] for o in orders]
    R = len(curves)
    for s in range(S):
        if perms is not None:
            curves = [curves[int(i)] for i in perms[s]]
        stg = Stage(stride_alloc=N0, grid=grid, batch=batch, cloud_off=cloud_off_all[s], curves=list(curves))
        stg.code = torch.empty((R, N0), device=dev, dtype=torch.int64)
        stg.order = torch.empty((R, N0), **i32)
        stg.inverse = torch.empty((R, N0), **i32)
        cv = (ctypes.c_int32 * 4)(*(curves + [0] * (4 - R)))
        nrows = min(R, max([0] + list(rows_needed[s])) + 1)
        if static is not None and static.get("curves_dev") is not None:
            # graph replay: this forward's shuffled curve ids are read from the static device array (S, 4)
            L.lotus_serialize_dev(_p(grid), _p(batch), _p(stg.cloud_off), B, N0, max_cloud, _p(meta[s]), _p(static["curves_dev"][s]), nrows, N0,
                                  _p(stg.code), _p(stg.order), _p(stg.inverse), _p(err), st)
        else:
            L.lotus_serialize(_p(grid), _p(batch), _p(stg.cloud_off), B, N0, max_cloud, _p(meta[s]), cv, nrows, N0, _p(stg.code),
                              _p(stg.order), _p(stg.inverse), _p(err), st)
        # sparse-conv neighbourhoods (hash built once per stage, probed for k=3 and, at stage 0, k=5)
        stg.nbr3 = torch.empty((27, N0), **i32)
        L.lotus_neighbor_table(_p(grid), _p(batch), _p(meta[s]), N0, 3, _p(hkeys), _p(hvals), hash_cap, 1, _p(stg.nbr3), N0, st)
        if need_pairs:
            stg.pairs_out = torch.empty((27, N0), **i32)
            stg.pairs_in = torch.empty((27, N0), **i32)
            stg.pair_cnt = torch.empty(27, **i32)
            L.lotus_pair_lists(_p(stg.nbr3), N0, _p(meta[s]), N0, 27, _p(stg.pairs_out), _p(stg.pairs_in), _p(stg.pair_cnt),
                               _p(pair_scratch), st)
        if s == 0 and need_stem:
            stg.nbr5 = torch.empty((125, N0), **i32)
            L.lotus_neighbor_table(_p(grid), _p(batch), _p(meta[s]), N0, 5, _p(hkeys), _p(hvals), hash_cap, 0, _p(stg.nbr5), N0, st)
        K = patch_sizes[s]
        npad_upper = N0 + B * K
        cu_len = N0 // K + B + 2
        for r in rows_needed[s]:
            pl = PatchPlan(gather_idx=torch.empty(npad_upper, **i32), slot_of_point=torch.empty(N0, **i32),
                           cu_seqlens=torch.empty(cu_len, **i32), pad=torch.empty(npad_upper, **i32),
                           unpad=torch.empty(N0, **i32), max_len=K)
            L.lotus_patch_plan(_p(stg.cloud_off), B, K, _p(stg.order[r]), _p(stg.inverse[r]), _p(pl.pad), _p(pl.unpad),
                               _p(pl.gather_idx), _p(pl.slot_of_point), _p(pl.cu_seqlens), cu_len, _p(meta[s]),
                               _p(scratch[8:]), st)
            stg.plans[r] = pl
        stages.append(stg)
        if s < S - 1:
            nxt_grid = torch.empty((N0, 3), **i32)
            nxt_batch = torch.empty(N0, **i32)
            cluster = torch.empty(N0, **i32)
            head = torch.empty(N0, **i32)
            seg_ptr = torch.empty(N0 + 1, **i32)
            L.lotus_pool_structure(_p(stg.code[0]), _p(stg.order[0]), _p(grid), _p(stg.cloud_off), B, int(strides[s]),
                                   _p(meta[s]), _p(cluster), _p(head), _p(seg_ptr), _p(nxt_grid), _p(nxt_batch),
                                   _p(cloud_off_all[s + 1]), _p(meta[s + 1]), _p(scratch), st)
            stg._next = (cluster, head, seg_ptr)
            grid, batch = nxt_grid, nxt_batch

    dev_sizes = torch.cat([meta.reshape(-1), cloud_off_all.reshape(-1), err])
    if static is not None:
        for s_, stg in enumerate(stages):
            nxt = getattr(stg, "_next", None)
            L.lotus_stage_sanitize(_p(meta[s_]), int(static["bounds"][s_]), _p(stg.nbr3), N0, _p(stg.nbr5) if stg.nbr5 is not None else None,
                                   _p(nxt[0]) if nxt is not None else None, _p(meta[s_ + 1]) if nxt is not None else None,
                                   int(static["bounds"][s_ + 1]) if nxt is not None else 0, _p(nxt[2]) if nxt is not None else None, st)
        pend = PendingTokens()
        pend.stages, pend.S, pend.B, pend.static, pend.dev_sizes = stages, S, B, static, dev_sizes
        return pend
    # ---- data-dependent sizes: one asynchronous copy into pinned memory + an event (the only host sync of the forward)
    pend = PendingTokens()
    pend.host = torch.empty(dev_sizes.shape, dtype=dev_sizes.dtype, pin_memory=True)
    pend.host.copy_(dev_sizes, non_blocking=True)
    pend.event = torch.cuda.Event()
    pend.event.record(torch.cuda.current_stream())
    pend.stages, pend.S, pend.B = stages, S, B
    pend.tensors = [coordc, cloud_off_all, meta, scratch, err, pair_scratch, hkeys, hvals, dev_sizes]
    for stg in stages:
        pend.tensors += [t for t in (stg.grid, stg.batch, stg.code, stg.order, stg.inverse, stg.nbr3, stg.pairs_out, stg.pairs_in,
                                     stg.pair_cnt, stg.nbr5) if t is not None]
        pend.tensors += [t for t in getattr(stg, "_next", ()) if t is not None]
        for pl in stg.plans.values():
            pend.tensors += [pl.gather_idx, pl.slot_of_point, pl.cu_seqlens, pl.pad, pl.unpad]
    return pend


def tokenize_finish_static(pend):
    """shape-static tokenisation: every stage is used over its bucket bound (rows past the actual count are inert, see
    lotus_stage_sanitize); the actual sizes stay on the device in `pend.dev_sizes` (meta, cloud offsets, error flag)"""
    stages, S, B, st = pend.stages, pend.S, pend.B, pend.static
    for s, stg in enumerate(stages):
        nb = int(st["bounds"][s])
        stg.n, stg.depth = nb, -1
        cb = int(st["cloud_bounds"][s])
        stg.cloud_off_host = [0, cb] + [cb] * (B - 1)           # only max_cloud is derived from it
        stg.grid, stg.batch = stg.grid[:nb], stg.batch[:nb]
        for pl in stg.plans.values():
            K = pl.max_len
            pl.n_pad, pl.n_patches = min(nb + B * K, pl.gather_idx.numel()), min(nb // K + B + 1, pl.cu_seqlens.numel() - 1)
            pl.slot_of_point, pl.unpad = pl.slot_of_point[:nb], pl.unpad[:nb]
        if s > 0:
            cluster, head, seg_ptr = stages[s - 1]._next
            stg.cluster = cluster[: stages[s - 1].n]
            stg.head = head[:nb]
            stg.seg_ptr = seg_ptr[: nb + 1]
            stg.member = stages[s - 1].order[0]
    return stages


def tokenize_finish(pend):
    """wait for the size copy of `tokenize_launch`, slice the tables to their true extents -> [Stage_0 .. Stage_{S-1}]"""
    if getattr(pend, "static", None) is not None:
        return tokenize_finish_static(pend)
    pend.event.synchronize()
    stages, S, B, host = pend.stages, pend.S, pend.B, pend.host
    meta_h = host[: S * META_STRIDE].view(S, META_STRIDE).tolist()
    off_h = host[S * META_STRIDE: S * META_STRIDE + S * (B + 1)].view(S, B + 1).tolist()
    if int(host[-1]) != 0:
        raise RuntimeError("lotus tokeniser: a cloud exceeds the shared-memory sort capacity (16384 points) or depth > 16")
    for s, stg in enumerate(stages):
        stg.n, stg.depth = meta_h[s][0], meta_h[s][1]
        stg.cloud_off_host = off_h[s]
        n = stg.n
        stg.grid, stg.batch = stg.grid[:n], stg.batch[:n]
        for pl in stg.plans.values():
            pl.n_pad, pl.n_patches = meta_h[s][2], meta_h[s][3]
            pl.gather_idx, pl.pad = pl.gather_idx[: pl.n_pad], pl.pad[: pl.n_pad]
            pl.slot_of_point, pl.unpad = pl.slot_of_point[:n], pl.unpad[:n]
            pl.cu_seqlens = pl.cu_seqlens[: pl.n_patches + 1]
        if s > 0:
            cluster, head, seg_ptr = stages[s - 1]._next
            stg.cluster = cluster[: stages[s - 1].n]
            stg.head = head[:n]
            stg.seg_ptr = seg_ptr[: n + 1]
            stg.member = stages[s - 1].order[0]
    assert stages[0].depth <= 16 and stages[0].depth * 3 + B.bit_length() <= 63, "serialization code overflow (model.py:105,110)"
    return stages
