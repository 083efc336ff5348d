"""standalone timing of the north-star tokeniser kernels (FPS, kNN, ball query, serialisation + rulebook) on BASELINE config-2
shapes (32 keysteps of <= 4096 points): CUDA-event time per call, algorithmic bytes / time, and the intrinsic work rate.

    python tools/fps_knn_bench.py            # prints one JSON object (profiles/r2_fps_knn.json)

FPS is a chain of `npoint` dependent arg-max steps over a cloud held in shared memory (HBM traffic: the cloud once + the indices):
its meaningful rate is distance evaluations per second; kNN / ball query are brute force inside each cloud (n_q x n_r distance
evaluations from shared-memory tiles)."""
import ctypes, importlib, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = lambda n: importlib.import_module("robot-3dlotus_b200." + n)
L = pkg("lib").load()
syn, tk = pkg("synthetic"), pkg("tokenizer")
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
dev = "cuda"
b = syn.make_batch(seed=2024, batch_size=32, with_targets=False)
sizes = [int(v) for v in b["npoints_in_batch"]]
N, B, mx = sum(sizes), len(sizes), max(sizes)
xyz = b["pc_fts"][:, :3].contiguous().to(dev)
off = torch.tensor(np.cumsum([0] + sizes), dtype=torch.int32, device=dev)


def timed(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3     # us


out = {"workload": f"{B} keysteps, {N} points (<= {mx} per keystep), synthetic voxel1cm clouds", "kernels": {}}
npoint = 1024
start = torch.zeros(B, dtype=torch.int32, device=dev)
fidx = torch.empty(B * npoint, dtype=torch.int32, device=dev)
us = timed(lambda: L.lotus_fps(P(xyz), 3, P(off), B, mx, npoint, P(start), P(fidx), None), 5)
by = N * 12 + B * npoint * 4
out["kernels"]["lotus_fps"] = {"us": us, "npoint": npoint, "algorithmic_bytes": by, "gbs": by / us / 1e3, "distance_evals_per_s": N * npoint / us * 1e6,
                               "note": "one CTA per cloud (32 of 148 SMs busy), npoint dependent block-wide arg-max steps: latency-bound by construction"}
for k in (16, 32):
    idx = torch.empty(N * k, dtype=torch.int32, device=dev); dist = torch.empty(N * k, device=dev)
    us = timed(lambda: L.lotus_knn(P(xyz), P(off), P(xyz), P(off), B, mx, k, P(idx), P(dist), None), 5)
    by = 2 * N * 12 + N * k * 8
    ev = sum(s * s for s in sizes)
    out["kernels"][f"lotus_knn_k{k}"] = {"us": us, "algorithmic_bytes": by, "gbs": by / us / 1e3, "distance_evals_per_s": ev / us * 1e6}
bq = torch.empty(N * 32, dtype=torch.int32, device=dev)
us = timed(lambda: L.lotus_ball_query(P(xyz), P(off), P(xyz), P(off), B, mx, 0.05, 32, P(bq), None), 5)
by = 2 * N * 12 + N * 32 * 4
out["kernels"]["lotus_ball_query_r5cm_n32"] = {"us": us, "algorithmic_bytes": by, "gbs": by / us / 1e3}
# the model's own tokeniser: grid coords, 4 curve codes + per-cloud sort, pool structure, patch plans, rulebooks (all five stages)
perms = [[0, 1, 2, 3]] * 5
us = timed(lambda: tk.tokenize(xyz, sizes, ["z", "z-trans", "hilbert", "hilbert-trans"], [2, 2, 2, 2], [128] * 5, [[0]] * 5, perms, need_pairs=True), 10)
stages = tk.tokenize(xyz, sizes, ["z", "z-trans", "hilbert", "hilbert-trans"], [2, 2, 2, 2], [128] * 5, [[0]] * 5, perms, need_pairs=True)
by = 0
for s in stages:
    n = s.n
    by += n * (12 + 4 + 4 * 8 + 4 * 4 * 2 + 27 * 4 * 3 + 4 * 4) + (125 * 4 * n if s.nbr5 is not None else 0)
out["kernels"]["tokenize(5 stages, host orchestration + one size read-back)"] = {
    "us": us, "algorithmic_bytes": by, "gbs": by / us / 1e3, "stage_points": [s.n for s in stages],
    "note": "bytes = coords in, codes / orders / inverses, neighbour tables + pair lists, plans out; includes the Python launcher loop and one D2H size read"}
print(json.dumps(out))
