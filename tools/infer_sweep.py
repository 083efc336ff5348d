#!/usr/bin/env python
"""Batched inference sweep (BASELINE.json configs[4]): points-per-keystep x batch, eval mode,
`compute_final_action=True`, p50 latency and throughput of `model(batch)` through the public API with pinned host
inputs (H2D inside the timed region, D2H of the (B,8) actions).

    python tools/infer_sweep.py [--points 1024,2048,4096,8192,16384] [--batches 1,2,4,...,256] [--iters 30] > profiles/rN_infer_sweep.json

Multi-GPU inference is replicas only: under torchrun every rank runs the same sweep on its own GPU with
its share of the batch (batch / world per rank); rank 0 reports the max-over-ranks latency and the summed throughput.
"""
import argparse
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

pkg = lambda n: importlib.import_module("robot-3dlotus_b200." + n)  # noqa: E731


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", default="1024,2048,4096,8192,16384")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--max-total-points", type=int, default=(1 << 22) + 1, help="skip cells whose packed batch exceeds this many points per rank")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfgm, syn, policy = pkg("config"), pkg("synthetic"), pkg("policy")
    torch.manual_seed(2024)
    model = policy.build_parameter_shell(cfgm.production_model_config()).to(dev).eval()
    model.engine().precision = "bf16"
    model.engine().use_native = True
    native = pkg("native")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, tf_peak = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops_sustained", 1400.0)
    cells = []
    clouds = {}
    for n in [int(x) for x in args.points.split(",")]:
        # a small pool of distinct clouds of n points, tiled to the batch size (surface-like occupancy as in synthetic.py)
        clouds[n] = syn.make_batch(seed=900 + n, sizes=[n] * 4, with_targets=False, max_points=n)
    for n in [int(x) for x in args.points.split(",")]:
        base = clouds[n]
        npts = base["npoints_in_batch"]
        offs = [0]
        for c in npts:
            offs.append(offs[-1] + c)
        toff = [0]
        for l in base["txt_lens"]:
            toff.append(toff[-1] + l)
        for B in [int(x) for x in args.batches.split(",")]:
            Bl = max(1, B // world)
            if Bl * n > args.max_total_points:
                continue
            ids = [i % 4 for i in range(Bl)]
            fts = torch.cat([base["pc_fts"][offs[i]: offs[i + 1]] for i in ids]).pin_memory()
            txt = torch.cat([base["txt_embeds"][toff[i]: toff[i + 1]] for i in ids]).pin_memory()
            batch = {"pc_fts": fts, "npoints_in_batch": [npts[i] for i in ids], "txt_embeds": txt,
                     "txt_lens": [base["txt_lens"][i] for i in ids],
                     "offset": torch.tensor([npts[i] for i in ids]).cumsum(0)}
            lat = []
            try:
                with torch.no_grad():
                    model(dict(batch), compute_loss=False, compute_final_action=True)
            except torch.OutOfMemoryError:
                torch.cuda.empty_cache()
                cells.append({"points": n, "batch": Bl * world, "batch_per_gpu": Bl, "skipped": "out of device memory"})
                continue
            with torch.no_grad():
                for it in range(args.warmup + args.iters):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    act = model(dict(batch), compute_loss=False, compute_final_action=True).cpu()
                    dt = time.perf_counter() - t0
                    if it >= args.warmup:
                        lat.append(dt * 1e3)
            lat.sort()
            p50 = lat[len(lat) // 2]
            t = torch.tensor([p50], device=dev, dtype=torch.float64)
            if world > 1:
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            p50 = float(t.item())
            # roofline column: algorithmic bytes / FLOPs of every kernel call of ONE forward of this cell (native profile mode: the
            # arguments of each launcher call; same read-once / write-once accounting as bench.py), floor = sum over calls of
            # max(bytes / HBM peak, FLOPs / tensor peak); device_ms = the same calls' summed CUDA-event time
            floor_ms = dev_ms = None
            if rank == 0:
                with torch.no_grad():
                    rec = native.profile_step(lambda: model(dict(batch), compute_loss=False, compute_final_action=True))
                stages = model.engine().last_stages
                ctx = {"conv_pairs": {}, "n_of_upper": {}, "npad_of_patches": {}, "n_of_cloud_off": {}, "Lt": int(sum(batch["txt_lens"]))}
                for stg in stages:
                    pl = next(iter(stg.plans.values()))
                    ctx["npad_of_patches"][pl.n_patches] = pl.n_pad
                    ctx["n_of_cloud_off"][stg.cloud_off.data_ptr()] = stg.n
                floor_ms = dev_ms = 0.0
                for name, a_, ms_ in rec:
                    fl, by = bench.algorithmic_work(name, a_, ctx)
                    floor_ms += max(fl / (tf_peak * 1e12), by / (hbm_peak * 1e9)) * 1e3
                    dev_ms += ms_
            cells.append({"points": n, "batch": Bl * world, "batch_per_gpu": Bl, "p50_ms": p50, "p50_ms_per_keystep": p50 / (Bl * world) * world,
                          "keysteps_per_s": Bl * world / (p50 * 1e-3), "h2d_bytes": bench.h2d_bytes(batch), "finite": bool(torch.isfinite(act).all()),
                          "roofline_floor_ms": floor_ms, "device_ms_profiled": dev_ms,
                          "frac_of_roofline": (floor_ms / p50) if floor_ms else None})
            if rank == 0:
                print(json.dumps(cells[-1]), file=sys.stderr)
    if rank == 0:
        print(json.dumps({"workload": "3D-LOTUS inference sweep, SimplePolicyPTV3CA (68.2M params), bf16 tensor-core mode, eval, "
                                      "compute_final_action=True, synthetic voxel1cm clouds", "n_gpus": world,
                          "roofline": "roofline_floor_ms = sum over the kernel calls of one forward of max(algorithmic bytes / %.0f GB/s, FLOPs / %.0f TFLOP/s) "
                                      "(MEASURED_PEAKS.json; conv FLOPs counted dense over 27 taps); frac_of_roofline = floor / p50 (host in -> host out); rank 0's replica" % (hbm_peak, tf_peak),
                          "timing": "host wall clock around model(batch).cpu() with pinned host inputs, p50 over %d iterations after %d warm-ups; "
                                    "multi-GPU = independent replicas, max latency over ranks" % (args.iters, args.warmup),
                          "cells": cells}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
