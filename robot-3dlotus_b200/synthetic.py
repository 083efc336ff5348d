"""Synthetic voxel1cm-shaped keystep batches (there is no dataset offline).

Produces the batch schema `ptv3_collate_fn` emits
(reference: genrobo3d/train/datasets/simple_policy_dataset.py:391-415):
packed `pc_fts (N,7)` = xyz(centred, metres) + rgb[-1,1] + height, `offset`, `npoints_in_batch`,
`txt_embeds (sum L, 512)`, `txt_lens`, `gt_actions (B,7)`, `disc_pos_probs` (soft position labels).

Clouds are *surface-like* (table patch + box shells + a tube), one point per occupied 1 cm voxel, and are
de-duplicated under the model's own quantisation `trunc((coord - min_over_batch)/0.01)`
(reference: models/PointTransformerV3/model.py:96-98) because duplicate voxels make the submanifold
conv / argsort undefined in the reference.
"""
import numpy as np
import torch

VOXEL = 0.01


def disc_gt_pos_prob(xyz, gt_pos, pos_bin_size=0.01, pos_bins=15, heatmap_type="dist"):
    """soft position labels, restating genrobo3d/utils/action_position_utils.py:7-46 (no robot mask).
    xyz (n,3) f32/f64, gt_pos (3,) -> (3, n*2*pos_bins) float (dtype follows numpy promotion; 'plain' is float32)."""
    shift = np.arange(-pos_bins, pos_bins) * pos_bin_size
    cands = xyz[:, :, None] + shift[None, None, :]                # (n,3,2b)
    dists = np.abs(gt_pos[None, :, None] - cands)
    dists = np.transpose(dists, (1, 0, 2)).reshape(3, -1)          # 'n c b -> c (n b)'
    if heatmap_type == "plain":
        prob = np.zeros(dists.shape, dtype=np.float32)
        prob[dists < 0.01] = 1
    else:
        prob = 1.0 / np.maximum(dists, 1e-4)
        prob[dists > 0.01] = 0
    for i in range(3):
        if prob[i].sum() == 0:
            prob[i, np.argmin(dists[i])] = 1
    return prob / prob.sum(-1, keepdims=True)


def _box_shell(rng, lo, hi):
    """all surface voxels of an axis-aligned box [lo,hi] (integer voxel coords, inclusive)."""
    xs, ys, zs = [np.arange(lo[a], hi[a] + 1) for a in range(3)]
    faces = []
    for a, (u, v) in enumerate([(ys, zs), (xs, zs), (xs, ys)]):
        uu, vv = np.meshgrid(u, v, indexing="ij")
        for side in (lo[a], hi[a]):
            f = np.empty((uu.size, 3), dtype=np.int64)
            oth = [d for d in range(3) if d != a]
            f[:, a] = side
            f[:, oth[0]] = uu.ravel()
            f[:, oth[1]] = vv.ravel()
            faces.append(f)
    return np.concatenate(faces, 0)


def _tube(rng, base, radius, height):
    th = np.linspace(0, 2 * np.pi, int(2 * np.pi * radius * 2) + 8, endpoint=False)
    z = np.arange(height)
    tt, zz = np.meshgrid(th, z, indexing="ij")
    p = np.stack([base[0] + radius * np.cos(tt), base[1] + radius * np.sin(tt), base[2] + zz], -1)
    return np.round(p.reshape(-1, 3)).astype(np.int64)


def surface_cloud_voxels(rng, n_target, extent=(100, 100, 60)):
    """Unique integer voxels on 2-D surfaces inside a 1.0 x 1.0 x 0.6 m workspace at 1 cm."""
    ex = np.asarray(extent)
    parts = []
    # a table patch
    w = rng.integers(20, 50, size=2)
    o = rng.integers(0, ex[:2] - w)
    xx, yy = np.meshgrid(np.arange(o[0], o[0] + w[0]), np.arange(o[1], o[1] + w[1]), indexing="ij")
    parts.append(np.stack([xx.ravel(), yy.ravel(), np.zeros(xx.size, dtype=np.int64)], -1))
    total = parts[0].shape[0]
    while total < n_target * 1.3:
        if rng.random() < 0.75:
            size = rng.integers(4, 28, size=3)
            lo = np.array([rng.integers(0, ex[0] - size[0]), rng.integers(0, ex[1] - size[1]),
                           rng.integers(0, max(1, ex[2] - size[2]) // 2 + 1)])
            parts.append(_box_shell(rng, lo, lo + size))
        else:
            r = rng.integers(2, 7)
            base = np.array([rng.integers(r, ex[0] - r), rng.integers(r, ex[1] - r), rng.integers(0, 20)])
            parts.append(_tube(rng, base, r, rng.integers(8, 35)))
        total += parts[-1].shape[0]
    vox = np.unique(np.concatenate(parts, 0), axis=0)
    vox = vox[(vox >= 0).all(1) & (vox < ex).all(1)]
    return vox


def _model_grid(coord, batch_ids):
    """the model's own quantisation, in torch fp32 exactly as model.py:96-98."""
    c = torch.from_numpy(coord)
    g = torch.div(c - c.min(0)[0], VOXEL, rounding_mode="trunc").int().numpy().astype(np.int64)
    return np.concatenate([batch_ids[:, None], g], 1)


def make_batch(seed=2024, batch_size=32, max_points=4096, min_raw=1500, max_raw=6000, sizes=None,
               txt_len_range=(8, 30), pos_bins=15, pos_bin_size=0.01, with_targets=True, in_channels=7):
    """Config-2 style batch.  `sizes` (list[int]) pins the per-cloud point counts."""
    rng = np.random.default_rng(seed)
    clouds = []
    B = len(sizes) if sizes is not None else batch_size
    for i in range(B):
        if sizes is not None:
            n = int(sizes[i])
            raw = max(n + 64, int(n * 1.2))
        else:
            raw = int(rng.integers(min_raw, max_raw + 1))
            n = int(np.floor(rng.uniform(0.95, 1.0) * min(raw, max_points)))
        vox = surface_cloud_voxels(rng, raw)
        sel = rng.permutation(vox.shape[0])[: max(n + 32, int(n * 1.05))]
        # a point somewhere inside its voxel (augmented data is not on voxel centres)
        xyz = (vox[sel] + rng.uniform(0.2, 0.8, size=(sel.size, 3))) * VOXEL
        xyz = (xyz - xyz.mean(0, keepdims=True)).astype(np.float32)        # xyz_shift 'center'
        clouds.append((xyz, n))
    # de-duplicate under the model's quantisation (global min over the packed batch); the points that define
    # the global min corner are always kept so the quantisation does not move between iterations
    for _ in range(50):
        coord = np.concatenate([c for c, _ in clouds], 0)
        bid = np.concatenate([np.full(c.shape[0], i, dtype=np.int64) for i, (c, _) in enumerate(clouds)])
        key = _model_grid(coord, bid)
        gmin = np.unique(np.argmin(coord, 0))
        prio = np.concatenate([gmin, np.setdiff1d(np.arange(key.shape[0]), gmin)])
        _, first = np.unique(key[prio], axis=0, return_index=True)
        if first.size == key.shape[0]:
            break
        keep = np.zeros(key.shape[0], dtype=bool)
        keep[prio[first]] = True
        off = 0
        new = []
        for c, n in clouds:
            k = keep[off: off + c.shape[0]]
            off += c.shape[0]
            new.append((c[k], n))
        clouds = new
    # trim to the requested counts (keeping the global-min points, which are kept by position order anyway)
    coord = np.concatenate([c for c, _ in clouds], 0)
    gmin_pts = set(np.argmin(coord, 0).tolist())
    off = 0
    out = []
    for c, n in clouds:
        n = min(n, c.shape[0])
        idx = np.arange(c.shape[0])
        must = [j for j in idx if (off + j) in gmin_pts]
        rest = [j for j in idx if (off + j) not in gmin_pts]
        sel = np.array(sorted(must + rest[: n - len(must)]), dtype=np.int64)
        off += c.shape[0]
        out.append(c[sel])
    coord = np.concatenate(out, 0)
    bid = np.concatenate([np.full(c.shape[0], i, dtype=np.int64) for i, c in enumerate(out)])
    key = _model_grid(coord, bid)
    assert np.unique(key, axis=0).shape[0] == key.shape[0], "synthetic batch has duplicate voxels"

    npoints = [c.shape[0] for c in out]
    fts = []
    for c in out:
        rgb = rng.uniform(-1, 1, size=(c.shape[0], 3)).astype(np.float32)
        height = (c[:, 2:3] - c[:, 2].min()).astype(np.float32)
        f = np.concatenate([c, rgb, height], 1)
        fts.append(f[:, :in_channels] if in_channels <= 7 else f)
    txt_lens = [int(x) for x in rng.integers(txt_len_range[0], txt_len_range[1] + 1, size=B)]
    batch = {
        "pc_fts": torch.from_numpy(np.concatenate(fts, 0)),
        "npoints_in_batch": npoints,
        "offset": torch.from_numpy(np.cumsum(npoints)).long(),
        "txt_embeds": torch.from_numpy(rng.standard_normal((sum(txt_lens), 512)).astype(np.float32)),
        "txt_lens": txt_lens,
        "ee_poses": torch.from_numpy(np.concatenate([
            rng.normal(0, 0.1, size=(B, 3)), np.tile([0, 0, 0, 1.0], (B, 1)),
            rng.integers(0, 2, size=(B, 1))], 1).astype(np.float32)),
        "step_ids": torch.from_numpy(rng.integers(0, 10, size=B)).long(),
    }
    if with_targets:
        gt = np.zeros((B, 7), dtype=np.float32)
        probs = []
        for i, c in enumerate(out):
            # ground-truth gripper position near a surface point (so the soft label is non-degenerate)
            gt[i, :3] = c[rng.integers(0, c.shape[0])] + rng.normal(0, 0.004, size=3)
            gt[i, 3:6] = rng.integers(0, 72, size=3)
            gt[i, 6] = rng.integers(0, 2)
            probs.append(torch.from_numpy(
                disc_gt_pos_prob(c, gt[i, :3], pos_bin_size=pos_bin_size, pos_bins=pos_bins).astype(np.float32)))
        batch["gt_actions"] = torch.from_numpy(gt)
        batch["disc_pos_probs"] = probs
    return batch


def randomize_poses(batch, seed=0):
    """random unit quaternions (xyzw) for `ee_poses` (make_batch keeps the identity rotation)"""
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((batch["ee_poses"].shape[0], 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    batch["ee_poses"][:, 3:7] = torch.from_numpy(q)
    return batch


def make_planner_batch(seed=2024, batch_size=32, max_traj_len=5, sizes=None, pos_bins=15, pos_bin_size=0.01,
                       txt_len_range=(3, 8), **kw):
    """Config-3 style batch: pc_fts (N,4) = xyz + height, pc_labels, trajectories of up to
    `max_traj_len` steps with stop flags / masks and one soft position label per step."""
    b = make_batch(seed=seed, batch_size=batch_size, sizes=sizes, txt_len_range=txt_len_range, pos_bins=pos_bins,
                   pos_bin_size=pos_bin_size, with_targets=False, **kw)
    rng = np.random.default_rng(seed + 7919)
    f = b["pc_fts"]
    b["pc_fts"] = torch.cat([f[:, :3], f[:, 6:7]], 1).contiguous()
    N, B, T = f.shape[0], len(b["npoints_in_batch"]), max_traj_len
    b["pc_labels"] = torch.from_numpy(rng.choice(4, size=N, p=[0.7, 0.15, 0.1, 0.05])).long()
    randomize_poses(b, seed + 1)
    lens = rng.integers(1, T + 1, size=B)
    gt = np.zeros((B, T, 7), dtype=np.float32)
    stop = np.zeros((B, T), dtype=bool)
    mask = np.zeros((B, T), dtype=bool)
    probs = []
    off = 0
    for i, n in enumerate(b["npoints_in_batch"]):
        c = f[off: off + n, :3].numpy()
        off += n
        pr = np.zeros((T, 3, n * 2 * pos_bins), dtype=np.float32)
        for t in range(T):
            gt[i, t, :3] = c[rng.integers(0, n)] + rng.normal(0, 0.004, size=3)
            gt[i, t, 3:6] = rng.integers(0, 72, size=3)
            gt[i, t, 6] = rng.integers(0, 2)
            pr[t] = disc_gt_pos_prob(c, gt[i, t, :3], pos_bin_size=pos_bin_size, pos_bins=pos_bins)
        mask[i, : lens[i]] = True
        stop[i, lens[i] - 1] = True
        probs.append(torch.from_numpy(pr))
    b.update(gt_trajs=torch.from_numpy(gt), gt_trajs_stop=torch.from_numpy(stop), traj_masks=torch.from_numpy(mask),
             traj_lens=[int(v) for v in lens], gt_trajs_disc_pos_probs=probs)
    return b


def batch_nbytes(batch):
    n = 0
    for v in batch.values():
        if isinstance(v, torch.Tensor):
            n += v.numel() * v.element_size()
        elif isinstance(v, list) and v and isinstance(v[0], torch.Tensor):
            n += sum(t.numel() * t.element_size() for t in v)
    return n
