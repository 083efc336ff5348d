"""Mint batch-producer golden vectors with the reference's own numpy helper (test infrastructure; build container only).

    python oracle/make_golden_producer.py     # writes tests/golden/producer.pt

The reference's dataset class needs LMDB / msgpack_numpy / open3d (absent offline), so the per-keystep transformation of
`SimplePolicyDataset.__getitem__` (train/datasets/simple_policy_dataset.py:281-346) is replayed here statement by statement around the
reference's `random_rotate_z` (train/datasets/common.py:81-87, imported from /root/reference), with numpy's global generator seeded
like a worker would use it.  Stored: the raw cloud, every random draw, and the resulting features / labels.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TABLE_HEIGHT = 0.7505          # genrobo3d/configs/rlbench/constants.py get_robot_workspace: TABLE_HEIGHT


def main():
    import_reference()
    import importlib.util
    # train/datasets/__init__ pulls lmdb etc.: load common.py directly by path
    spec = importlib.util.spec_from_file_location("ref_common", "/root/reference/genrobo3d/train/datasets/common.py")
    common = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(common)
    random_rotate_z = common.random_rotate_z
    rng = np.random.default_rng(3)
    cases = []
    for case, (m, num_points, use_color) in enumerate([(5000, 4096, True), (3000, 4096, True), (1500, 4096, False), (130, 4096, True)]):
        vox = rng.permutation(90 * 90 * 40)[:m]
        xyz = (np.stack([vox // (90 * 40), (vox // 40) % 90, vox % 40], 1).astype(np.float32) * 0.01 + np.array([-0.3, -0.45, 0.76], dtype=np.float32)).astype(np.float32)
        rgb = rng.integers(0, 256, size=(m, 3)).astype(np.uint8)
        gt_action = np.concatenate([xyz[rng.integers(0, m)] + rng.normal(0, 0.004, 3), [0, 0, 0, 1, 1.0]]).astype(np.float32)
        np.random.seed(100 + case)
        # ---- simple_policy_dataset.py:281-294
        if len(xyz) > num_points:
            point_idxs = np.random.choice(len(xyz), num_points, replace=False)
        else:
            max_npoints = int(len(xyz) * np.random.uniform(0.95, 1))
            point_idxs = np.random.permutation(len(xyz))[:max_npoints]
        x = xyz[point_idxs]
        c = rgb[point_idxs]
        height = x[:, -1] - TABLE_HEIGHT                                        # :296
        # ---- _augment_pc :158-181 (aug_max_rot 45 degrees, configured in radians by the dataset: np.deg2rad)
        aug_max_rot = np.deg2rad(45)
        angle = np.random.uniform(-1, 1) * aug_max_rot
        x = random_rotate_z(x, angle=angle)
        gt_pos = random_rotate_z(gt_action[:3], angle=angle)
        pc_noises = np.random.uniform(0, 0.002, size=x.shape)
        x = pc_noises + x
        # ---- :321-346 (xyz_shift center, xyz_norm False)
        centroid = np.mean(x, 0)
        x = (x - centroid) / 1
        gt_pos = (gt_pos - centroid) / 1
        cc = (c / 255.) * 2 - 1
        pc_ft = np.concatenate([x, cc], 1) if use_color else x
        pc_ft = np.concatenate([pc_ft, height[:, None]], 1)
        cases.append({"raw_xyz": torch.from_numpy(xyz), "raw_rgb": torch.from_numpy(rgb), "use_color": use_color, "gt_pos_in": torch.from_numpy(gt_action[:3].astype(np.float64)),
                      "point_idxs": torch.from_numpy(point_idxs.astype(np.int64)), "angle": float(angle), "noise": torch.from_numpy(pc_noises),
                      "table_height": TABLE_HEIGHT, "pc_ft": torch.from_numpy(pc_ft).float(), "centroid": torch.from_numpy(centroid),
                      "gt_pos_out": torch.from_numpy(gt_pos).float()})
    torch.save(cases, os.path.join(GOLD, "producer.pt"))
    print("producer.pt:", [(tuple(c["raw_xyz"].shape), tuple(c["pc_ft"].shape)) for c in cases], "MB", os.path.getsize(os.path.join(GOLD, "producer.pt")) / 1e6)


if __name__ == "__main__":
    main()
