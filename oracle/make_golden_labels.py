"""Mint golden vectors for the soft position labels from the UNMODIFIED reference function (test infrastructure; run in the
build container):      python oracle/make_golden_labels.py      # writes tests/golden/labels.pt

reference: genrobo3d/utils/action_position_utils.py:7-46 (`get_disc_gt_pos_prob`), both heatmap types, including an axis whose
ground truth is farther than 1 cm from every candidate (the `argmin` one-hot fallback) and an exact hit (distance 0).
The device-side producer `lotus_disc_pos_prob` (csrc/batch_producer.cu) and the numpy restatement
`synthetic.disc_gt_pos_prob` are checked against these.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402


def main():
    import_reference()
    from genrobo3d.utils.action_position_utils import get_disc_gt_pos_prob
    rng = np.random.default_rng(11)
    cases = []
    for n in (257, 1, 64):
        xyz = rng.uniform(-0.3, 0.3, size=(n, 3)).astype(np.float32)
        gt = (xyz[rng.integers(0, n)] + rng.normal(0, 0.003, 3)).astype(np.float32)
        cases.append((xyz, gt))
    xyz = rng.uniform(-0.3, 0.3, size=(100, 3)).astype(np.float32)
    gt = xyz[5].copy()
    gt[1] = 5.0                                   # axis 1: no candidate within 1 cm -> one-hot at the first closest candidate
    cases.append((xyz, gt))                       # axes 0 and 2: exact hit on point 5 (distance 0 -> weight 1 / 1e-4)
    out = []
    for xyz, gt in cases:
        rec = {"xyz": torch.from_numpy(xyz), "gt": torch.from_numpy(gt)}
        for kind in ("dist", "plain"):
            rec[kind] = torch.from_numpy(np.asarray(get_disc_gt_pos_prob(xyz, gt, pos_bin_size=0.01, pos_bins=15, heatmap_type=kind)))
        out.append(rec)
    path = os.path.join(ROOT, "tests", "golden", "labels.pt")
    torch.save(out, path)
    print("labels: %d cases, %.3f MB" % (len(out), os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    main()
