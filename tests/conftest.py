import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def pkg(name=None):
    return importlib.import_module("robot-3dlotus_b200" + ("." + name if name else ""))


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


@pytest.fixture(scope="session")
def units():
    return load_golden("units.pt")


@pytest.fixture(scope="session")
def gold_small():
    return load_golden("policy_small.pt")


@pytest.fixture(scope="session")
def gold_prod():
    return load_golden("policy_production.pt")


@pytest.fixture(scope="session")
def gold_variants():
    return load_golden("variants.pt")


@pytest.fixture(scope="session")
def gold_big():
    """BASELINE config-2-sized fixture (two keysteps of 4096 / 3000 points, production width): the batch is regenerated from
    (seed, sizes) and checked against the digest the generator stored (oracle/make_golden_big.py)."""
    g = load_golden("policy_production_big.pt")
    batch = pkg("synthetic").make_batch(seed=g["batch_seed"], sizes=g["batch_sizes"])
    d = g["batch_digest"]
    assert abs(batch["pc_fts"].double().sum().item() - d["pc_fts_sum"]) <= 1e-9 * max(1.0, abs(d["pc_fts_sum"]))
    assert abs(batch["pc_fts"].double().abs().sum().item() - d["pc_fts_abs_sum"]) <= 1e-9 * d["pc_fts_abs_sum"]
    assert abs(batch["txt_embeds"].double().sum().item() - d["txt_sum"]) <= 1e-9 * max(1.0, abs(d["txt_sum"]))
    assert abs(batch["gt_actions"].double().sum().item() - d["gt_sum"]) <= 1e-9 * max(1.0, abs(d["gt_sum"]))
    assert list(batch["txt_lens"]) == d["txt_lens"]
    g["batch"] = batch
    return g
