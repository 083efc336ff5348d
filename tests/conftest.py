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
