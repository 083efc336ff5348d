"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol declared in
include/lotus.h (no compute calls without a GPU), the header is in sync with the sources, and the product
path refuses to run without CUDA (no silent fallback)."""
import ctypes
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, pkg


def test_header_in_sync_with_sources():
    before = open(os.path.join(ROOT, "include", "lotus.h")).read()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_header.py")], stdout=subprocess.DEVNULL)
    assert open(os.path.join(ROOT, "include", "lotus.h")).read() == before, "include/lotus.h is stale: run tools/gen_header.py"


def test_library_exports_every_declared_symbol():
    lib = pkg("lib")
    if not os.path.exists(lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    protos = lib.parse_header()
    assert len(protos) >= 38
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), f"{name} declared in include/lotus.h but not exported by liblotus.so"
    L = lib.load()
    assert L.cdll.lotus_version() >= 100


def test_every_entry_point_cites_the_reference():
    src = open(os.path.join(ROOT, "include", "lotus.h")).read()
    assert "(undocumented)" not in src


def test_no_cpu_fallback():
    """the product path must fail loudly without a CUDA device instead of falling back to CPU math"""
    cfgm, policy = pkg("config"), pkg("policy")
    cfg = cfgm.small_model_config()
    model = policy.build_parameter_shell(cfg)
    syn = pkg("synthetic")
    batch = syn.make_batch(seed=0, sizes=[50, 60])
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        model(batch, compute_loss=True)


def test_product_never_imports_oracle():
    pdir = os.path.join(ROOT, "robot-3dlotus_b200")
    for fn in os.listdir(pdir):
        if fn.endswith(".py"):
            src = open(os.path.join(pdir, fn)).read()
            assert "oracle" not in src.replace("oracle/", "").replace("the oracle", "").replace("CPU oracle", "") or fn == "testing.py", fn
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_conv_split_rule_is_host_logic():
    """lotus_subm_conv_splits is pure host arithmetic (no launch): only grids under half a wave of 128-row tiles get a split reduction
    (the 976-row stage of the production step), never the wide stages"""
    lib = pkg("lib")
    cdll = ctypes.CDLL(lib.LIB_PATH)
    f = cdll.lotus_subm_conv_splits
    f.restype = ctypes.c_int
    assert f(976, 768, 27 * 768) > 1 and f(976, 768, 27 * 768) <= 8
    assert f(3210, 512, 27 * 512) == 1 and f(12140, 256, 27 * 256) == 1
    assert f(45262, 128, 27 * 128) == 1 and f(96652, 64, 27 * 64) == 1
    assert f(500, 64, 64) == 1                                     # a short reduction is never split
