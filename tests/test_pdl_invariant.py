"""Programmatic dependent launch invariant (CPU test, reads the built library's SASS).

A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor on the stream is still running;
it is only correct if it executes `griddepcontrol.wait` (SASS: ACQBULK) before touching global memory.  Every kernel that is launched
through `launch_pdl(...)` (common.cuh) or with the attribute directly (linear_tc.cu) must therefore contain that instruction, and the
wait must come before the first `griddepcontrol.launch_dependents` (SASS: PREEXIT)."""
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "robot-3dlotus_b200", "liblotus.so")
CSRC = os.path.join(ROOT, "robot-3dlotus_b200", "csrc")


def _pdl_kernel_names():
    names = set()
    for path in glob.glob(os.path.join(CSRC, "*.cu")):
        src = open(path).read()
        for m in re.finditer(r"launch_pdl\(\s*([A-Za-z_][A-Za-z0-9_:]*)", src):
            names.add(m.group(1).split("::")[-1])
        if "cudaLaunchAttributeProgrammaticStreamSerialization" in src:
            for m in re.finditer(r"cudaLaunchKernelEx\(&cfg,\s*([A-Za-z_][A-Za-z0-9_:]*)", src):
                names.add(m.group(1).split("::")[-1])
    names.discard("kern")            # the parameter name inside launch_pdl itself
    return names


@pytest.mark.skipif(not os.path.exists(SO) or shutil.which("cuobjdump") is None, reason="needs the built library and cuobjdump")
def test_every_pdl_launched_kernel_waits_before_global_memory():
    names = _pdl_kernel_names()
    assert {"k_lin", "k_gemm_tc", "k_pattn_fwd", "k_pattn_bwd", "k_ln_fwd_vec", "k_ln_bwd_vec"} <= names, names
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        if cur is not None:
            mm = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if mm:
                funcs[cur].append(mm.group(1))
    checked = 0
    for mangled, ops in funcs.items():
        base = [n for n in names if re.search(r"\d+%s(I|E|P|v|$)" % re.escape(n), mangled)]
        if not base:
            # a kernel that triggers its dependents early without being PDL-launched itself would be harmless, but none should exist
            assert "PREEXIT" not in ops or "ACQBULK" in ops, mangled
            continue
        checked += 1
        assert "ACQBULK" in ops, f"{mangled}: launched with the PDL attribute but never executes griddepcontrol.wait"
        if "PREEXIT" in ops:
            assert ops.index("ACQBULK") < ops.index("PREEXIT"), f"{mangled}: launch_dependents before wait"
    assert checked >= 12, checked        # template instantiations of the kernels above
