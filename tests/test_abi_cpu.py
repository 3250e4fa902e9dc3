"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, exports every
symbol include/omok_b200.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_declared_symbols(omk):
    omk.build()
    L = omk.load_library()
    names = omk.declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", omk.lib_path()], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported
    assert all(s.startswith("omk_") for s in exported if not s.startswith("_")), exported


def test_library_targets_sm_100a(omk):
    out = subprocess.run(["cuobjdump", "-lelf", omk.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(omk):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(omk.OmkError) as e:
        omk.Context(device=0, capacity_envs=4, capacity_trees=1)
    assert e.value.code == -2


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "omok-ai_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} refers to the oracle"
