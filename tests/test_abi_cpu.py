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


def test_rust_sys_crate_declares_every_entry_point_with_the_same_arity(omk):
    """shims/omok-b200-sys cannot be compiled here (no Rust toolchain): at least keep its `extern "C"` block in step with
    the header -- every declared symbol present, same number of parameters."""
    import re

    header = open(os.path.join(ROOT, "include", "omok_b200.h")).read()
    rust = open(os.path.join(ROOT, "shims", "omok-b200-sys", "src", "lib.rs")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    for name in omk.declared_symbols():
        c = re.search(r"\b" + name + r"\s*\(([^)]*)\)", header)
        r = re.search(r"pub fn " + name + r"\s*\(([^)]*)\)", rust)
        assert c and r, name
        c_args = [a for a in c.group(1).split(",") if a.strip() and a.strip() != "void"]
        r_args = [a for a in r.group(1).split(",") if a.strip()]
        assert len(c_args) == len(r_args), (name, c_args, r_args)
        # ... and the same parameter TYPES, position by position
        for ca, ra in zip(c_args, r_args):
            assert _c_type_as_rust(ca) == ra.split(":", 1)[1].strip(), (name, ca, ra)


_SCALARS = {"int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64", "uint8_t": "u8", "int8_t": "i8",
            "uint16_t": "u16", "float": "f32", "void": "c_void", "omk_ctx": "omk_ctx", "omk_selfplay_config": "omk_selfplay_config",
            "omk_selfplay_stats": "omk_selfplay_stats"}


def _c_type_as_rust(decl: str) -> str:
    """`const float *const *tensors` -> `*const *const f32` (the parameter name is dropped)."""
    import re

    toks = re.findall(r"[A-Za-z_]\w*|\*", decl)
    toks = toks[:-1]  # parameter name
    base = [t for t in toks if t in _SCALARS]
    assert len(base) == 1, decl
    # walk the declarator left to right: qualifiers before a `*` say whether THAT pointer's pointee is const
    out, const_pending, seen_base = _SCALARS[base[0]], False, False
    for t in toks:
        if t == "const":
            const_pending = True
        elif t in _SCALARS:
            seen_base = True
        elif t == "*":
            assert seen_base, decl
            out = ("*const " if const_pending else "*mut ") + out
            const_pending = False
    return out
