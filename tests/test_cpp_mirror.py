"""The C++ host mirror (include/omok_b200.hpp) compiles and links against libomok_b200.so (CPU),
and passes the reference-shaped tests on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_mirror")


def build(omk):
    omk.build()
    libdir = os.path.dirname(omk.lib_path())
    subprocess.check_call(
        ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"),
         "-o", EXE, "-L", libdir, "-l:libomok_b200.so", f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])


def test_cpp_mirror_compiles_and_links(omk):
    build(omk)
    out = subprocess.run([EXE], capture_output=True, text=True)
    assert out.returncode == 0 and "linked ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_mirror_runs_reference_tests(omk):
    build(omk)
    import importlib
    import tempfile

    ckpt = os.path.join(tempfile.mkdtemp(), "alpha-zero")
    out = subprocess.run([EXE, "--run", ckpt], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "cpp host mirror ok" in out.stdout, out.stdout + out.stderr
    # the checkpoint the C++ mirror wrote is the same bincode image the Python mirror reads and writes
    model_io = importlib.import_module("omok-ai_b200.model_io")
    names, params = model_io.load(ckpt)
    assert names == model_io.VARIABLE_NAMES and len(params) == 31
    assert model_io.dumps(params) == open(ckpt, "rb").read()
