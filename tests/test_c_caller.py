"""A compiled C program calling the C ABI (no Python, no ctypes in the process) with the reference test driver's buffer
and lwork formulas (test_driver/test_zhegvdx.F90:266-293, test_dsygvdx.F90:292-314): builds on CPU, runs on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "csrc", "c_caller.c")
EXE = os.path.join(ROOT, "tests", "csrc", "_c_caller")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build():
    from eigensolver_gpu_b200 import build
    lib = build.build()
    libdir = os.path.dirname(lib)
    cmd = ["gcc", "-O2", "-std=c99", "-o", EXE, SRC, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           "-L", libdir, "-leigb200", "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(CUDA, "lib64")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_c_caller_compiles_and_links_against_the_abi():
    exe = _build()
    assert os.path.exists(exe)
    out = subprocess.run([exe], capture_output=True, text=True)      # no arguments: usage, no CUDA call
    assert out.returncode == 2 and "usage" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,m", [("z", 300, 300), ("d", 512, 64), ("z", 129, 17)])
def test_c_caller_solves(kind, n, m):
    exe = _build()
    out = subprocess.run([exe, kind, str(n), str(m)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().splitlines()[-1].startswith("OK ")
