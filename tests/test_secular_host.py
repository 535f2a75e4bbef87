"""CPU test of the host/device secular-equation core (csrc/secular.cuh) against LAPACK dlaed4 (oracle) and
against the Gu-Eisenstat orthogonality property it must support."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import lapack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sec():
    out = os.path.join(ROOT, "tests", "csrc", "_secular_host.so")
    src = os.path.join(ROOT, "tests", "csrc", "secular_host.cpp")
    hdr = os.path.join(ROOT, "eigensolver_gpu_b200", "csrc")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-I", hdr, src, "-o", out])
    lib = C.CDLL(out)
    lib.secular_all.restype = C.c_int
    return lib


def _solve(lib, d, z, rho):
    k = len(d)
    lam = np.zeros(k)
    delta = np.zeros((k, k), order="F")
    iters = np.zeros(k, dtype=np.int32)
    mx = lib.secular_all(k, d.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), C.c_double(rho),
                         lam.ctypes.data_as(C.c_void_p), delta.ctypes.data_as(C.c_void_p),
                         iters.ctypes.data_as(C.c_void_p))
    return lam, delta, iters, mx


def _cases():
    rng = np.random.default_rng(0)
    out = []
    for k in (1, 2, 3, 10, 64, 300):
        d = np.sort(rng.standard_normal(k))
        z = rng.standard_normal(k)
        z /= np.linalg.norm(z)
        out.append((f"rand{k}", d, z, abs(rng.standard_normal()) + 0.1))
    k = 100
    d = np.linspace(0, 1, k)
    z = np.ones(k) / np.sqrt(k)
    out.append(("uniform", d, z, 2.0))
    out.append(("tiny_rho", d, z, 1e-10))
    out.append(("huge_rho", d, z, 1e8))
    z2 = z.copy(); z2[::2] *= 1e-7; z2 /= np.linalg.norm(z2)
    out.append(("tiny_z", d, z2, 1.0))
    dc = np.sort(np.concatenate([1 + 1e-9 * np.arange(50), 2 + 1e-12 * np.arange(50)]))
    out.append(("clusters", dc, z, 0.5))
    return out


@pytest.mark.parametrize("name,d,z,rho", _cases(), ids=[c[0] for c in _cases()])
def test_secular_matches_dlaed4(sec, name, d, z, rho):
    k = len(d)
    lam, delta, iters, mx = _solve(sec, d, z, rho)
    assert mx <= 40, f"too many iterations: {mx}"
    for j in range(k):
        dl, lref, info = lapack.laed4(k, j + 1, d, z, rho)
        assert info == 0
        scale = max(abs(d).max(), rho)
        assert abs(lam[j] - lref) <= 8 * np.finfo(float).eps * scale, (j, lam[j], lref)
        # differences d_i - lambda_j agree to high relative accuracy (what the eigenvectors are built from)
        if k > 2:     # (dlaed4 returns other quantities in delta for k <= 2)
            rel = np.abs(delta[:, j] - dl) / np.maximum(np.abs(dl), 1e-300)
            assert rel.max() < 1e-6, (j, rel.max())
    # interlacing
    assert np.all(lam[:-1] > d[:-1]) and np.all(lam[:-1] < d[1:]) if k > 1 else True
    assert lam[-1] > d[-1]
    # Gu-Eisenstat: recomputed z-hat gives numerically orthogonal eigenvectors
    if k > 1:
        w = np.ones(k)
        for i in range(k):
            p = delta[i, i]
            for j in range(k):
                if j != i:
                    p *= delta[i, j] / (d[i] - d[j])
            w[i] = np.copysign(np.sqrt(-p), z[i])
        u = w[:, None] / delta
        u /= np.linalg.norm(u, axis=0)
        assert np.abs(u.T @ u - np.eye(k)).max() < 50 * k * np.finfo(float).eps
