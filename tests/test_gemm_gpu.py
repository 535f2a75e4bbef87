"""GPU parity: the DMMA GEMM family against numpy on the same seeded inputs (through the C ABI)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand(shape, cplx, rng):
    a = rng.standard_normal(shape)
    if cplx:
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a)


def _op(a, t):
    return a if t == "N" else (a.T if t == "T" else a.conj().T)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("ta,tb", [("N", "N"), ("N", "C"), ("C", "N"), ("T", "T"), ("C", "C")])
@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (257, 131, 77), (64, 300, 1), (5, 7, 3), (513, 384, 200)])
def test_gemm(cplx, ta, tb, m, n, k):
    from eigensolver_gpu_b200 import stages as S
    rng = np.random.default_rng(m * 1000 + n + k)
    a = _rand((m, k) if ta == "N" else (k, m), cplx, rng)
    b = _rand((k, n) if tb == "N" else (n, k), cplx, rng)
    c = _rand((m, n), cplx, rng)
    ref = 0.75 * _op(a, ta) @ _op(b, tb) - 0.5 * c
    cd = S.to_dev(c)
    S.gemm(ta, tb, 0.75, S.to_dev(a), S.to_dev(b), -0.5, cd)
    got = S.to_host(cd)
    tol = 1e-13 * (k + 4) * max(1.0, np.abs(ref).max())
    assert np.abs(got - ref).max() < tol


@pytest.mark.parametrize("cplx", [False, True])
def test_gemm_unaligned_submatrix(cplx):
    """odd leading dimension and odd offsets force the 8-byte cp.async path for real data"""
    from eigensolver_gpu_b200 import stages as S
    import torch
    rng = np.random.default_rng(5)
    big_a = _rand((301, 90), cplx, rng)
    big_b = _rand((301, 140), cplx, rng)
    a = big_a[3:203, 1:80]      # 200 x 79, ld 301
    b = big_b[5:84, 7:130]      # 79 x 123
    ref = a @ b
    ad, bd = S.to_dev(big_a), S.to_dev(big_b)
    cd = torch.zeros((123, 200), dtype=ad.dtype, device="cuda")
    av = ad[1:80, 3:]
    bv = bd[7:130, 5:]
    from eigensolver_gpu_b200._lib import load, check
    import ctypes as C
    lib = load()
    fn = lib.eigb200_zgemm if cplx else lib.eigb200_dgemm
    check(fn(b"N", b"N", 200, 123, 79, 1.0, C.c_void_p(av.data_ptr()), 301, C.c_void_p(bv.data_ptr()), 301, 0.0,
             C.c_void_p(cd.data_ptr()), 200), "gemm")
    got = S.to_host(cd)
    assert np.abs(got - ref).max() < 1e-11


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n,k", [(128, 32), (300, 64), (1000, 37)])
def test_her2k_upper(cplx, n, k):
    from eigensolver_gpu_b200 import stages as S
    rng = np.random.default_rng(n + k)
    a = _rand((n, k), cplx, rng)
    b = _rand((n, k), cplx, rng)
    c = _rand((n, n), cplx, rng)
    c = c + c.conj().T
    ref = c - (a @ b.conj().T + b @ a.conj().T)
    cd = S.to_dev(c)
    S.her2k(-1.0, S.to_dev(a), S.to_dev(b), 1.0, cd)
    got = S.to_host(cd)
    assert np.abs(np.triu(got) - np.triu(ref)).max() < 1e-12 * k * max(1, np.abs(ref).max())
    # strict lower part untouched
    assert np.array_equal(np.tril(got, -1), np.tril(c, -1))
    if cplx:
        assert np.all(np.diag(got).imag == 0)
