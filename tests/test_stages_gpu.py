"""GPU parity for potrf / trsm / hegst / ormtr against the oracle (LAPACK + numpy restatement)."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import lapack, matgen, metrics, restatement as R

pytestmark = pytest.mark.gpu


def _rand(shape, cplx, rng):
    a = rng.standard_normal(shape)
    if cplx:
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n", [1, 5, 64, 65, 200, 777])
def test_potrf(cplx, n):
    from eigensolver_gpu_b200 import stages as S
    _, b = matgen.family_c(n, cplx, seed=n)
    junk = np.tril(np.ones((n, n)), -1) * 7.0
    bd = S.to_dev(np.triu(b) + junk)
    info = S.potrf(bd)
    assert info == 0
    got = np.array(S.to_host(bd))
    u = lapack.potrf(b)
    assert np.abs(np.triu(got) - u).max() <= 50 * n * metrics.EPS * np.abs(u).max()
    assert np.array_equal(np.tril(got, -1), junk)


def test_potrf_reports_not_positive_definite():
    from eigensolver_gpu_b200 import stages as S
    n = 100
    _, b = matgen.family_c(n, False, seed=1)
    b[70, 70] = -5.0
    assert S.potrf(S.to_dev(b)) != 0


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("side,trans,m,n", [("L", "N", 130, 70), ("L", "C", 200, 200), ("R", "N", 90, 257),
                                            ("L", "N", 64, 1), ("R", "N", 1, 64), ("L", "N", 500, 333)])
def test_trsm(cplx, side, trans, m, n):
    from eigensolver_gpu_b200 import stages as S
    rng = np.random.default_rng(m + n)
    nu = m if side == "L" else n
    u = np.triu(_rand((nu, nu), cplx, rng)) + 4 * np.eye(nu)
    b = _rand((m, n), cplx, rng)
    bd = S.to_dev(b)
    S.trsm(side, trans, S.to_dev(u + np.tril(np.ones((nu, nu)), -1) * 1e20), bd)
    got = S.to_host(bd)
    if side == "L":
        ref = sla.solve_triangular(u, b, trans="C" if trans == "C" else "N", lower=False)
    else:
        ref = sla.solve_triangular(u, b.conj().T, trans="C", lower=False).conj().T
    assert np.abs(got - ref).max() <= 100 * nu * metrics.EPS * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n", [3, 64, 150, 500])
def test_hegst(cplx, n):
    from eigensolver_gpu_b200 import stages as S
    a, b = matgen.family_c(n, cplx, seed=3 * n)
    u = lapack.potrf(b)
    ad = S.to_dev(np.triu(a))
    S.hegst(ad, S.to_dev(u))
    got = np.array(S.to_host(ad))
    ref = R.hegst_reference(a, u, 448)                 # the reference's blocked variant (restatement)
    ref2 = np.triu(lapack.hegst(a, u))                 # LAPACK ?hegst
    scale = np.abs(ref2).max()
    assert np.abs(np.triu(got) - ref2).max() <= 100 * n * metrics.EPS * scale
    assert np.abs(np.triu(got) - ref).max() <= 100 * n * metrics.EPS * scale


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n,m,nb", [(2, 2, 128), (65, 10, 64), (300, 300, 128), (513, 100, 128), (700, 64, 256)])
def test_ormtr(cplx, n, m, nb):
    from eigensolver_gpu_b200 import stages as S
    from eigensolver_gpu_b200._lib import load
    load().eigb200_set_option(b"bt_nb", nb)
    a, _ = matgen.family_c(n, cplx, seed=n + 1)
    a2, d, e, tau = lapack.hetrd(a)
    rng = np.random.default_rng(0)
    z = _rand((n, m), cplx, rng)
    zd = S.to_dev(z)
    S.ormtr(S.to_dev(a2), S.to_dev(tau if n > 1 else np.zeros(1, dtype=a2.dtype)), zd, m=m)
    load().eigb200_set_option(b"bt_nb", 128)
    got = S.to_host(zd)
    ref = lapack.ormtr("L", "U", "N", a2, tau, z) if n > 1 else z
    assert np.abs(got - ref).max() <= 50 * n * metrics.EPS * np.abs(z).max()


@pytest.mark.parametrize("cplx", [False, True])
def test_trsm_leaf_sizes_agree(cplx):
    """solves with a finished factor: 256x256 inverted leaves (default) and 64x64 leaves give the same solution"""
    import numpy as np
    from eigensolver_gpu_b200 import stages as S
    from eigensolver_gpu_b200._lib import load
    from oracle import matgen, metrics
    lib = load()
    n, m = 900, 333                               # partial last 256-block, partial last 64-block
    _, b = matgen.family_c(n, cplx, seed=4)
    rng = np.random.default_rng(2)
    x = rng.standard_normal((n, m)) + (1j * rng.standard_normal((n, m)) if cplx else 0)
    bd = S.to_dev(np.triu(b))
    assert S.potrf(bd) == 0
    u = np.triu(np.array(S.to_host(bd)))
    sols = []
    for leaf in (1, 0):
        assert lib.eigb200_set_option(b"trsm_leaf256", leaf) == 0
        xd = S.to_dev(x)
        S.trsm("L", "N", bd, xd, m=n, n=m)
        sols.append(np.array(S.to_host(xd)))
    lib.eigb200_set_option(b"trsm_leaf256", 1)
    for sol in sols:
        r = u @ sol - x
        assert np.abs(r).max() <= 50 * n * metrics.EPS * np.abs(u).max() * np.abs(sol).max()
