"""GPU parity: symv/hemv and the blocked tridiagonalization against LAPACK (oracle) through the C ABI."""
import numpy as np
import pytest

from oracle import lapack, matgen, metrics

pytestmark = pytest.mark.gpu


def _herm(n, cplx, seed):
    a, _ = matgen.family_c(n, cplx, seed)
    return a


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 200, 1000, 2049])
def test_hemv_matches_oracle(cplx, n):
    from eigensolver_gpu_b200 import stages as S
    a = _herm(n, cplx, 100 + n)
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
    junk = np.tril(rng.standard_normal((n, n)), -1)          # the lower triangle must never be read
    ad = S.to_dev(np.triu(a) + junk * 1e30)
    y = S.to_host(S.hemv(ad, S.to_dev(x)))
    ref = lapack.hemv_upper(a, x)
    assert np.abs(y - ref).max() <= 4 * n * metrics.EPS * np.abs(a).max() * np.abs(x).max() + 1e-300


@pytest.mark.parametrize("coop", [1, 0])
@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n,nb", [(1, 64), (2, 64), (3, 64), (33, 32), (64, 64), (65, 64), (130, 32), (257, 64),
                                  (700, 64), (1500, 48)])
def test_hetrd_matches_lapack(cplx, n, nb, coop):
    from eigensolver_gpu_b200 import stages as S
    from eigensolver_gpu_b200._lib import load
    lib = load()
    assert lib.eigb200_set_option(b"trd_nb", nb) == 0
    assert lib.eigb200_set_option(b"trd_coop", coop) == 0
    a = _herm(n, cplx, 7 + n)
    rng = np.random.default_rng(n)
    junk = np.tril(rng.standard_normal((n, n)), -1)
    ad = S.to_dev(np.triu(a) + junk)
    d, e, tau = S.hetrd(ad)
    d, e, tau = S.to_host(d), S.to_host(e), S.to_host(tau)
    aout = np.array(S.to_host(ad))
    lib.eigb200_set_option(b"trd_nb", 64)
    lib.eigb200_set_option(b"trd_coop", 1)
    # strict lower triangle untouched
    assert np.array_equal(np.tril(aout, -1), junk)
    _, dl, el, taul = lapack.hetrd(a)
    an = np.abs(a).sum(axis=0).max()
    tol = 20 * n * metrics.EPS * an
    assert np.abs(d - dl).max() <= tol
    if n > 1:
        assert np.abs(e - el).max() <= tol
        assert np.abs(tau - taul).max() <= 200 * n * metrics.EPS
    # similarity check through the oracle's own stedc + ormtr: A Z = Z diag(w)
    if n > 1:
        w, zt, info = lapack.stedc(d, e)
        assert info == 0
        z = lapack.ormtr("L", "U", "N", np.asfortranarray(aout), tau, zt.astype(aout.dtype))
        g = metrics.std_gates(a, w, z)
        assert g["residual_max"] < 30 and g["orth"] < 30


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("opt,val", [("trd_l2keep_mb", 0), ("trd_l2keep_mb", 1), ("trd_upc", 6), ("trd_upc", 1 + (2 << 8)),
                                     ("trd_prefetch", 4)])
def test_hetrd_tuning_options_do_not_change_the_result_class(cplx, opt, val):
    """scheduling / cache-policy knobs of the tile engine: same (d, e) as LAPACK to rounding whatever their value"""
    from eigensolver_gpu_b200 import stages as S
    from eigensolver_gpu_b200._lib import load
    lib = load()
    n = 1700
    old = lib.eigb200_get_option(opt.encode())
    assert lib.eigb200_set_option(opt.encode(), val) == 0
    try:
        a = _herm(n, cplx, 11)
        ad = S.to_dev(np.triu(a))
        d, e, tau = S.hetrd(ad)
        d, e = S.to_host(d), S.to_host(e)
    finally:
        lib.eigb200_set_option(opt.encode(), old)
    _, dl, el, _ = lapack.hetrd(a)
    tol = 20 * n * metrics.EPS * np.abs(a).sum(axis=0).max()
    assert np.abs(d - dl).max() <= tol and np.abs(e - el).max() <= tol


def test_hetrd_is_run_to_run_deterministic():
    """the dynamic tile queue decides WHO runs a unit, never WHAT is summed in which order: bitwise identical repeats"""
    from eigensolver_gpu_b200 import stages as S
    a = _herm(2300, True, 5)
    outs = []
    for _ in range(3):
        ad = S.to_dev(np.triu(a))
        d, e, tau = S.hetrd(ad)
        outs.append((S.to_host(d).copy(), S.to_host(e).copy(), S.to_host(tau).copy()))
    for o in outs[1:]:
        assert all(np.array_equal(x, y) for x, y in zip(outs[0], o))
