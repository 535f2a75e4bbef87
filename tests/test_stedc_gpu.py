"""GPU parity: on-device tridiagonal divide & conquer against LAPACK dstedc (oracle)."""
import numpy as np
import pytest

from oracle import lapack, matgen, metrics

pytestmark = pytest.mark.gpu


def _run(d, e):
    from eigensolver_gpu_b200 import stages as S
    w, q = S.stedc(S.to_dev(d), S.to_dev(e) if len(e) else S.to_dev(np.zeros(1)))
    return S.to_host(w), np.array(S.to_host(q))


@pytest.mark.parametrize("kind", ["random", "toeplitz", "wilkinson", "glued", "graded", "zero_e", "identity"])
@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 64, 65, 100, 257, 1000, 2500])
def test_stedc_matches_lapack(kind, n):
    d, e = matgen.tridiag_family(n, kind, seed=n)
    w, q = _run(d, e)
    wr, zr, info = lapack.stedc(d, e)
    assert info == 0
    t = np.diag(d) + (np.diag(e, 1) + np.diag(e, -1) if n > 1 else 0)
    tn = max(np.abs(t).sum(axis=0).max(), 1e-300)
    assert np.all(np.isfinite(w)) and np.all(np.isfinite(q))
    assert np.all(np.diff(w) >= 0), "eigenvalues must be ascending"
    assert np.abs(w - wr).max() <= 4 * n * metrics.EPS * tn
    g = metrics.std_gates(t, w, q)
    assert g["residual_max"] < 10, g
    assert g["orth"] < 10, g


def test_stedc_large_random():
    n = 4096
    d, e = matgen.tridiag_family(n, "random", seed=1)
    w, q = _run(d, e)
    wr, zr, info = lapack.stedc(d, e)
    tn = np.abs(d).max() + 2 * np.abs(e).max()
    assert np.abs(w - wr).max() <= 4 * n * metrics.EPS * tn
    r = (d[:, None] * q)
    r[1:] += e[:, None] * q[:-1]
    r[:-1] += e[:, None] * q[1:]
    r -= q * w[None, :]
    assert np.abs(r).max() <= 20 * n * metrics.EPS * tn
    assert np.abs(q.T @ q - np.eye(n)).max() <= 20 * n * metrics.EPS


@pytest.mark.parametrize("kind", ["random", "toeplitz", "wilkinson", "glued"])
@pytest.mark.parametrize("n,c_lo,c_hi", [(100, 0, 10), (257, 30, 200), (1000, 0, 125), (1000, 990, 1000), (2500, 600, 601),
                                         (2500, 0, 0), (33, 5, 20), (20, 3, 9)])
def test_stedc_column_range_equals_the_full_solve(kind, n, c_lo, c_hi):
    """root merge restricted to the wanted eigenvector columns (il..iu subsets, the ranks' shares in the multi-GPU driver):
    same eigenvalues, and the wanted columns are those of the full solve"""
    from eigensolver_gpu_b200 import stages as S
    d, e = matgen.tridiag_family(n, kind, seed=n + 1)
    wf, qf = _run(d, e)
    w, q = S.stedc(S.to_dev(d), S.to_dev(e), cols=(c_lo, c_hi))
    w, q = S.to_host(w), np.array(S.to_host(q))
    assert np.array_equal(w, wf)
    assert np.abs(q[:, c_lo:c_hi] - qf[:, c_lo:c_hi]).max(initial=0.0) <= 1e-14


@pytest.mark.parametrize("cplx,n,il,iu", [(False, 1500, 1, 190), (True, 900, 401, 520), (False, 2100, 2000, 2100)])
def test_driver_subset_after_restricted_merge(cplx, n, il, iu):
    """il..iu subsets through the drop-in entry point (the restricted root merge is on its path), all gates vs LAPACK"""
    from eigensolver_gpu_b200 import api, stages as S
    a, b = matgen.family_c(n, cplx, seed=n)
    info, w, z, ws = api.solve_generalized(S.to_dev(np.triu(a)), S.to_dev(np.triu(b)), il, iu, skip_host_copy=True)
    assert info == 0
    w, z = S.to_host(w), np.array(S.to_host(z))
    wr, zr, ur, linfo = lapack.hegvd(a, b)
    assert np.abs(w - wr).max() < n * metrics.EPS * np.linalg.norm(a, 2)
    g = metrics.eig_gates(a, b, w[il - 1:iu], z)
    assert g["residual_max"] < 30 and g["b_orth"] < 30, g
    assert metrics.compare_2d_abs(zr[:, il - 1:iu], z)[0] < 1e-8
