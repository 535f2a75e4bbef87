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
