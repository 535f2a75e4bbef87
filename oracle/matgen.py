"""Seeded input families for the generalized eigenproblem (test infrastructure).

Family R: the reference recipe -- T Hermitian with entries U[0,1) (+ i U[0,1) off the diagonal, real
diagonal), A = T_A T_A^H, B = T_B T_B^H (test_driver/test_zhegvdx.F90:28-66, test_dsygvdx.F90:28-64).
Used for like-for-like timing; cond(B) ~ 1e8..3e10, so LAPACK itself misses the n*eps residual gate on it
(SURVEY.md section 4).

Family C: conditioned pair for the parity gates -- A = (G + G^H)/2 with G iid N(0,1) (+ i N(0,1)),
B = T_B T_B^H / N + I (cond ~ 1e2).
"""
import numpy as np


def _rng(seed):
    return np.random.Generator(np.random.Philox(int(seed)))


def _herm_uniform(n, cplx, rng):
    t = rng.random((n, n))
    if cplx:
        t = t + 1j * rng.random((n, n))
    t = np.triu(t) + np.triu(t, 1).conj().T
    if cplx:
        t[np.diag_indices(n)] = t.diagonal().real
    return t


def family_r(n, cplx=False, seed=1234):
    rng = _rng(seed)
    ta = _herm_uniform(n, cplx, rng)
    tb = _herm_uniform(n, cplx, rng)
    a = ta @ ta.conj().T
    b = tb @ tb.conj().T
    return np.asfortranarray(_sym(a)), np.asfortranarray(_sym(b))


def family_c(n, cplx=False, seed=1234):
    rng = _rng(seed)
    g = rng.standard_normal((n, n))
    if cplx:
        g = g + 1j * rng.standard_normal((n, n))
    a = (g + g.conj().T) / 2
    tb = _herm_uniform(n, cplx, rng)
    b = tb @ tb.conj().T / n + np.eye(n)
    return np.asfortranarray(_sym(a)), np.asfortranarray(_sym(b))


def _sym(a):
    a = (a + a.conj().T) / 2
    if np.iscomplexobj(a):
        a[np.diag_indices_from(a)] = a.diagonal().real
    return a


def tridiag_family(n, kind, seed=1234):
    """Tridiagonal test matrices for the stedc stage: (d, e)."""
    rng = _rng(seed)
    if kind == "random":
        return rng.standard_normal(n), rng.standard_normal(n - 1)
    if kind == "toeplitz":       # 1-2-1: clustered at the ends, heavy deflation at high levels
        return 2.0 * np.ones(n), np.ones(n - 1)
    if kind == "wilkinson":      # W_n^+: pairs of pathologically close eigenvalues
        m = (n - 1) / 2.0
        return np.abs(np.arange(n) - m), np.ones(n - 1)
    if kind == "glued":          # nearly-split blocks: tiny off-diagonals
        d = rng.standard_normal(n)
        e = rng.standard_normal(n - 1)
        e[:: max(2, n // 7)] *= 1e-14
        return d, e
    if kind == "graded":
        s = np.logspace(0, -10, n)
        return s * rng.standard_normal(n), s[:-1] * rng.standard_normal(n - 1)
    if kind == "zero_e":
        return rng.standard_normal(n), np.zeros(n - 1)
    if kind == "identity":
        return np.ones(n), np.zeros(n - 1)
    raise ValueError(kind)
