"""CPU tests of the oracle: the numpy restatement of the reference stages against the reference's own ground
truth (LAPACK ?sygvd/?hegvd, test_driver/test_zhegvdx.F90:163-182) and against the committed fixtures."""
import numpy as np
import pytest

from oracle import lapack, matgen, metrics, restatement as R


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n,m", [(1, 1), (2, 2), (33, 33), (97, 20), (300, 64)])
def test_restatement_matches_lapack_hegvd(cplx, n, m):
    a, b = matgen.family_c(n, cplx, seed=11 + n)
    w, z, u = R.hegvdx_reference(a, b, 1, m)
    wr, zr, ur, info = lapack.hegvd(a, b)
    assert info == 0
    g = metrics.eig_gates(a, b, w, z, wr)
    assert g["dlambda_over_gate"] < 1.0          # |dlambda| < n eps ||A||
    assert g["residual_max"] < 30.0              # north_star residual gate
    assert g["b_orth"] < 30.0
    rel, _ = metrics.compare_2d_abs(zr[:, :m], z)
    assert rel < 1e-9


@pytest.mark.parametrize("cplx", [False, True])
def test_restatement_stage_outputs_match_lapack(cplx):
    n = 157
    a, b = matgen.family_c(n, cplx, seed=5)
    u = R.potrf_upper(b)
    assert np.allclose(u.conj().T @ u, b, atol=1e-12 * np.linalg.norm(b))
    c = R.hegst_reference(a, u, 64)
    cl = np.triu(lapack.hegst(a, u))
    assert np.abs(c - cl).max() < 1e-11 * np.abs(cl).max()
    cf = metrics.full_from_upper(c)
    a2, d, e, tau = R.hetrd_reference(cf)
    al, dl, el, taul = lapack.hetrd(cf)
    assert np.abs(d - dl).max() < 1e-11 * np.abs(dl).max()
    assert np.abs(np.abs(e) - np.abs(el)).max() < 1e-11 * np.abs(el).max()
    w, zt, info = lapack.stedc(d, e)
    assert info == 0
    z = R.unmtr_reference(a2, tau, zt.astype(cf.dtype))
    zl = lapack.ormtr("L", "U", "N", a2, tau, zt.astype(cf.dtype))
    assert np.abs(z - zl).max() < 1e-12
    g = metrics.std_gates(cf, w, z)
    assert g["residual_max"] < 10 and g["orth"] < 10


def test_workspace_checks_follow_reference():
    n = 100
    assert R.check_workspace(n, True, 2 * 64 * 64 + 65 * n, n, n, 1 + 5 * n + 2 * n * n, n) == 0
    assert R.check_workspace(n, True, 2 * 64 * 64 + 65 * n - 1, n, n, 1 + 5 * n + 2 * n * n, n) == -1
    assert R.check_workspace(n, True, 2 * 64 * 64 + 65 * n, n - 1, n, 1 + 5 * n + 2 * n * n, n) == -1
    assert R.check_workspace(n, False, 2 * 64 * 64 + 66 * n, 0, 1 + 6 * n + 2 * n * n, 0, n) == 0
    assert R.check_workspace(n, False, 2 * 64 * 64 + 66 * n, 0, 6 * n + 2 * n * n, 0, n) == -1
    assert R.check_workspace(n, False, 2 * 64 * 64 + 66 * n, 0, 1 + 6 * n + 2 * n * n, 0, n - 1) == -1


@pytest.mark.parametrize("kind", ["random", "toeplitz", "wilkinson", "glued", "graded", "zero_e", "identity"])
def test_stedc_oracle_is_accurate(kind):
    n = 201
    d, e = matgen.tridiag_family(n, kind, seed=3)
    w, z, info = lapack.stedc(d, e)
    assert info == 0
    t = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    g = metrics.std_gates(t, w, z)
    assert g["residual_max"] < 5 and g["orth"] < 5
    assert np.all(np.diff(w) >= 0)
