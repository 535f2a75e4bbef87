"""Parity at the BENCHMARKED orders: the code paths that only switch on for large matrices -- L2
eviction hints of the tile stream (triangle > 100 MB), hegst block 2048, the chunked 256-leaf TRSM, 64-bit indexing at
N=16384 -- compared with LAPACK where the CPU finishes in seconds and through size-independent gates (residual,
B-orthogonality, ascending eigenvalues; computed on the device with torch matmul) at BASELINE.json's full sizes.
Reference comparator: test_driver/test_zhegvdx.F90:163-182, 297-303."""
import numpy as np
import pytest
import torch

import bench
from oracle import lapack, matgen, metrics

pytestmark = pytest.mark.gpu


def _gates(a0, b0, w, z, m):
    g = bench.parity_metrics(torch, a0, b0, w, z, m)
    assert g["finite"] and g["w_ascending"], g
    assert g["residual_max"] < 30, g
    assert g["b_orth"] < 30, g
    return g


def test_dsygvdx_n4096_m512_matches_lapack():
    """BASELINE.json configs[1]: DSYGVDX N=4096, il=1..512 -- every gate, against LAPACK dsygvd on the same input"""
    from eigensolver_gpu_b200 import api, stages as S
    n, m = 4096, 512
    a, b = matgen.family_c(n, False, seed=4096)
    ad, bd = S.to_dev(np.triu(a)), S.to_dev(np.triu(b))
    info, w, z, ws = api.solve_generalized(ad, bd, 1, m, skip_host_copy=False)
    assert info == 0
    w = S.to_host(w)
    z = np.array(S.to_host(z))
    lapack.set_num_threads(__import__("os").cpu_count())
    wr, zr, ur, linfo = lapack.hegvd(a, b)
    assert linfo == 0
    v = np.ones(n)
    for _ in range(60):                       # ||A||_2 by power iteration (from below: the stricter gate)
        v = a @ v
        an = np.linalg.norm(v)
        v /= an
    assert np.abs(w - wr).max() < n * metrics.EPS * an
    g = metrics.eig_gates(a, b, w[:m], z)
    assert g["residual_max"] < 30 and g["b_orth"] < 30, g
    assert metrics.compare_1d(wr, w)[0] < 1e-13
    assert metrics.compare_2d_abs(zr[:, :m], z)[0] < 1e-8
    assert np.array_equal(ws.Z_h.numpy().T[:, :m], z)


@pytest.mark.parametrize("cplx,n,m", [(True, 8192, 8192), (False, 16384, 2048)])
def test_full_size_configs_pass_the_gates(cplx, n, m):
    """BASELINE.json configs[2] (ZHEGVDX N=8192 full spectrum) and configs[3] (DSYGVDX N=16384, il=1..2048) on one GPU"""
    from eigensolver_gpu_b200 import api
    a0, b0 = bench.make_inputs(torch, n, cplx, "C", 99)
    A, B = a0.clone(), b0.clone()
    info, w, z, ws = api.solve_generalized(A, B, 1, m, skip_host_copy=True)
    assert info == 0
    del A, B
    _gates(a0, b0, w, z, m)
    # eigenvalue cross-check that does not involve this library: generalized Rayleigh quotients of the computed vectors
    Z = z[:m].T
    num = (Z.conj() * (a0.T @ Z)).sum(dim=0).real
    den = (Z.conj() * (b0.T @ Z)).sum(dim=0).real
    assert float((num / den - w[:m]).abs().max()) < n * metrics.EPS * float(torch.linalg.matrix_norm(a0, 1))


def test_hetrd_n6000_l2_hints_do_not_change_results():
    """the eviction-priority hints only steer the cache: (d, e, tau) must be bitwise identical with and without them,
    and T must have A's eigenvalues (checked against torch.linalg.eigvalsh, an independent implementation)"""
    from eigensolver_gpu_b200 import stages as S
    from eigensolver_gpu_b200._lib import load
    import scipy.linalg as sl
    n = 6000
    lib = load()
    a0, _ = bench.make_inputs(torch, n, True, "C", 6)
    outs = []
    try:
        for mb in (0, 32):
            assert lib.eigb200_set_option(b"trd_l2keep_mb", mb) == 0
            A = a0.clone()
            d, e, tau = S.hetrd(A)
            outs.append((d.clone(), e.clone(), tau.clone()))
    finally:
        lib.eigb200_set_option(b"trd_l2keep_mb", 32)
    for x, y in zip(outs[0], outs[1]):
        assert torch.equal(x, y)
    d, e, _ = outs[0]
    wt = sl.eigvalsh_tridiagonal(d.cpu().numpy(), e.cpu().numpy())
    wa = torch.linalg.eigvalsh(a0.T).cpu().numpy()
    an = float(torch.linalg.matrix_norm(a0, 1))
    assert np.abs(wt - wa).max() < n * metrics.EPS * an


@pytest.mark.parametrize("cplx", [False, True])
def test_hegst_large_blocks_match_lapack(cplx):
    """n = 4200 > 4096: block size 2048 and 256-wide inverted leaves (trsm.cu) -- vs LAPACK ?hegst on the same input"""
    from eigensolver_gpu_b200 import stages as S
    n = 4200
    a, b = matgen.family_c(n, cplx, seed=42)
    u = lapack.potrf(b)
    ad, ud = S.to_dev(np.triu(a)), S.to_dev(u)
    S.hegst(ad, ud)
    c = np.triu(np.array(S.to_host(ad)))
    cr = np.triu(lapack.hegst(a, u))
    assert np.abs(c - cr).max() <= 50 * n * metrics.EPS * np.abs(cr).max()


def test_tma_gemm_equals_cpasync_gemm_on_a_long_reduction():
    """Regression for round 2's TMA incident ("TMA-fed GEMM: what went wrong"): hegst at n = 10240 with 2048-wide
    blocks is several hundred back-to-back GEMM launches with long k loops; the TMA-fed kernel and the cp.async kernel
    accumulate in the same order, so the two reductions must agree BITWISE"""
    from eigensolver_gpu_b200 import stages as S
    from eigensolver_gpu_b200._lib import load
    lib = load()
    n = 10240
    t = torch.rand((n, n), dtype=torch.float64, device="cuda")
    bm = t @ t.T / n + torch.eye(n, dtype=torch.float64, device="cuda")
    del t
    g = torch.randn((n, n), dtype=torch.float64, device="cuda")
    am = (g + g.T) / 2
    del g
    outs = []
    try:
        assert lib.eigb200_set_option(b"hegst_hb", 2048) == 0
        for tma in (1, 0, 1):
            assert lib.eigb200_set_option(b"gemm_tma", tma) == 0
            u = bm.clone()
            assert S.potrf(u) == 0
            outs.append((torch.tril(u), torch.tril(S.hegst(am.clone(), u))))
    finally:
        lib.eigb200_set_option(b"gemm_tma", 1)
        lib.eigb200_set_option(b"hegst_hb", 0)
    for k in (0, 2):
        assert torch.equal(outs[k][0], outs[1][0]) and torch.equal(outs[k][1], outs[1][1])


def test_dsygvdx_n12288_gates():
    """an order between the benchmarked ones (the first size at which round 2's TMA incident showed)"""
    from eigensolver_gpu_b200 import api
    n, m = 12288, 1536
    a0, b0 = bench.make_inputs(torch, n, False, "C", 7)
    A, B = a0.clone(), b0.clone()
    info, w, z, ws = api.solve_generalized(A, B, 1, m, skip_host_copy=True)
    assert info == 0
    del A, B
    _gates(a0, b0, w, z, m)
