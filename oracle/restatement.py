"""numpy restatement of the reference's stage algorithms (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Every function follows the reference routine named in its docstring (file:line relative to
/root/reference/lib_eigsolve).  Arrays are column-major numpy arrays, indices 0-based here (the citations
are 1-based Fortran).  Pure-numpy BLAS-2/3 calls stand in for the cuBLAS calls of the reference; the
host ?stedc call of the reference is LAPACK itself (oracle.lapack.stedc).

Parity pinning -- PARITY UNPINNED in the task's sense: the reference ships no golden vectors or known-answer tests for
this path and cannot be built or run here; as the strongest substitute this restatement is pinned in
tests/test_oracle.py against the reference's own ground truth, LAPACK ?sygvd/?hegvd, and against the
fixtures in tests/golden/.
"""
import numpy as np
import scipy.linalg as sla

from . import lapack


def _H(x):
    return x.conj().T


# ----------------------------------------------------------------------------------------------- stage (i)
def potrf_upper(b):
    """cusolverDn?potrf(UPPER) call site: zhegvdx_gpu.F90:135 / dsygvdx_gpu.F90:121.  B = U^H U."""
    return lapack.potrf(np.array(b, order="F"))


def hegst_reference(a, u, nb=448):
    """zhegst_gpu.F90:51-107 / dsygst_gpu.F90:51-96: A <- U^-H A U^-1 (upper), blocked, nb=448.

    Differences from LAPACK ?hegst that are encoded here: the diagonal block is Hermitian-completed and
    solved with two full TRSMs instead of ?hegs2 (zhegst_gpu.F90:58-71), its diagonal is re-realified
    (:74-81), and the two half updates use GEMM with the completed block instead of HEMM (:93-101).
    Input a: full or upper-populated matrix (only the upper triangle is read). Returns the upper triangle
    of C (strict lower part zero)."""
    a = np.triu(np.array(a, order="F"))
    n = a.shape[0]
    cplx = np.iscomplexobj(a)
    for k in range(0, n, nb):
        kb = min(n - k, nb)
        ukk = u[k:k + kb, k:k + kb]
        akk = a[k:k + kb, k:k + kb]
        akk = np.triu(akk) + _H(np.triu(akk, 1))                      # :58-65 complete the block
        akk = sla.solve_triangular(ukk, akk, trans="C", lower=False)   # :68-69 U_kk^-H A_kk
        akk = sla.solve_triangular(ukk, _H(akk), trans="C", lower=False)
        akk = _H(akk)                                                  # :70-71 (..) U_kk^-1
        if cplx:
            akk[np.diag_indices(kb)] = akk.diagonal().real             # :74-81
        a[k:k + kb, k:k + kb] = akk
        r = n - k - kb
        if r > 0:
            s = slice(k + kb, n)
            ak = sla.solve_triangular(ukk, a[k:k + kb, s], trans="C", lower=False)   # :87-88
            bk = u[k:k + kb, s]
            ak = ak - 0.5 * akk @ bk                                                 # :93-94
            a[s, s] -= np.triu(_H(ak) @ bk + _H(bk) @ ak)                            # :95-96 her2k, upper
            ak = ak - 0.5 * akk @ bk                                                 # :100-101
            # :103-104 right TRSM with U_22:  X U22 = ak
            ak = _H(sla.solve_triangular(u[s, s], _H(ak), trans="C", lower=False))
            a[k:k + kb, s] = ak
    return np.triu(a)


# ---------------------------------------------------------------------------------------------- stage (ii)
def larfg_reference(alpha, x):
    """zlarfg_kernel zhetrd_gpu.F90:211-333 / dlarfg_kernel dsytrd_gpu.F90:200-301.

    LAPACK ?larfg without the safmin rescaling loop.  Returns (beta, tau, scale) with v = scale * x,
    or (alpha.real, 0, None) when nothing is to be done (xnorm == 0 and Im alpha == 0)."""
    xnorm = np.linalg.norm(x)
    if np.iscomplexobj(alpha) or isinstance(alpha, complex):
        ar, ai = float(np.real(alpha)), float(np.imag(alpha))
        if xnorm == 0.0 and ai == 0.0:
            return ar, 0.0, None
        sc = max(abs(ar), abs(ai), xnorm)
        nrm = sc * np.sqrt((ar / sc) ** 2 + (ai / sc) ** 2 + (xnorm / sc) ** 2)   # dlapy3
        beta = -np.copysign(nrm, ar)
        tau = complex((beta - ar) / beta, -ai / beta)
        scale = 1.0 / (alpha - beta)                                              # zladiv in the kernel
        return beta, tau, scale
    a = float(alpha)
    if xnorm == 0.0:
        return a, 0.0, None
    beta = -np.copysign(np.hypot(a, xnorm), a)
    return beta, (beta - a) / beta, 1.0 / (a - beta)


def latrd_reference(a, n, nb, e, tau):
    """zlatrd_gpu zhetrd_gpu.F90:99-165 (+ kernels K10-K16) / dlatrd_gpu dsytrd_gpu.F90:98-164.

    Reduces the last nb columns of the leading n x n block of `a` (upper) in place; returns W (n x nb).
    The column update is K11 (zhetrd_gpu.F90:365-389), the reflector K10/K11-tail (:406-509), hemv K13
    (zhemv_gpu.F90:33-193, upper triangle only), stacked gemvs K14/K15 (:513-616, :750-879) and the
    tau/alpha fix-up of K15/K16 (:873-877, :660-748)."""
    cplx = np.iscomplexobj(a)
    w = np.zeros((n, nb), dtype=a.dtype)
    for i in range(n - 1, n - nb - 1, -1):          # 0-based column index i (Fortran i+1)
        iw = i - n + nb
        if i < n - 1:
            v = a[: i + 1, i + 1:n]
            ww = w[: i + 1, iw + 1:nb]
            a[: i + 1, i] -= v @ ww[i, :].conj() + ww @ v[i, :].conj()
            if cplx:
                a[i, i] = a[i, i].real
        if i > 0:
            alpha = a[i - 1, i]
            beta, t, scale = larfg_reference(alpha, a[: i - 1, i])
            if scale is not None:
                a[: i - 1, i] *= scale
                e[i - 1] = beta
            else:
                e[i - 1] = beta
            tau[i - 1] = t
            a[i - 1, i] = 1.0
            x = a[:i, i]
            au = a[:i, :i]
            full = np.triu(au) + _H(np.triu(au, 1))
            if cplx:
                full[np.diag_indices(i)] = full.diagonal().real
            wi = full @ x
            if i < n - 1:
                vv = a[:i, i + 1:n]
                ww = w[:i, iw + 1:nb]
                z1 = _H(vv) @ x
                z2 = _H(ww) @ x
                wi = wi - ww @ z1 - vv @ z2
            wi = t * wi
            al = -0.5 * t * np.vdot(wi, x)
            w[:i, iw] = wi + al * x
    return w


def hetrd_reference(a, nb=32):
    """zhetrd_gpu zhetrd_gpu.F90:30-96 / dsytrd_gpu dsytrd_gpu.F90:30-95: blocked tridiagonalization, UPLO='U'.

    Panels of nb from the last column leftwards (:60-71), a remainder panel so that a 32x32 block is left
    (:74-83), the final block by the unblocked kernel zhetd2_gpu.F90:41-188 (restated with the same panel
    recurrence, nb = remaining order), d(j)=A(j,j) (:90-94).  Returns (A_out, d, e, tau) with the reflectors
    v_j in A(0:j, j+1), unit element stored explicitly for panel columns (:92)."""
    a = np.array(a, order="F")
    n = a.shape[0]
    a = np.triu(a) + np.tril(a, -1)        # lower part is never read
    e = np.zeros(max(n - 1, 0))
    tau = np.zeros(max(n - 1, 0), dtype=a.dtype)
    if n > 32:
        kk = n - ((n - 32) // nb) * nb
        k = n
        i = n - nb
        while i >= kk:
            w = latrd_reference(a, i + nb, nb, e, tau)
            v = a[:i, i:i + nb]
            a[:i, :i] -= np.triu(v @ _H(w[:i]) + w[:i] @ _H(v))
            k -= nb
            i -= nb
        nbr = k - 32
        if nbr > 0:
            i = k - nbr
            w = latrd_reference(a, i + nbr, nbr, e, tau)
            v = a[:i, i:i + nbr]
            a[:i, :i] -= np.triu(v @ _H(w[:i]) + w[:i] @ _H(v))
    m = min(32, n)
    if m > 0:
        latrd_reference(a, m, m, e, tau)    # zhetd2_gpu.F90: same recurrence on the final block
    d = a.diagonal().real.copy()
    return a, d, e, tau


# --------------------------------------------------------------------------------------------- stage (iii)
def larft_reference(v, tau):
    """zlarft_gpu + finish_T_block_kernel zheevd_gpu.F90:136-176, 215-279 (d: dsyevd_gpu.F90:134-174, 212-276).

    v: mi x ib block already in 'unit lower-trapezoidal at the bottom' form (K22).  T0 = V^H V (lower),
    T(j,j)=tau_j, T(r,j) = -tau_j T0(r,j), then the backward column recurrence -> lower-triangular T
    (LAPACK ?larft('Backward','Columnwise'))."""
    ib = v.shape[1]
    t0 = _H(v) @ v
    t = np.zeros((ib, ib), dtype=v.dtype)
    for j in range(ib):
        t[j, j] = tau[j]
        t[j + 1:, j] = -tau[j] * t0[j + 1:, j]
    for c in range(ib - 2, -1, -1):
        t[c + 1:, c] = t[c + 1:, c + 1:] @ t[c + 1:, c]
    return t


def unmtr_reference(a, tau, z, nb2=64):
    """Back-transformation loop zheevd_gpu.F90:119-130 with zlarfb_gpu :178-213 (d: dsyevd_gpu.F90:117-128, 176-210).

    Z <- Q Z, Q = H(n-1)...H(1), applied in ascending blocks of nb2 reflectors:
    work = Z^H V; work <- work T^H; Z -= V work^H."""
    z = np.array(z, order="F")
    n = a.shape[0]
    k = n - 1
    for i in range(0, k, nb2):
        ib = min(nb2, k - i)
        mi = i + ib
        v = np.array(a[:mi, i + 1:i + 1 + ib])
        for j in range(ib):                     # K22: unit diagonal at the bottom, zeros below
            v[mi - ib + j, j] = 1.0
            v[mi - ib + j + 1:, j] = 0.0
        t = larft_reference(v, tau[i:i + ib])
        wk = _H(z[:mi]) @ v
        wk = wk @ _H(t)
        z[:mi] -= v @ _H(wk)
    return z


# ------------------------------------------------------------------------------------------------- drivers
def heevd_reference(a, il, iu, nb1=32, nb2=64):
    """zheevd_gpu zheevd_gpu.F90:32-134 / dsyevd_gpu dsyevd_gpu.F90:32-132: hetrd -> host stedc -> back-transform.
    il, iu 1-based inclusive.  Returns (w[all n], Z[:, :iu-il+1])."""
    a2, d, e, tau = hetrd_reference(a, nb1)
    w, zt, info = lapack.stedc(d, e)
    if info != 0:
        raise RuntimeError("stedc failed")
    z = zt[:, il - 1:iu].astype(a2.dtype)
    z = unmtr_reference(a2, tau, z, nb2)
    return w, z


def hegvdx_reference(a, b, il, iu):
    """zhegvdx_gpu zhegvdx_gpu.F90:75-182 / dsygvdx_gpu dsygvdx_gpu.F90:71-168.
    Returns (w[all n], Z[:, :m], U) with Z^H B Z = I."""
    u = potrf_upper(b)                                   # :135
    c = hegst_reference(a, u, 448)                       # :156-158
    w, z = heevd_reference(c, il, iu)                    # :163-164
    z = sla.solve_triangular(u, z, lower=False)          # :169
    return w, z, u


def check_workspace(n, cplx, lwork, lrwork, lwork_h, lrwork_h, liwork_h):
    """Workspace checks zhegvdx_gpu.F90:106-127 / dsygvdx_gpu.F90:100-113. Returns info (0 or -1)."""
    if cplx:
        if lwork < 2 * 64 * 64 + 65 * n or lrwork < n or lwork_h < n or lrwork_h < 1 + 5 * n + 2 * n * n or liwork_h < n:
            return -1
    else:
        if lwork < 2 * 64 * 64 + 66 * n or lwork_h < 1 + 6 * n + 2 * n * n or liwork_h < n:
            return -1
    return 0
