"""ctypes bindings to the Fortran-ABI LAPACK/BLAS in SciPy's bundled OpenBLAS (test infrastructure).

The reference's ground truth is CPU LAPACK ?hegvd / ?sygvd (test_driver/test_zhegvdx.F90:163-182,
test_driver/test_dsygvdx.F90:189-208); its host stage is ?stedc('I') (lib_eigsolve/zheevd_gpu.F90:101,
dsyevd_gpu.F90:99).  scipy.linalg.lapack does not wrap ?stedc / ?ormtr / ?unmtr / ?laed4, so they are
bound here directly.  All arrays are Fortran (column-major) order.
"""
import ctypes as C
import glob
import os

import numpy as np

_lib = None


def lib():
    global _lib
    if _lib is None:
        import scipy
        pat = os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so")
        cands = sorted(glob.glob(pat))
        if not cands:
            raise RuntimeError("oracle: libscipy_openblas not found next to scipy")
        _lib = C.CDLL(cands[0])
    return _lib


def num_threads():
    return int(lib().scipy_openblas_get_num_threads())


def set_num_threads(n):
    lib().scipy_openblas_set_num_threads(C.c_int(int(n)))


def _i(v):
    return C.byref(C.c_int(int(v)))


def _d(v):
    return C.byref(C.c_double(float(v)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _ch(c):
    return C.c_char_p(c.encode())


_ONE = C.c_size_t(1)


def _f(a, dtype):
    a = np.asarray(a, dtype=dtype)
    return np.array(a, order="F", copy=True)


def stedc(d, e, compz="I"):
    """?stedc('I'): all eigenpairs of the symmetric tridiagonal (d, e). Returns (w, Z, info)."""
    n = len(d)
    d = np.array(d, dtype=np.float64, copy=True)
    e = np.array(np.concatenate([e, [0.0]]) if len(e) == n - 1 else e, dtype=np.float64, copy=True)
    z = np.zeros((n, n), dtype=np.float64, order="F")
    lwork = 1 + 4 * n + n * n + 64
    liwork = 3 + 5 * n + 64
    work = np.zeros(lwork)
    iwork = np.zeros(liwork, dtype=np.int32)
    info = C.c_int(0)
    lib().scipy_dstedc_(_ch(compz), _i(n), _p(d), _p(e), _p(z), _i(max(1, n)), _p(work), _i(lwork), _p(iwork),
                        _i(liwork), C.byref(info), _ONE)
    return d, z, info.value


def laed4(n, i, d, z, rho):
    """dlaed4: i-th (1-based) root of the secular equation. Returns (delta, lam, info)."""
    d = np.array(d, dtype=np.float64)
    z = np.array(z, dtype=np.float64)
    delta = np.zeros(n)
    lam = C.c_double(0)
    info = C.c_int(0)
    lib().scipy_dlaed4_(_i(n), _i(i), _p(d), _p(z), _p(delta), _d(rho), C.byref(lam), C.byref(info))
    return delta, lam.value, info.value


def ormtr(side, uplo, trans, a, tau, c):
    """?ormtr / ?unmtr: C <- op(Q) C with Q from ?sytrd/?hetrd. Returns new C."""
    cplx = np.iscomplexobj(a)
    dt = np.complex128 if cplx else np.float64
    a = _f(a, dt)
    c = _f(c, dt)
    tau = np.array(tau, dtype=dt)
    m, n = c.shape
    lwork = max(1, 64 * max(m, n))
    work = np.zeros(lwork, dtype=dt)
    info = C.c_int(0)
    fn = lib().scipy_zunmtr_ if cplx else lib().scipy_dormtr_
    fn(_ch(side), _ch(uplo), _ch(trans), _i(m), _i(n), _p(a), _i(a.shape[0]), _p(tau), _p(c), _i(m), _p(work),
       _i(lwork), C.byref(info), _ONE, _ONE, _ONE)
    if info.value != 0:
        raise RuntimeError(f"ormtr info={info.value}")
    return c


def hegvd(a, b, jobz="V", uplo="U"):
    """?sygvd / ?hegvd (ITYPE=1) -- the reference's ground truth. Returns (w, Z, U, info).

    Called directly (not via scipy.linalg) so that the bench can time exactly one LAPACK call with
    pre-sized workspaces, as test_driver/test_zhegvdx.F90:163-182 does."""
    cplx = np.iscomplexobj(a) or np.iscomplexobj(b)
    dt = np.complex128 if cplx else np.float64
    a = _f(a, dt)
    b = _f(b, dt)
    n = a.shape[0]
    w = np.zeros(n)
    info = C.c_int(0)
    if cplx:
        lwork, lrwork, liwork = 2 * n + n * n + 64, 1 + 5 * n + 2 * n * n + 64, 3 + 5 * n + 64
        work = np.zeros(lwork, dtype=dt)
        rwork = np.zeros(lrwork)
        iwork = np.zeros(liwork, dtype=np.int32)
        lib().scipy_zhegvd_(_i(1), _ch(jobz), _ch(uplo), _i(n), _p(a), _i(n), _p(b), _i(n), _p(w), _p(work), _i(lwork),
                            _p(rwork), _i(lrwork), _p(iwork), _i(liwork), C.byref(info), _ONE, _ONE)
    else:
        lwork, liwork = 1 + 6 * n + 2 * n * n + 64, 3 + 5 * n + 64
        work = np.zeros(lwork)
        iwork = np.zeros(liwork, dtype=np.int32)
        lib().scipy_dsygvd_(_i(1), _ch(jobz), _ch(uplo), _i(n), _p(a), _i(n), _p(b), _i(n), _p(w), _p(work), _i(lwork),
                            _p(iwork), _i(liwork), C.byref(info), _ONE, _ONE)
    return w, a, b, info.value


def potrf(b):
    """?potrf('U'): B = U^H U. Returns upper-triangular U (strict lower part zeroed)."""
    import scipy.linalg.lapack as L
    fn = L.zpotrf if np.iscomplexobj(b) else L.dpotrf
    u, info = fn(b, lower=0, clean=1)
    if info != 0:
        raise RuntimeError(f"potrf info={info}")
    return u


def hegst(a, u):
    """?sygst / ?hegst(ITYPE=1,'U'): A <- U^-H A U^-1 (upper triangle returned, Hermitian-completed)."""
    import scipy.linalg.lapack as L
    cplx = np.iscomplexobj(a) or np.iscomplexobj(u)
    fn = L.zhegst if cplx else L.dsygst
    c, info = fn(a, u, itype=1, lower=0)
    if info != 0:
        raise RuntimeError(f"hegst info={info}")
    c = np.triu(c)
    c = c + np.triu(c, 1).conj().T
    return c


def hetrd(a):
    """?sytrd / ?hetrd('U'). Returns (A_out, d, e, tau)."""
    import scipy.linalg.lapack as L
    fn = L.zhetrd if np.iscomplexobj(a) else L.dsytrd
    c, d, e, tau, info = fn(a, lower=0)
    if info != 0:
        raise RuntimeError(f"hetrd info={info}")
    return c, d, e, tau


def hemv_upper(a, x):
    """?symv / ?hemv('U'): y = A x reading only the upper triangle of A."""
    au = np.triu(a)
    full = au + np.triu(a, 1).conj().T
    if np.iscomplexobj(a):
        full[np.diag_indices_from(full)] = np.real(np.diag(a))
    return full @ x
