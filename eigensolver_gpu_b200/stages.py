"""Stage-level Python wrappers over the C ABI (device tensors in, device tensors out).

Column-major convention: a matrix with n rows and m columns is held in a torch tensor of shape (m, n)
(row-major torch memory == column-major (n, m)); `to_dev` / `to_host` convert from/to numpy arrays.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import check, sync_stream as load     # every stage call is issued on torch's current stream


def to_dev(a, device="cuda"):
    """numpy (n, m) array -> device tensor of shape (m, n) whose memory is the column-major matrix."""
    a = np.asarray(a)
    if a.ndim == 1:
        return torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return torch.from_numpy(np.ascontiguousarray(a.T)).to(device)


def to_host(t):
    """inverse of to_dev: returns an (n, m) numpy array (Fortran-ordered view)."""
    a = t.detach().cpu().numpy()
    return a if a.ndim == 1 else a.T


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _ld(t):
    return int(t.shape[1]) if t.dim() == 2 else int(t.shape[0])


def _is_c(t):
    return t.dtype == torch.complex128


def _ch(c):
    return C.c_char(c.encode())


def gemm(ta, tb, alpha, a, b, beta, c, m=None, n=None, k=None):
    """C = alpha op(A) op(B) + beta C on column-major device tensors (shape (cols, ld))."""
    lib = load()
    if m is None:
        m = a.shape[1] if ta == "N" else a.shape[0]
    if k is None:
        k = a.shape[0] if ta == "N" else a.shape[1]
    if n is None:
        n = b.shape[0] if tb == "N" else b.shape[1]
    fn = lib.eigb200_zgemm if _is_c(a) else lib.eigb200_dgemm
    check(fn(_ch(ta), _ch(tb), m, n, k, alpha, _ptr(a), _ld(a), _ptr(b), _ld(b), beta, _ptr(c), _ld(c)), "gemm")
    return c


def her2k(alpha, a, b, beta, c, n=None, k=None):
    lib = load()
    n = a.shape[1] if n is None else n
    k = a.shape[0] if k is None else k
    fn = lib.eigb200_zher2k if _is_c(a) else lib.eigb200_dsyr2k
    check(fn(n, k, alpha, _ptr(a), _ld(a), _ptr(b), _ld(b), beta, _ptr(c), _ld(c)), "her2k")
    return c


def hemv(a, x, n=None):
    lib = load()
    n = a.shape[0] if n is None else n
    y = torch.empty_like(x)
    fn = lib.eigb200_zhemv if _is_c(a) else lib.eigb200_dsymv
    check(fn(n, _ptr(a), _ld(a), _ptr(x), _ptr(y)), "hemv")
    return y


def hetrd(a):
    """In-place tridiagonalization (upper). Returns (d, e, tau) device tensors."""
    lib = load()
    n = a.shape[0]
    d = torch.zeros(n, dtype=torch.float64, device=a.device)
    e = torch.zeros(max(n - 1, 1), dtype=torch.float64, device=a.device)
    tau = torch.zeros(max(n - 1, 1), dtype=a.dtype, device=a.device)
    fn = lib.eigb200_zhetrd if _is_c(a) else lib.eigb200_dsytrd
    check(fn(n, _ptr(a), _ld(a), _ptr(d), _ptr(e), _ptr(tau)), "hetrd")
    return d, e[: n - 1], tau[: n - 1]


def stedc(d, e, cols=None):
    """All eigenvalues of the tridiagonal (d, e) on the device and the eigenvectors (all, or only the sorted columns
    cols = (c_lo, c_hi), 0-based half open -- the others are undefined). Returns (w, Q), Q an (n, n) column-major tensor."""
    lib = load()
    n = d.shape[0]
    w = d.clone()
    ee = torch.zeros(max(n, 1), dtype=torch.float64, device=d.device)
    ee[: n - 1] = e[: n - 1]
    q = torch.zeros((n, n), dtype=torch.float64, device=d.device)
    if cols is None:
        check(lib.eigb200_dstedc(n, _ptr(w), _ptr(ee), _ptr(q), n), "stedc")
    else:
        check(lib.eigb200_dstedc_range(n, _ptr(w), _ptr(ee), _ptr(q), n, cols[0], cols[1]), "stedc")
    return w, q


def ormtr(a, tau, z, m=None):
    lib = load()
    n = a.shape[0]
    m = z.shape[0] if m is None else m
    fn = lib.eigb200_zunmtr if _is_c(a) else lib.eigb200_dormtr
    check(fn(n, m, _ptr(a), _ld(a), _ptr(tau), _ptr(z), _ld(z)), "ormtr")
    return z


def potrf(b):
    lib = load()
    n = b.shape[0]
    info = C.c_int(0)
    fn = lib.eigb200_zpotrf if _is_c(b) else lib.eigb200_dpotrf
    check(fn(n, _ptr(b), _ld(b), C.byref(info)), "potrf")
    return info.value


def hegst(a, u):
    lib = load()
    n = a.shape[0]
    fn = lib.eigb200_zhegst if _is_c(a) else lib.eigb200_dsygst
    check(fn(n, _ptr(a), _ld(a), _ptr(u), _ld(u)), "hegst")
    return a


def trsm(side, trans, u, b, m=None, n=None):
    lib = load()
    m = b.shape[1] if m is None else m
    n = b.shape[0] if n is None else n
    fn = lib.eigb200_ztrsm if _is_c(b) else lib.eigb200_dtrsm
    check(fn(_ch(side), _ch(trans), m, n, _ptr(u), _ld(u), _ptr(b), _ld(b)), "trsm")
    return b
