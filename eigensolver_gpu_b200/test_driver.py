"""Host-side mirror of the reference's test programs (test_driver/test_zhegvdx.F90:75-305, test_dsygvdx.F90:90-330).

    python -m eigensolver_gpu_b200.test_driver z|d N [il iu]          # random SPD pair, the reference recipe
    python -m eigensolver_gpu_b200.test_driver z|d fileA fileB [il iu] # matrices dumped by the application

Like the reference program it solves the same generalized problem with the CPU LAPACK driver (?hegvd / ?sygvd -- here
SciPy's bundled OpenBLAS instead of MKL) and with the custom GPU solver (`zhegvdx_gpu` / `dsygvdx_gpu` through the C ABI)
and prints the comparison the reference prints (`compare` in test_driver/toolbox.F90:26-176: relative L2 error and
largest per-element error in percent, eigenvalues first, then |Z|).  The MAGMA and cuSOLVER cases of the reference
program are not reproduced: neither library is on this product's path.

File format (test_dsygvdx.F90:120-145): Fortran unformatted sequential, record 1 = three default integers
`n, m, lda`, record 2 = `A(1:n,1:n)` (column-major, real(8) or complex(8)); every record is framed by 4-byte length
markers (gfortran / nvfortran / ifort default).
"""
import struct
import sys
import time

import numpy as np


# ----------------------------------------------------------------------------------------- unformatted files
def write_unformatted(path, a, m=None, lda=None):
    """Writes `n, m, lda` and A(1:n,1:n) the way the reference's dump is laid out (record markers: int32)."""
    a = np.asarray(a)
    n = a.shape[0]
    m = n if m is None else m
    lda = n if lda is None else lda
    with open(path, "wb") as f:
        hdr = struct.pack("<3i", n, m, lda)
        f.write(struct.pack("<i", len(hdr)) + hdr + struct.pack("<i", len(hdr)))
        body = np.asfortranarray(a).tobytes(order="F")
        if len(body) >= 2 ** 31:
            raise ValueError("record larger than 2 GiB: sub-records are not supported")
        f.write(struct.pack("<i", len(body)) + body + struct.pack("<i", len(body)))


def read_unformatted(path, cplx):
    """Returns (A[n,n], n, m, lda) from a file written by the application (see module docstring)."""
    with open(path, "rb") as f:
        (l0,) = struct.unpack("<i", f.read(4))
        if l0 != 12:
            raise ValueError("%s: first record is not three default integers (length %d)" % (path, l0))
        n, m, lda = struct.unpack("<3i", f.read(12))
        (l1,) = struct.unpack("<i", f.read(4))
        if l1 != l0:
            raise ValueError("%s: corrupt record marker" % path)
        (lb,) = struct.unpack("<i", f.read(4))
        es = 16 if cplx else 8
        if lb != n * n * es:
            raise ValueError("%s: matrix record has %d bytes, expected n*n*%d = %d" % (path, lb, es, n * n * es))
        a = np.frombuffer(f.read(lb), dtype=np.complex128 if cplx else np.float64).reshape((n, n), order="F").copy()
        (le,) = struct.unpack("<i", f.read(4))
        if le != lb:
            raise ValueError("%s: corrupt trailing record marker" % path)
    return a, n, m, lda


# ----------------------------------------------------------------------------------------- reference recipe
def create_random_pd(n, cplx, seed=0):
    """create_random_hermetian_pd / create_random_symmetric_pd (test_zhegvdx.F90:28-66, test_dsygvdx.F90:28-60):
    T Hermitian with U[0,1) entries (real diagonal), A = T T^H."""
    rng = np.random.default_rng(seed)
    t = rng.random((n, n))
    if cplx:
        t = t + 1j * rng.random((n, n))
    t = np.triu(t) + np.triu(t, 1).conj().T
    if cplx:
        t[np.diag_indices(n)] = t[np.diag_indices(n)].real
    a = t @ t.conj().T
    return (a + a.conj().T) / 2


# ----------------------------------------------------------------------------------------- compare (toolbox.F90)
def compare_1d(ref, x):
    ref, x = np.asarray(ref, dtype=float), np.asarray(x, dtype=float)
    sel = np.abs(ref) >= 1e-10
    l2 = np.sqrt(np.sum((ref[sel] - x[sel]) ** 2)) / max(np.sqrt(np.sum(ref[sel] ** 2)), 1e-300)
    perr = np.zeros_like(ref)
    perr[sel] = np.abs(ref[sel] - x[sel]) / np.abs(ref[sel]) * 100.0
    i = int(np.argmax(perr))
    return l2, perr[i], i


def compare_2d_abs(ref, x):
    """the reference compares moduli (eigenvectors are defined up to a phase): toolbox.F90:100-176"""
    r, y = np.abs(ref), np.abs(x)
    sel = r >= 1e-10
    l2 = np.sqrt(np.sum((r[sel] - y[sel]) ** 2)) / max(np.sqrt(np.sum(r[sel] ** 2)), 1e-300)
    perr = np.zeros_like(r)
    perr[sel] = np.abs(r[sel] - y[sel]) / r[sel] * 100.0
    i, j = np.unravel_index(int(np.argmax(perr)), perr.shape)
    return l2, perr[i, j], i, j


def _report(w_ref, z_ref, w, z):
    l2, mx, i = compare_1d(w_ref, w)
    if l2 == 0.0:
        print("     EXACT MATCH")
    else:
        print("    l2norm error  %10.3E   max error%10.3E  %% at%5d   cpu=  %20.14E   gpu=  %20.14E" % (l2, mx, i + 1, w_ref[i], w[i]))
    l2, mx, i, j = compare_2d_abs(z_ref, z)
    if l2 == 0.0:
        print("     EXACT MATCH")
    else:
        print("    l2norm error  %10.3E   max error%10.3E  %% at%5d%5d   |cpu|=  %20.14E   |gpu|=  %20.14E"
              % (l2, mx, i + 1, j + 1, abs(z_ref[i, j]), abs(z[i, j])))
    return l2


def run(cplx, a, b, il=1, iu=None, out=sys.stdout):
    """The CPU and CUSTOM cases of the reference program on one (A, B) pair; returns (w_gpu, z_gpu, info)."""
    import scipy.linalg as sla
    import torch
    from . import api, stages as S
    n = a.shape[0]
    iu = n if iu is None else iu
    m = iu - il + 1
    print(" Running with N = %d" % n)
    print("\n CPU_____________________")
    sla.eigh(a, b, driver="gvd")                                   # run once before timing (test_zhegvdx.F90:171)
    t0 = time.time()
    w1, z1 = sla.eigh(a, b, driver="gvd")
    print(" \tTime for CPU %s = %10.3f" % ("zhegvd" if cplx else "dsygvd", (time.time() - t0) * 1e3))
    print("\n CUSTOM_____________________")
    api.init_eigsolve_gpu()
    ws = api.Workspace(n, cplx, host_z=True)
    info = -1
    for rep in range(2):                                           # the reference times the second call as well
        ad, bd = S.to_dev(np.triu(a)), S.to_dev(np.triu(b))
        torch.cuda.synchronize()
        t0 = time.time()
        info, w, z, _ = api.solve_generalized(ad, bd, il, iu, ws=ws, skip_host_copy=False)
        torch.cuda.synchronize()
        t1 = time.time() - t0
    print(" evalues/evector accuracy: (compared to CPU results)")
    w2 = ws.w_h.numpy()[:n].copy()
    z2 = ws.Z_h.numpy().T[:, :m].copy()
    _report(w1, z1[:, il - 1:iu], w2, z2)
    print("\n Time for CUSTOM %s = %10.3f   (info = %d)" % ("zhegvdx_gpu" if cplx else "dsygvdx_gpu", t1 * 1e3, info))
    return w2, z2, info


def main(argv):
    if len(argv) < 2 or argv[0] not in ("z", "d"):
        print("Usage:\n\t python -m eigensolver_gpu_b200.test_driver z|d N [il iu]\n\t python -m eigensolver_gpu_b200.test_driver "
              "z|d fileA fileB [il iu]")
        return 2
    cplx = argv[0] == "z"
    rest = argv[1:]
    if rest[0].isdigit():
        print(" Using randomly-generated matrices...")
        n = int(rest[0])
        a, b = create_random_pd(n, cplx, 1), create_random_pd(n, cplx, 2)
        rest = rest[1:]
    else:
        print(" Reading  matrices from files ...\n Unformatted files with n,m,lda \n A(lda,n) B(lda,n)")
        a, n1, m1, lda1 = read_unformatted(rest[0], cplx)
        b, n2, m2, lda2 = read_unformatted(rest[1], cplx)
        if (n1, m1, lda1) != (n2, m2, lda2):
            print(" expecting A and B to have same N,M,LDA")
            return 1
        print(" n,m,lda from files: %d %d %d" % (n1, m1, lda1))
        n = n1
        rest = rest[2:]
        if not rest:
            rest = ["1", str(m1)]                                  # the dump's M = number of wanted eigenpairs
    il, iu = (int(rest[0]), int(rest[1])) if len(rest) >= 2 else (1, n)
    w, z, info = run(cplx, a, b, il, iu)
    return 0 if info == 0 else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
