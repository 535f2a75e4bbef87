"""cuSOLVER comparator (REPORTED BASELINE ONLY -- never on the product path).

The reference's own test driver times `cusolverDnZhegvdx` / `cusolverDnDsygvdx` next to its custom solver
(test_driver/test_zhegvdx.F90:213-263, test_dsygvdx.F90:240-290).  The reference itself cannot be built here
(CUDA Fortran), so this library call on the same B200 stands in for "the reference's algorithm re-timed through
cuSOLVER/cuBLAS" (SURVEY.md section 2.2).  Bound through ctypes; torch only provides device memory.
"""
import ctypes as C
import glob
import os

_EIG_TYPE_1 = 1            # cusolverEigType_t
_EIG_MODE_VECTOR = 1       # cusolverEigMode_t
_EIG_RANGE_ALL = 1001      # cusolverEigRange_t
_EIG_RANGE_I = 1002
_FILL_MODE_UPPER = 1       # cublasFillMode_t


def _load():
    import torch
    cands = []
    base = os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "cusolver", "lib")
    cands += sorted(glob.glob(os.path.join(base, "libcusolver.so*")))
    cands += ["libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so"]
    last = None
    for c in cands:
        try:
            return C.CDLL(c, mode=C.RTLD_GLOBAL)
        except OSError as e:
            last = e
    raise RuntimeError(f"cusolver baseline: libcusolver not loadable ({last})")


class CusolverHegvdx:
    """Pre-sized cusolverDn?{sy,he}gvdx plan for order n, eigenpairs il..iu (1-based), UPLO='U', ITYPE=1, JOBZ='V'."""

    def __init__(self, n, il, iu, cplx):
        import torch
        self.torch = torch
        self.lib = _load()
        self.n, self.il, self.iu, self.cplx = n, il, iu, cplx
        self.h = C.c_void_p()
        st = self.lib.cusolverDnCreate(C.byref(self.h))
        if st != 0:
            raise RuntimeError(f"cusolverDnCreate -> {st}")
        self.lib.cusolverDnSetStream(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self.range = _EIG_RANGE_ALL if (il == 1 and iu == n) else _EIG_RANGE_I
        self.w = torch.empty(n, dtype=torch.float64, device="cuda")
        self.info = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.meig = C.c_int(0)
        self.fn = self.lib.cusolverDnZhegvdx if cplx else self.lib.cusolverDnDsygvdx
        self.fnb = self.lib.cusolverDnZhegvdx_bufferSize if cplx else self.lib.cusolverDnDsygvdx_bufferSize
        self.lwork = C.c_int(0)
        self.work = None

    def _common(self, a, b):
        n = self.n
        return [self.h, C.c_int(_EIG_TYPE_1), C.c_int(_EIG_MODE_VECTOR), C.c_int(self.range), C.c_int(_FILL_MODE_UPPER),
                C.c_int(n), C.c_void_p(a.data_ptr()), C.c_int(n), C.c_void_p(b.data_ptr()), C.c_int(n),
                C.c_double(0.0), C.c_double(0.0), C.c_int(self.il), C.c_int(self.iu), C.byref(self.meig),
                C.c_void_p(self.w.data_ptr())]

    def plan(self, a, b):
        st = self.fnb(*self._common(a, b), C.byref(self.lwork))
        if st != 0:
            raise RuntimeError(f"cusolverDn?gvdx_bufferSize -> {st}")
        dt = self.torch.complex128 if self.cplx else self.torch.float64
        self.work = self.torch.empty(max(self.lwork.value, 1), dtype=dt, device="cuda")

    def solve(self, a, b):
        """a, b: column-major device tensors (upper triangles read); a is overwritten by the eigenvectors, b by its factor."""
        if self.work is None:
            self.plan(a, b)
        st = self.fn(*self._common(a, b), C.c_void_p(self.work.data_ptr()), self.lwork, C.c_void_p(self.info.data_ptr()))
        if st != 0:
            raise RuntimeError(f"cusolverDn?gvdx -> {st}")
        return self.w

    def close(self):
        if self.h:
            self.lib.cusolverDnDestroy(self.h)
            self.h = C.c_void_p()
