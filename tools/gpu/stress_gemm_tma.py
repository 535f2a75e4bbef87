import sys, ctypes as C
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200._lib import load
lib = load(); lib.eigb200_init()
dt = torch.float64
n, kb = 10240, 2048
A0 = torch.randn((n, n), dtype=dt, device="cuda")
U0 = torch.randn((n, n), dtype=dt, device="cuda")
ptr = lambda t, row, col: C.c_void_p(t.data_ptr() + 8 * (row + col * n))
ch = lambda c: C.c_char(c.encode())
def prod(A, ta, rr, kk):
    if ta == "C":
        lib.eigb200_dgemm(ch("C"), ch("N"), rr, rr, kk, -1.0, ptr(A, 0, kb), n, ptr(U0, 0, kb), n, 1.0, ptr(A, kb, kb), n)
    else:
        lib.eigb200_dgemm(ch("N"), ch("N"), rr, rr, kk, -1.0, ptr(A, kb, 0), n, ptr(U0, 0, kb), n, 1.0, ptr(A, kb, kb), n)
for ta, rr, kk in (("C", 7168, 2048), ("N", 7168, 2048), ("C", 4096, 4096)):
    lib.eigb200_set_option(b"gemm_tma", 0)
    ref = A0.clone(); prod(ref, ta, rr, kk); torch.cuda.synchronize()
    for dbg in (0, 1, 2, 3):
        lib.eigb200_set_option(b"gemm_tma", 1); lib.eigb200_set_option(b"gemm_tma_dbg", dbg)
        bad = 0; worst = 0.0; nbad = 0
        for rep in range(12):
            A = A0.clone(); prod(A, ta, rr, kk); torch.cuda.synchronize()
            d = (A - ref).abs()
            m = float(d.max())
            if m > 0:
                bad += 1; worst = max(worst, m); nbad = int((d > 0).sum())
        print(f"{ta}N r={rr} K={kk} dbg={dbg}: {bad}/12 runs differ, worst {worst:.3g}, wrong elements in last bad run {nbad}", flush=True)
