import sys, ctypes as C
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200._lib import load
lib = load(); lib.eigb200_init()
dt = torch.float64
n, kb = 10240, 2048
k = 0
r = n - k - kb
A0 = torch.randn((n, n), dtype=dt, device="cuda")
U0 = torch.randn((n, n), dtype=dt, device="cuda")
def ptr(t, row, col):
    return C.c_void_p(t.data_ptr() + 8 * (row + col * n))
def run(tma, fn):
    lib.eigb200_set_option(b"gemm_tma", tma)
    A = A0.clone(); fn(A); torch.cuda.synchronize(); return A
def cmp(name, fn):
    a, b = run(1, fn), run(0, fn)
    d = (a - b).abs(); idx = int(d.argmax())
    print(f"{name}: max diff {float(d.max()):.3g} (ref max {float(b.abs().max()):.3g}) at col {idx // n} row {idx % n}", flush=True)
ch = lambda c: C.c_char(c.encode())
# gemm -1/2: Akr -= 0.5 Akk Ukr    (kb x r) = (kb x kb)(kb x r)
cmp("gemm NN Akr -= .5 Akk Ukr", lambda A: lib.eigb200_dgemm(ch("N"), ch("N"), kb, r, kb, -0.5, ptr(A, k, k), n, ptr(U0, k, k + kb), n, 1.0, ptr(A, k, k + kb), n))
# rank-2k pieces as plain products: Arr -= Akr^H Ukr   (r x r) = (kb x r)^H (kb x r)
cmp("gemm CN Arr -= Akr^H Ukr", lambda A: lib.eigb200_dgemm(ch("C"), ch("N"), r, r, kb, -1.0, ptr(A, k, k + kb), n, ptr(U0, k, k + kb), n, 1.0, ptr(A, k + kb, k + kb), n))
cmp("gemm CN Arr -= Ukr^H Akr", lambda A: lib.eigb200_dgemm(ch("C"), ch("N"), r, r, kb, -1.0, ptr(U0, k, k + kb), n, ptr(A, k, k + kb), n, 1.0, ptr(A, k + kb, k + kb), n))
for rr in (6144, 7168, 8192):
    cmp(f"gemm CN r={rr}", lambda A: lib.eigb200_dgemm(ch("C"), ch("N"), rr, rr, kb, -1.0, ptr(A, k, k + kb), n, ptr(U0, k, k + kb), n, 1.0, ptr(A, k + kb, k + kb), n))
for kk in (512, 1024, 1536):
    cmp(f"gemm CN r=8192 K={kk}", lambda A: lib.eigb200_dgemm(ch("C"), ch("N"), r, r, kk, -1.0, ptr(A, k, k + kb), n, ptr(U0, k, k + kb), n, 1.0, ptr(A, k + kb, k + kb), n))
