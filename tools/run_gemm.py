"""One launch each of a few representative GEMM shapes (for ncu): python tools/run_gemm.py"""
import sys
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
for cplx in (True, False):
    dt = torch.complex128 if cplx else torch.float64
    m = n = k = 4096
    a = torch.randn((k, m), dtype=dt, device="cuda"); b = torch.randn((n, k), dtype=dt, device="cuda")
    c = torch.zeros((n, m), dtype=dt, device="cuda")
    S.gemm("N", "N", 1.0, a, b, 1.0, c, m=m, n=n, k=k)
    n, k = 8192, 64
    a = torch.randn((k, n), dtype=dt, device="cuda"); b = torch.randn((k, n), dtype=dt, device="cuda")
    c = torch.zeros((n, n), dtype=dt, device="cuda")
    S.her2k(-1.0, a, b, 1.0, c)
torch.cuda.synchronize()
