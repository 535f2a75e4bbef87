"""One complex NN GEMM 4096^3 (for ncu): python tools/run_gemm1.py"""
import sys
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
dt = torch.complex128
m = n = k = 4096
a = torch.randn((k, m), dtype=dt, device="cuda"); b = torch.randn((n, k), dtype=dt, device="cuda")
c = torch.zeros((n, m), dtype=dt, device="cuda")
S.gemm("N", "N", 1.0, a, b, 1.0, c, m=m, n=n, k=k)
S.gemm("N", "N", 1.0, a, b, 1.0, c, m=m, n=n, k=k)
torch.cuda.synchronize()
