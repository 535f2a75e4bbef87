import sys
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
n = int(sys.argv[1]); cplx = sys.argv[2] == "z"
dt = torch.complex128 if cplx else torch.float64
a = torch.randn((n, n), dtype=dt, device="cuda")
x = torch.randn(n, dtype=dt, device="cuda")
for _ in range(4):
    y = S.hemv(a, x)
torch.cuda.synchronize()
print("done")
