import sys
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
from eigensolver_gpu_b200._lib import load
n = int(sys.argv[1]); cplx = sys.argv[2] == "z"; coop = int(sys.argv[3]); nb = int(sys.argv[4]) if len(sys.argv) > 4 else 64
lib = load()
lib.eigb200_set_option(b"trd_coop", coop)
lib.eigb200_set_option(b"trd_nb", nb)
dt = torch.complex128 if cplx else torch.float64
g = torch.randn((n, n), dtype=dt, device="cuda")
a = g + g.conj().T
torch.cuda.synchronize()
S.hetrd(a)
torch.cuda.synchronize()
print("done")
