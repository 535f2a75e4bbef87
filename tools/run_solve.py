"""One generalized solve after a warm-up solve (for an ncu launch list; profiler range = the second solve):
python tools/run_solve.py N d|z [m]"""
import sys
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import api
n, cplx = int(sys.argv[1]), sys.argv[2] == "z"
m = int(sys.argv[3]) if len(sys.argv) > 3 else n
dt = torch.complex128 if cplx else torch.float64
gen = torch.Generator(device="cuda").manual_seed(1234)
def herm():
    t = torch.rand((n, n), dtype=torch.float64, device="cuda", generator=gen)
    if cplx:
        t = torch.complex(t, torch.rand((n, n), dtype=torch.float64, device="cuda", generator=gen))
    t = torch.triu(t) + torch.triu(t, 1).conj().T
    x = t @ t.conj().T
    return ((x + x.conj().T) / 2).contiguous()
a0, b0 = herm(), herm()
ws = api.Workspace(n, cplx, host_z=False)
for rep in range(2):
    A, B = a0.clone(), b0.clone()
    torch.cuda.synchronize()
    if rep == 1:
        torch.cuda.profiler.start()
    info, w, z, _ = api.solve_generalized(A, B, 1, m, ws=ws, skip_host_copy=True)
    torch.cuda.synchronize()
    if rep == 1:
        torch.cuda.profiler.stop()
    assert info == 0
print("done")
