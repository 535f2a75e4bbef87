import sys, time
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
from eigensolver_gpu_b200._lib import load
lib = load()
for cplx, n in ((True, 8192), (False, 8192), (False, 4096)):
    dt = torch.complex128 if cplx else torch.float64
    g = torch.randn((n, n), dtype=dt, device="cuda")
    a0 = g + g.conj().T
    for tma in (1, 0):
        lib.eigb200_set_option(b"symv_tma", tma)
        for rep in range(2):
            a = a0.clone(); torch.cuda.synchronize(); t0 = time.time(); S.hetrd(a); torch.cuda.synchronize(); t = time.time() - t0
        print(f"hetrd {'z' if cplx else 'd'} n={n} tma={tma}: {t*1e3:.1f} ms", flush=True)
