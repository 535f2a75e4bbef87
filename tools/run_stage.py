"""Runs one stage once (after a warm-up call) for an ncu launch list: python tools/run_stage.py potrf|hegst|ormtr|stedc N d|z [m]"""
import sys
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
stage, n, cplx = sys.argv[1], int(sys.argv[2]), sys.argv[3] == "z"
m = int(sys.argv[4]) if len(sys.argv) > 4 else n
dt = torch.complex128 if cplx else torch.float64
t = torch.rand((n, n), dtype=dt, device="cuda")
b = t @ t.conj().T / n + torch.eye(n, dtype=dt, device="cuda")
g = torch.randn((n, n), dtype=dt, device="cuda"); a = (g + g.conj().T) / 2
del g, t
reps = 2
for rep in range(reps):
    B = b.clone(); A = a.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if stage == "potrf":
        torch.cuda.profiler.start() if rep == reps - 1 else None
        e0.record(); S.potrf(B); e1.record()
    elif stage == "hegst":
        S.potrf(B)
        torch.cuda.profiler.start() if rep == reps - 1 else None
        e0.record(); S.hegst(A, B); e1.record()
    elif stage == "ormtr":
        d, e, tau = S.hetrd(A)
        z = torch.eye(n, dtype=dt, device="cuda")[:m].contiguous()
        torch.cuda.profiler.start() if rep == reps - 1 else None
        e0.record(); S.ormtr(A, tau, z, m=m); e1.record()
    elif stage == "stedc":
        d = torch.randn(n, dtype=torch.float64, device="cuda"); e = torch.randn(n - 1, dtype=torch.float64, device="cuda")
        torch.cuda.profiler.start() if rep == reps - 1 else None
        e0.record(); S.stedc(d, e); e1.record()
    torch.cuda.synchronize()
    if rep == reps - 1:
        torch.cuda.profiler.stop()
    print(f"{stage} {'z' if cplx else 'd'} n={n}: {e0.elapsed_time(e1):.2f} ms", flush=True)
