"""A/B timing of one stage under option settings: python tools/bench_stage_opts.py potrf|hegst|trsm|ormtr|stedc N d|z name=value ..."""
import sys
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
from eigensolver_gpu_b200._lib import load
lib = load()
stage, n, cplx = sys.argv[1], int(sys.argv[2]), sys.argv[3] == "z"
dt = torch.complex128 if cplx else torch.float64
t = torch.rand((n, n), dtype=dt, device="cuda")
b = t @ t.conj().T / n + torch.eye(n, dtype=dt, device="cuda")
g = torch.randn((n, n), dtype=dt, device="cuda"); a = (g + g.conj().T) / 2
del g, t
for setting in sys.argv[4:]:
    saved = {}
    for kv in setting.split(","):
        k, v = kv.split("=")
        saved[k] = lib.eigb200_get_option(k.encode())
        assert lib.eigb200_set_option(k.encode(), int(v)) == 0, k
    best = 1e9
    for rep in range(3):
        B = b.clone(); A = a.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if stage == "potrf":
            e0.record(); S.potrf(B); e1.record()
        elif stage == "hegst":
            S.potrf(B); e0.record(); S.hegst(A, B); e1.record()
        elif stage == "trsm":
            S.potrf(B); e0.record(); S.trsm("L", "N", B, A, m=n, n=n); e1.record()
        elif stage == "ormtr":
            d, e, tau = S.hetrd(A) if rep == 0 else (None, None, tau)
            if rep == 0:
                Ared = A.clone()
            z = torch.eye(n, dtype=dt, device="cuda")
            e0.record(); S.ormtr(Ared, tau, z, m=n); e1.record()
        elif stage == "stedc":
            dd = torch.randn(n, dtype=torch.float64, device="cuda"); ee = torch.randn(n - 1, dtype=torch.float64, device="cuda")
            e0.record(); S.stedc(dd, ee); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{stage} {'z' if cplx else 'd'} n={n} [{setting}]: {best:.2f} ms", flush=True)
    for k, v in saved.items():
        lib.eigb200_set_option(k.encode(), v)
