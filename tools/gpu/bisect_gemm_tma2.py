import sys
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
from eigensolver_gpu_b200._lib import load
lib = load()
dt = torch.float64
def opt(v):
    assert lib.eigb200_set_option(b"gemm_tma", v) == 0
def chk(name, fn):
    opt(1); a = fn(); opt(0); b = fn()
    den = float(b.abs().max()) + 1e-300
    d = (a - b).abs()
    idx = int(d.argmax())
    print(f"{name}: tma-vs-cpasync {float(d.max())/den:.3g} at flat index {idx} (col {idx // a.shape[1]}, row {idx % a.shape[1]})", flush=True)
for n in (10240, 12288):
    t = torch.rand((n, n), dtype=dt, device="cuda"); bm = t @ t.T / n + torch.eye(n, dtype=dt, device="cuda")
    del t
    opt(0); U = bm.clone(); S.potrf(U); del bm
    X = torch.randn((n, 2048), dtype=dt, device="cuda")          # matrix 2048 x n (rows x cols): tensor (n, 2048)
    chk(f"trsm R N 2048 x {n}", lambda: S.trsm("R", "N", U, X.clone(), m=2048, n=n))
    Y = torch.randn((2048, n), dtype=dt, device="cuda")          # matrix n x 2048
    chk(f"trsm L C {n} x 2048", lambda: S.trsm("L", "C", U, Y.clone(), m=n, n=2048))
    chk(f"trsm L N {n} x 2048", lambda: S.trsm("L", "N", U, Y.clone(), m=n, n=2048))
    for hb in (1024, 2048):
        lib.eigb200_set_option(b"hegst_hb", hb)
        g = torch.randn((n, n), dtype=dt, device="cuda"); am = (g + g.T) / 2
        del g
        chk(f"hegst n={n} hb={hb}", lambda: torch.tril(S.hegst(am.clone(), U)))
        del am
    lib.eigb200_set_option(b"hegst_hb", 0)
    del U, X, Y
    torch.cuda.empty_cache()
