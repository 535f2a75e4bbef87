"""Large real shapes: TMA-fed GEMM vs cp.async GEMM vs torch (which product goes wrong beyond n = 8192?)."""
import sys
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
from eigensolver_gpu_b200._lib import load
lib = load()
dt = torch.float64
def opt(v):
    assert lib.eigb200_set_option(b"gemm_tma", v) == 0
def chk(name, fn, ref):
    opt(1); a = fn(); opt(0); b = fn()
    r = ref() if ref is not None else b
    den = float(r.abs().max()) + 1e-300
    print(f"{name}: tma-vs-ref {float((a - r).abs().max())/den:.3g}  cpasync-vs-ref {float((b - r).abs().max())/den:.3g}", flush=True)
def gemm_case(ta, tb, m, n, k, ldpad=0):
    a = torch.randn((k, m) if ta == "N" else (m, k), dtype=dt, device="cuda")
    b = torch.randn((n, k) if tb == "N" else (k, n), dtype=dt, device="cuda")
    def fn():
        c = torch.zeros((n, m), dtype=dt, device="cuda")
        S.gemm(ta, tb, 1.0, a, b, 0.0, c, m=m, n=n, k=k); return c
    def ref():
        A = a.T if ta == "N" else a      # tensor (cols, rows): matrix = tensor.T
        B = b.T if tb == "N" else b
        return (A @ B).T.contiguous()
    chk(f"gemm {ta}{tb} {m}x{n}x{k}", fn, ref)
for shp in [("N", "N", 12288, 12288, 64), ("N", "N", 16384, 4096, 128), ("C", "N", 128, 2048, 12288), ("N", "N", 6144, 6144, 6144),
            ("N", "N", 2048, 10240, 2048), ("C", "N", 10240, 2048, 2048), ("N", "C", 12288, 2048, 64)]:
    gemm_case(*shp)
# her2k on a big real matrix
n, k = 12288, 64
a = torch.randn((k, n), dtype=dt, device="cuda"); b = torch.randn((k, n), dtype=dt, device="cuda")
def fn():
    c = torch.zeros((n, n), dtype=dt, device="cuda"); S.her2k(-1.0, a, b, 1.0, c); return torch.tril(c)   # tensor lower = matrix upper
chk("her2k n=12288 k=64", fn, lambda: torch.tril(-(b.T @ a + a.T @ b)))
# stages at n = 12288: potrf, hegst, trsm (TMA vs cp.async)
n = 12288
t = torch.rand((n, n), dtype=dt, device="cuda"); bm = t @ t.T / n + torch.eye(n, dtype=dt, device="cuda")
g = torch.randn((n, n), dtype=dt, device="cuda"); am = (g + g.T) / 2
del t, g
def potrf():
    B = bm.clone(); S.potrf(B); return torch.tril(B)
chk("potrf n=12288", potrf, None)
opt(0); U = bm.clone(); S.potrf(U)
def hegst():
    A = am.clone(); S.hegst(A, U); return torch.tril(A)
chk("hegst n=12288", hegst, None)
def trsm():
    X = am[:2048].clone(); S.trsm("L", "N", U, X, m=n, n=2048); return X
chk("trsm LN n=12288 x 2048", trsm, None)
def ormtr():
    A = am.clone(); d, e, tau = S.hetrd(A); z = torch.eye(n, dtype=dt, device="cuda")[:1536].contiguous(); S.ormtr(A, tau, z, m=1536); return z
chk("hetrd+ormtr n=12288 m=1536", ormtr, None)
