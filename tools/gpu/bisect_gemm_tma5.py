import sys
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
from eigensolver_gpu_b200._lib import load
lib = load()
dt = torch.float64
def setopt(k, v):
    assert lib.eigb200_set_option(k.encode(), v) == 0
order = [int(c) for c in sys.argv[1]]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10240
t = torch.rand((n, n), dtype=dt, device="cuda"); bm = t @ t.T / n + torch.eye(n, dtype=dt, device="cuda")
del t
setopt("gemm_tma", order[0]); U = bm.clone(); S.potrf(U); del bm
g = torch.randn((n, n), dtype=dt, device="cuda"); am = (g + g.T) / 2
del g
setopt("hegst_hb", 2048)
outs = []
for o in order:
    setopt("gemm_tma", o)
    outs.append(torch.tril(S.hegst(am.clone(), U)))
for i, o in enumerate(order):
    d = (outs[i] - outs[-1]).abs()
    print(f"order {sys.argv[1]} n={n}: run {i} (tma={o}) vs last: max diff {float(d.max()):.3g}, wrong {int((d > 0).sum())}", flush=True)
