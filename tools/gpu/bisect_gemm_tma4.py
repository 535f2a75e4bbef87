import sys
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
from eigensolver_gpu_b200._lib import load
lib = load()
dt = torch.float64
def setopt(k, v):
    assert lib.eigb200_set_option(k.encode(), v) == 0
n = 10240
t = torch.rand((n, n), dtype=dt, device="cuda"); bm = t @ t.T / n + torch.eye(n, dtype=dt, device="cuda")
del t
setopt("gemm_tma", 0); U = bm.clone(); S.potrf(U); del bm
g = torch.randn((n, n), dtype=dt, device="cuda"); am = (g + g.T) / 2
del g
setopt("hegst_hb", 2048)
ref = torch.tril(S.hegst(am.clone(), U))
for dbg in (0, 1, 2, 3, 0):
    setopt("gemm_tma", 1); setopt("gemm_tma_dbg", dbg)
    res = []
    for rep in range(3):
        a = torch.tril(S.hegst(am.clone(), U))
        d = (a - ref).abs()
        res.append((float(d.max()), int((d > 0).sum())))
    print(f"hegst n={n} hb=2048 dbg={dbg}: (max diff, wrong elements) per run {res}", flush=True)
# where are the wrong elements? (last run, dbg 0)
d = (a - ref).abs()
bad = (d > 0).nonzero()
if bad.numel():
    cols, rows = bad[:, 0], bad[:, 1]
    print("wrong elements: rows", int(rows.min()), "..", int(rows.max()), " cols", int(cols.min()), "..", int(cols.max()), " count", bad.shape[0])
    # histogram by 2048-blocks
    import collections
    h = collections.Counter((int(r) // 2048, int(c) // 2048) for c, r in bad[:: max(1, bad.shape[0] // 2000)].tolist())
    print("block (row, col) histogram of a sample:", sorted(h.items()))
