import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import lapack, matgen, metrics
from eigensolver_gpu_b200 import stages as S

for kind in ["random", "toeplitz", "wilkinson", "glued", "graded", "zero_e"]:
    for n in [3, 8, 31, 32, 33, 64, 100, 257, 1000]:
        d, e = matgen.tridiag_family(n, kind, seed=n)
        w, q = S.stedc(S.to_dev(d), S.to_dev(e))
        w, q = S.to_host(w), np.array(S.to_host(q))
        wr, zr, info = lapack.stedc(d, e)
        t = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        tn = np.abs(t).sum(axis=0).max()
        fin = np.all(np.isfinite(w)) and np.all(np.isfinite(q))
        err = np.abs(np.sort(w) - wr).max() / (n * metrics.EPS * tn) if fin else np.inf
        g = metrics.std_gates(t, w, q) if fin else {}
        print(f"{kind:10s} n={n:5d} finite={fin} sorted={bool(np.all(np.diff(w) >= 0))} dw/gate={err:.3g} {g}", flush=True)
