"""Per-stage device timing of the full generalized solve (CUDA events), through the stage-level C ABI."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from eigensolver_gpu_b200 import api, stages as S  # noqa: E402


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def run(n, cplx, m):
    dt = torch.complex128 if cplx else torch.float64
    g = torch.randn((n, n), dtype=dt, device="cuda")
    a = (g + g.conj().T) / 2
    t = torch.rand((n, n), dtype=dt, device="cuda")
    b = t @ t.conj().T / n + torch.eye(n, dtype=dt, device="cuda")
    del g, t
    for rep in range(2):
        A, B = a.clone(), b.clone()
        torch.cuda.synchronize()
        marks = [("start", ev())]
        info = S.potrf(B); marks.append(("potrf", ev()))
        S.hegst(A, B); marks.append(("hegst", ev()))
        d, e, tau = S.hetrd(A); marks.append(("hetrd", ev()))
        w, q = S.stedc(d, e); marks.append(("stedc", ev()))
        z = q[:m].to(dt).contiguous(); marks.append(("select", ev()))
        S.ormtr(A, tau, z, m=m); marks.append(("ormtr", ev()))
        S.trsm("L", "N", B, z, m=n, n=m); marks.append(("trsm", ev()))
        torch.cuda.synchronize()
        if rep == 1:
            tot = marks[0][1].elapsed_time(marks[-1][1])
            parts = ", ".join(f"{marks[i][0]} {marks[i-1][1].elapsed_time(marks[i][1]):.1f}" for i in range(1, len(marks)))
            k = 4 if cplx else 1
            fl = k * (8 / 3 * n ** 3 + 3 * n * n * m)
            print(f"{'z' if cplx else 'd'} n={n} m={m}: total {tot:.1f} ms ({fl/tot*1e-9:.2f} TFLOP/s nominal) | {parts}", flush=True)
    # the real driver call
    ws = api.Workspace(n, cplx, host_z=False)
    for rep in range(2):
        A, B = a.clone(), b.clone()
        torch.cuda.synchronize()
        t0 = time.time()
        info, w, z, _ = api.solve_generalized(A, B, 1, m, ws=ws, skip_host_copy=True)
        torch.cuda.synchronize()
        t1 = time.time() - t0
    print(f"  driver call: info={info} {t1*1e3:.1f} ms", flush=True)


if __name__ == "__main__":
    run(4096, False, 512)
    run(4096, True, 4096)
    run(8192, True, 8192)
