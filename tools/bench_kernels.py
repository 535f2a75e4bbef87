"""Device-side timing of individual kernels through the C ABI (CUDA events, warm-up, L2-sized inputs)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S  # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_gemm():
    for cplx in (False, True):
        dt = torch.complex128 if cplx else torch.float64
        fl = 8 if cplx else 2
        for (ta, tb, m, n, k) in [("N", "N", 8192, 8192, 8192), ("N", "N", 4096, 4096, 128), ("N", "C", 8192, 8192, 64),
                                  ("C", "N", 128, 8192, 8192), ("N", "N", 8192, 8192, 128), ("N", "N", 4096, 4096, 4096)]:
            if cplx and m * n * k > 4096 ** 3 * 2:
                m //= 2; n //= 2
            a = torch.randn((k, m) if ta == "N" else (m, k), dtype=dt, device="cuda")
            b = torch.randn((n, k) if tb == "N" else (k, n), dtype=dt, device="cuda")
            c = torch.zeros((n, m), dtype=dt, device="cuda")
            ms = timeit(lambda: S.gemm(ta, tb, 1.0, a, b, 0.0, c, m=m, n=n, k=k))
            print(f"gemm {'z' if cplx else 'd'} {ta}{tb} {m}x{n}x{k}: {ms:.3f} ms  {fl*m*n*k/ms*1e-9:.2f} TFLOP/s", flush=True)
        for (n, k) in [(8192, 64), (8192, 32), (4096, 64)]:
            a = torch.randn((k, n), dtype=dt, device="cuda")
            b = torch.randn((k, n), dtype=dt, device="cuda")
            c = torch.zeros((n, n), dtype=dt, device="cuda")
            ms = timeit(lambda: S.her2k(-1.0, a, b, 1.0, c))
            byt = n * n * (16 if cplx else 8)
            print(f"her2k {'z' if cplx else 'd'} n={n} k={k}: {ms:.3f} ms  {fl*n*n*k/ms*1e-9:.2f} TFLOP/s  {byt/ms*1e-6:.0f} GB/s(R+W tri)", flush=True)




def bench_hemv():
    for cplx in (False, True):
        dt = torch.complex128 if cplx else torch.float64
        es = 16 if cplx else 8
        for n in (2048, 4096, 8192, 16384):
            a = torch.randn((n, n), dtype=dt, device="cuda")
            x = torch.randn(n, dtype=dt, device="cuda")
            ms = timeit(lambda: S.hemv(a, x), reps=20)
            byt = es * (n * (n + 1) / 2 + 3 * n)
            print(f"hemv {'z' if cplx else 'd'} n={n}: {ms*1e3:.1f} us  {byt/ms*1e-6:.0f} GB/s", flush=True)


def bench_hetrd():
    from eigensolver_gpu_b200._lib import load
    lib = load()
    for cplx, n in ((False, 2048), (False, 4096), (True, 4096), (True, 8192), (False, 8192)):
        dt = torch.complex128 if cplx else torch.float64
        g = torch.randn((n, n), dtype=dt, device="cuda")
        a0 = g + g.conj().T
        for nb, coop in ((64, 1), (32, 1), (64, 0)):
            lib.eigb200_set_option(b"trd_nb", nb)
            lib.eigb200_set_option(b"trd_coop", coop)
            a = a0.clone()
            S.hetrd(a)
            a = a0.clone()
            torch.cuda.synchronize()
            t0 = time.time()
            S.hetrd(a)
            torch.cuda.synchronize()
            dt_s = time.time() - t0
            k = 4 if cplx else 1
            print(f"hetrd {'z' if cplx else 'd'} n={n} nb={nb} coop={coop}: {dt_s*1e3:.1f} ms  {k*4/3*n**3/dt_s*1e-12:.2f} TFLOP/s "
                  f"(hemv bytes {(16 if cplx else 8)*n**3/6/dt_s*1e-9:.0f} GB/s-equiv)", flush=True)
        lib.eigb200_set_option(b"trd_nb", 64)
        lib.eigb200_set_option(b"trd_coop", 1)


if __name__ == "__main__":
    from eigensolver_gpu_b200._lib import load
    which = [a for a in sys.argv[1:] if "=" not in a] or ["gemm"]
    for kv in [a for a in sys.argv[1:] if "=" in a]:          # name=value library options, e.g. gemm_tma=0
        k, v = kv.split("=")
        assert load().eigb200_set_option(k.encode(), int(v)) == 0, kv
        print(f"option {k} = {v}", flush=True)
    for w in which:
        globals()["bench_" + w]()
