"""Per-CTA begin/end of phase B (tile phase) for the product of order J: python tools/trace_spread.py N d|z J"""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
from eigensolver_gpu_b200._lib import load
lib = load()
n = int(sys.argv[1]); cplx = sys.argv[2] == "z"; J = int(sys.argv[3])
dt = torch.complex128 if cplx else torch.float64
g = torch.randn((n, n), dtype=dt, device="cuda")
a0 = g + g.conj().T
a = a0.clone(); S.hetrd(a)
lib.eigb200_set_option(b"trd_trace", J)
a = a0.clone(); S.hetrd(a); torch.cuda.synchronize()
NS = 16
buf = np.zeros(n * NS + 1024, dtype=np.uint64)
lib.eigb200_trace_read(buf.ctypes.data_as(C.c_void_p), n * NS + 1024)
e = buf[n * NS:].astype(np.float64)
b, f = e[512:512 + 148], e[768:768 + 148]
t0 = b.min()
b = (b - t0) / 1e3; f = (f - t0) / 1e3
print(f"n={n} {'z' if cplx else 'd'} order {J}: phase B begin spread {b.max()-b.min():.2f} us; end: min {f.min():.2f} median {np.median(f):.2f} "
      f"p90 {np.percentile(f,90):.2f} max {f.max():.2f} us  (CTA0 end {f[0]:.2f})")
order = np.argsort(f)
print(" last 8 CTAs to finish:", [(int(i), round(float(f[i]), 2)) for i in order[-8:]])
print(" first 8 CTAs to finish:", [(int(i), round(float(f[i]), 2)) for i in order[:8]])
