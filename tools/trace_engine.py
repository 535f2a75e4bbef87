"""Per-tile timeline of the tile engine in one CTA for the product of order J: python tools/trace_engine.py N d|z J [cta]"""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
from eigensolver_gpu_b200._lib import load
lib = load()
n = int(sys.argv[1]); cplx = sys.argv[2] == "z"; J = int(sys.argv[3]); cta = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dt = torch.complex128 if cplx else torch.float64
g = torch.randn((n, n), dtype=dt, device="cuda")
a0 = g + g.conj().T
a = a0.clone(); S.hetrd(a)
lib.eigb200_set_option(b"trd_trace", J); lib.eigb200_set_option(b"trd_trace_cta", cta)
a = a0.clone(); S.hetrd(a); torch.cuda.synchronize()
NS = 16
buf = np.zeros(n * NS + 512, dtype=np.uint64)
lib.eigb200_trace_read(buf.ctypes.data_as(C.c_void_p), n * NS + 512)
e = buf[n * NS:n * NS + 480].reshape(-1, 8).astype(np.int64)
t0 = e[0, 0]
print(f"n={n} {'z' if cplx else 'd'} order {J} cta {cta}: tile: issue | full(wait done) | +loaded | +fma | +reduce | +stored | +done  [cycles], flags")
for i in range(len(e)):
    if e[i, 0] == 0 and e[i, 1] == 0: break
    print(f"  {i:3d}: {e[i,0]-t0:8d} | {e[i,1]-t0:8d} | {e[i,2]-e[i,1]:6d} {e[i,3]-e[i,2]:6d} {e[i,4]-e[i,3]:6d} {e[i,5]-e[i,4]:6d} {e[i,6]-e[i,5]:6d}  fl={e[i,7]}")
