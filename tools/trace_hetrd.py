"""Per-column timeline of the panel kernel (CTA 0): python tools/trace_hetrd.py N d|z tma [nb]"""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
from eigensolver_gpu_b200._lib import load
lib = load()
n = int(sys.argv[1]); cplx = sys.argv[2] == "z"; tma = int(sys.argv[3])
if len(sys.argv) > 4:
    lib.eigb200_set_option(b"trd_nb", int(sys.argv[4]))
lib.eigb200_set_option(b"symv_tma", tma)
dt = torch.complex128 if cplx else torch.float64
g = torch.randn((n, n), dtype=dt, device="cuda")
a0 = g + g.conj().T
a = a0.clone(); S.hetrd(a)
lib.eigb200_set_option(b"trd_trace", 1)
a = a0.clone(); S.hetrd(a); torch.cuda.synchronize()
NS = 16
buf = np.zeros(n * NS, dtype=np.uint64)
cnt = lib.eigb200_trace_read(buf.ctypes.data_as(C.c_void_p), n * NS)
t = buf.reshape(n, NS).astype(np.float64)
print(f"n={n} {'z' if cplx else 'd'} tma={tma}: columns traced {int((t[:,0]>0).sum())}")
es = 16 if cplx else 8
def med(sl, a, b):
    v = (t[sl, b] - t[sl, a]) / 1e3
    ok = (t[sl, a] > 0) & (t[sl, b] > 0)
    return np.median(v[ok]) if ok.any() else float("nan")
for lo in range(n - 512, -1, -1024):
    sl = slice(max(lo, 1), lo + 512)
    jm = (sl.start + sl.stop) / 2
    hb = es * jm * jm / 2
    print(f" cols {sl.start:5d}-{sl.stop:5d}: phaseA {med(sl,0,1):6.1f} us | wait1 {med(sl,1,2):5.1f} | phaseB(cta0) {med(sl,2,3):6.1f} | wait2 {med(sl,3,4):5.1f} | "
          f"column total {med(sl,0,4):6.1f} us | hemv-equivalent {hb/(med(sl,2,4)*1e3):.0f} B/ns")
    print(f"        A: loads+gather {med(sl,0,5):5.2f} | sync1 {med(sl,5,7):5.2f} | VW fma {med(sl,7,9):5.2f} | sync2 {med(sl,9,10):5.2f} | finish {med(sl,10,1):5.2f}"
          f"   B: larfg {med(sl,2,12):5.2f} | zdots {med(sl,12,13):5.2f} | engine {med(sl,13,3):6.2f}")
