import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
from eigensolver_gpu_b200._lib import load
lib = load()
n = int(sys.argv[1]); cplx = sys.argv[2] == "z"; tma = int(sys.argv[3])
lib.eigb200_set_option(b"symv_tma", tma)
dt = torch.complex128 if cplx else torch.float64
g = torch.randn((n, n), dtype=dt, device="cuda")
a0 = g + g.conj().T
a = a0.clone(); S.hetrd(a)
lib.eigb200_set_option(b"trd_trace", 1)
a = a0.clone(); S.hetrd(a); torch.cuda.synchronize()
buf = np.zeros(n * 5, dtype=np.uint64)
cnt = lib.eigb200_trace_read(buf.ctypes.data_as(C.c_void_p), n * 5)
t = buf.reshape(n, 5).astype(np.float64)
dA = (t[:, 1] - t[:, 0]) / 1e3; b1 = (t[:, 2] - t[:, 1]) / 1e3; dB = (t[:, 3] - t[:, 2]) / 1e3; b2 = (t[:, 4] - t[:, 3]) / 1e3
print(f"n={n} {'z' if cplx else 'd'} tma={tma}: columns traced {int((t[:,0]>0).sum())}")
es = 16 if cplx else 8
for lo in range(n - 512, -1, -1024):
    sl = slice(max(lo, 1), lo + 512)
    jm = (sl.start + sl.stop) / 2
    hb = es * jm * jm / 2
    print(f" cols {sl.start:5d}-{sl.stop:5d}: phaseA {np.median(dA[sl]):6.1f} us | wait1 {np.median(b1[sl]):5.1f} | phaseB(cta0) {np.median(dB[sl]):6.1f} | wait2 {np.median(b2[sl]):5.1f} | "
          f"column total {np.median((t[sl,4]-t[sl,0]))/1e3:6.1f} us | hemv-equivalent {hb/np.median((t[sl,4]-t[sl,2]))*1e-0/1e0:.0f} B/ns")
