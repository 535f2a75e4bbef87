"""A/B timing of the tridiagonalization under option settings:
python tools/bench_hetrd_opts.py N d|z name=value[,name=value...] [more settings ...]"""
import sys, time
import torch
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S
from eigensolver_gpu_b200._lib import load
lib = load()
n = int(sys.argv[1]); cplx = sys.argv[2] == "z"
dt = torch.complex128 if cplx else torch.float64
g = torch.randn((n, n), dtype=dt, device="cuda")
a0 = g + g.conj().T
defaults = {}
for setting in sys.argv[3:]:
    kv = [x.split("=") for x in setting.split(",") if x]
    for k, v in kv:
        if k not in defaults:
            defaults[k] = lib.eigb200_get_option(k.encode())
        assert lib.eigb200_set_option(k.encode(), int(v)) == 0, k
    best = 1e9
    for rep in range(3):
        a = a0.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); S.hetrd(a); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"hetrd {'z' if cplx else 'd'} n={n} [{setting}]: {best:.1f} ms", flush=True)
    for k, v in defaults.items():
        lib.eigb200_set_option(k.encode(), v)
