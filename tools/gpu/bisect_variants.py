import sys
import torch
sys.path.insert(0, ".")
import bench
from eigensolver_gpu_b200 import api
from eigensolver_gpu_b200._lib import load
lib = load()
n = int(sys.argv[1])
for rep in range(2):
    a0, b0 = bench.make_inputs(torch, n, False, "C", 99)
    A, B = a0.clone(), b0.clone()
    info, w, z, ws = api.solve_generalized(A, B, 1, n // 8, skip_host_copy=True)
    g = bench.parity_metrics(torch, a0, b0, w, z, n // 8)
    print(f"{sys.argv[2]} n={n}: residual_max={g['residual_max']:.3g}", flush=True)
    del a0, b0, A, B, ws
