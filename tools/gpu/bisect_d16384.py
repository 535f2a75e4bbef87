"""Which change breaks DSYGVDX N=16384 m=2048?  Solve + gates under option combinations."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from eigensolver_gpu_b200 import api
from eigensolver_gpu_b200._lib import load
lib = load()
def run(n, m, cplx, opts):
    for k, v in opts.items():
        assert lib.eigb200_set_option(k.encode(), v) == 0
    a0, b0 = bench.make_inputs(torch, n, cplx, "C", 99)
    A, B = a0.clone(), b0.clone()
    info, w, z, ws = api.solve_generalized(A, B, 1, m, skip_host_copy=True)
    g = bench.parity_metrics(torch, a0, b0, w, z, m)
    print(f"n={n} m={m} {'z' if cplx else 'd'} {opts}: info={info} residual_max={g['residual_max']:.3g} b_orth={g['b_orth']:.3g}", flush=True)
    del a0, b0, A, B, ws
    torch.cuda.empty_cache()
run(16384, 2048, False, {"gemm_tma": 0})
run(16384, 2048, False, {"gemm_tma": 1})
run(8192, 1024, False, {"gemm_tma": 1})
run(12288, 1536, False, {"gemm_tma": 1})
run(16384, 2048, False, {"gemm_tma": 1, "trd_l2keep_mb": 0})
run(16384, 16384, False, {"gemm_tma": 1, "trd_l2keep_mb": 32})
