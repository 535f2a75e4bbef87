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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12288
for o in ({"gemm_tma": 1, "gemm_tma_dbg": 0}, {"gemm_tma": 1, "gemm_tma_dbg": 1}, {"gemm_tma": 1, "gemm_tma_dbg": 2}, {"gemm_tma": 0, "gemm_tma_dbg": 0},
          {"gemm_tma": 1, "gemm_tma_dbg": 0, "potrf_pb": -1}, {"gemm_tma": 1, "gemm_tma_dbg": 0, "potrf_pb": 0, "bt_nb": 64}):
    run(n, n // 8, False, o)
