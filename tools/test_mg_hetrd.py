"""torchrun --nproc-per-node P tools/test_mg_hetrd.py N z|d : distributed tridiagonalization vs the single-GPU result."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from eigensolver_gpu_b200 import stages as S, multi_gpu as MG

n = int(sys.argv[1]); cplx = sys.argv[2] == "z"
from eigensolver_gpu_b200._lib import load as _load
for kv in os.environ.get("EIGB_OPTS", "").split(","):          # e.g. EIGB_OPTS=trd_upc=1027,mg_switch_n=2048
    if "=" in kv:
        k, v = kv.split("="); assert _load().eigb200_set_option(k.encode(), int(v)) == 0, kv
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dt = torch.complex128 if cplx else torch.float64
gen = torch.Generator(device="cuda").manual_seed(7)
g = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=gen)
if cplx:
    g = torch.complex(g, torch.randn((n, n), dtype=torch.float64, device="cuda", generator=gen))
a0 = (g + g.conj().T).contiguous()
a1 = a0.clone()
d1, e1, t1 = S.hetrd(a1)
torch.cuda.synchronize()
be = MG.CudaStages(); be.dist_hetrd_min_n = staticmethod(lambda world: 0)
a2 = a0.clone()
d2, e2, t2 = be.hetrd_dist(a2)
torch.cuda.synchronize()
for rep in range(2):
    a2 = a0.clone(); dist.barrier(); torch.cuda.synchronize(); t0 = time.time()
    d2, e2, t2 = be.hetrd_dist(a2); torch.cuda.synchronize(); tm = time.time() - t0
a1 = a0.clone(); torch.cuda.synchronize(); t0 = time.time(); S.hetrd.__wrapped__ if False else None
an = a0.abs().sum(dim=0).max().item()
err_d = (d1 - d2).abs().max().item() / (n * 2.2e-16 * an)
err_e = (e1 - e2).abs().max().item() / (n * 2.2e-16 * an)
err_t = (t1 - t2).abs().max().item()
mask = torch.triu(torch.ones((n, n), device="cuda", dtype=torch.bool), 1)   # tensor [c, r], r < c: reflector storage
err_v = ((a1_ := a1) is None) or 0
a1b = a0.clone(); S_d = None
print(f"rank {rank}: n={n} {'z' if cplx else 'd'} world={dist.get_world_size()} dist hetrd {tm*1e3:.1f} ms | "
      f"|dd|/gate={err_d:.3g} |de|/gate={err_e:.3g} |dtau|={err_t:.3g} finite={bool(torch.isfinite(d2).all())}", flush=True)
if rank == 0:
    import ctypes as C
    from eigensolver_gpu_b200._lib import load
    lib = load()
    lib.eigb200_set_option(b"trd_trace", 1)
dist.barrier()
a2 = a0.clone(); d2, e2, t2 = be.hetrd_dist(a2); torch.cuda.synchronize()
if rank == 0:
    NS = 16
    buf = np.zeros(n * NS, dtype=np.uint64)
    lib.eigb200_trace_read(buf.ctypes.data_as(C.c_void_p), n * NS)
    lib.eigb200_set_option(b"trd_trace", 0)
    t = buf.reshape(n, NS).astype(np.float64)
    def med(sl, a, b):
        ok = (t[sl, a] > 0) & (t[sl, b] > 0)
        return np.median((t[sl, b] - t[sl, a])[ok]) / 1e3 if ok.any() else float("nan")
    for lo in range(n - 512, -1, -1024):
        sl = slice(max(lo, 1), lo + 512)
        print(f" cols {sl.start:5d}-{sl.stop:5d}: phaseA {med(sl,0,1):5.1f} us [gather {med(sl,0,8):4.1f} combine+stores {med(sl,8,6):4.1f} fence+flags {med(sl,6,5):4.1f} sync1 {med(sl,5,7):4.1f} "
              f"VW {med(sl,7,9):4.1f} sync2 {med(sl,9,10):4.1f} flagwait {med(sl,10,11):4.1f} finish {med(sl,11,1):4.1f}] | wait1 {med(sl,1,2):5.1f} | "
              f"phaseB {med(sl,2,3):6.1f} | wait2 {med(sl,3,4):5.1f} | total {med(sl,0,4):6.1f}", flush=True)
# consistency across ranks (must be bitwise identical)
buf = [torch.zeros_like(d2) for _ in range(dist.get_world_size())]
dist.all_gather(buf, d2)
print(f"rank {rank}: d identical across ranks: {all(torch.equal(buf[0], b) for b in buf)}", flush=True)
dist.barrier()
dist.destroy_process_group()
