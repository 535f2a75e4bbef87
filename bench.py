#!/usr/bin/env python
"""bench.py -- headline benchmark of the dsygvdx_gpu / zhegvdx_gpu hot path (contract: see README/DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl eigb200|reference|cusolver] [--n 8192] [--dtype z|d] [--m M]

One "step" = one ZHEGVDX solve (N=8192, il=1, iu=8192, complex FP64: BASELINE.json configs[2], the
configuration the metric is quoted on) over synthetic family-R inputs (the reference recipe
A = T T^H, test_driver/test_zhegvdx.F90:28-66).  `value` is achieved GFLOP/s under the nominal flop model
F(N,m) = k [(8/3) N^3 + 3 N^2 m], k = 4 for complex, inputs resident in HBM; `e2e` is the same
metric through the reference-facing call with HOST buffers (H2D of A,B and D2H of Z,w inside the timed
region).  N > 1: one process per GPU (torchrun), ONE problem solved by all ranks (strong scaling):
eigensolver_gpu_b200/multi_gpu.py (DESIGN.md multi-GPU section).

After the timed region every run verifies what it computed (`parity` in the JSON line): residual and
B-orthogonality of the last timed solve (family R, cond(B) ~ 1e9: reported, LAPACK itself misses the gate there) and
the north-star gates on one untimed solve of a conditioned (family C) pair of the same order; with N > 1 ranks
also the eigenvalues against a single-GPU solve of the same pair on rank 0.

--impl reference : the reference's own CPU comparator (LAPACK ?hegvd, test_driver/test_zhegvdx.F90:163-182) on all host
                   cores through the oracle binding: ONE solve of the real order (about 2 min at N=8192) when it is
                   predicted to fit the time budget, else a labelled smaller sample.
--impl cusolver  : cusolverDn?{sy,he}gvdx on the same GPU -- the reference driver's GPU comparator
                   (test_driver/test_zhegvdx.F90:213-263); a reported secondary baseline, never on the product path.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
EPS = 2.220446049250313e-16


def flops_model(n, m, cplx):
    k = 4.0 if cplx else 1.0
    return k * ((8.0 / 3.0) * n ** 3 + 3.0 * n * n * m)


def gemm_flops_model(n, m, cplx):
    """the tensor-core (DMMA) share of the nominal model: everything but the symv/hemv half of the tridiagonalization"""
    k = 4.0 if cplx else 1.0
    return k * ((1.0 / 3.0 + 1.0 + 2.0 / 3.0) * n ** 3 + 3.0 * n * n * m)


def hemv_bytes_model(n, cplx):
    return (16.0 if cplx else 8.0) * n ** 3 / 6.0


# ----------------------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm)); out["sm_max_mhz"] = float(max(mx)); out["reasons"] = sorted(reasons)
            out["samples"] = len(sm)
        return out


# ----------------------------------------------------------------------------------------- naming
def prefix(args):
    return "zhegvdx" if args.dtype == "z" else "dsygvdx"


def metric_name(args, n=None):
    return f"{prefix(args)}_n{n or args.n}_gflops"


def workload_text(args, n=None, m=None):
    n = n or args.n
    m = m or (args.m if n == args.n else n)
    return (f"{prefix(args).upper()} N={n} il=1 iu={m} ({'complex' if args.dtype == 'z' else 'real'} FP64, ITYPE=1 JOBZ=V "
            f"RANGE=I UPLO=U)")


def workload_config(args):
    return {"workload": workload_text(args),
            "inputs": "family R (reference recipe A=T*T^H, B=T*T^H), seed 1234 (same problem on every rank)",
            "l2": "inputs (N^2*16 B each) larger than the 126 MB L2; A,B restored from pristine device copies "
                  "inside the timed region (2 D2D copies per step)",
            "parallelism": ("1 problem over %d GPUs through eigb200_zhegvdx_mg/dsygvdx_mg (C ABI, library-owned NCCL "
                            "communicator): hegst solves, back-transform and final trsm split by columns (NCCL exchanges); "
                            "hetrd trailing matrix 1-D block-cyclic with in-kernel NVLink exchange; potrf/stedc replicated "
                            "(bitwise deterministic)" % args.gpus)
            if args.gpus > 1 else "single"}


# ----------------------------------------------------------------------------------------- CPU comparator
def cpu_reference_run(n, cplx, steps, warmup):
    """Times the reference's own CPU comparator (LAPACK ?hegvd, test_driver/test_zhegvdx.F90:163-182) on the box's
    host cores through the oracle binding.  Returns (GFLOP/s under the nominal model, seconds per step, threads)."""
    from oracle import lapack, matgen
    # torchrun exports OMP_NUM_THREADS=1: the comparator gets all host cores whatever the launcher says
    lapack.set_num_threads(os.cpu_count() or 1)
    a, b = matgen.family_r(n, cplx, seed=1234)
    threads = lapack.num_threads()
    for _ in range(warmup):
        lapack.hegvd(a, b)
    t0 = time.time()
    for _ in range(steps):
        w, z, u, info = lapack.hegvd(a, b)
        assert info == 0
    dt = (time.time() - t0) / max(steps, 1)
    return flops_model(n, n, cplx) / dt * 1e-9, dt, threads


def run_reference(args):
    """One JSON line for the CPU arm.  The real order is solved ONCE (a step of the real workload takes minutes on
    the host: K steps would not fit any driver budget) when the N^3 extrapolation of a small sample says it fits
    `--ref-budget-s`; otherwise the sample itself is reported under ITS OWN metric/workload name."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cplx = args.dtype == "z"
    n_s = min(args.ref_n, args.n)
    val_s, dt_s, threads = cpu_reference_run(n_s, cplx, 1, 1)
    predicted = dt_s * (args.n / n_s) ** 3
    full = (not args.ref_sample_only) and args.m == args.n and n_s < args.n and predicted <= args.ref_budget_s
    if full:
        val, dt, threads = cpu_reference_run(args.n, cplx, 1, 0)
        n_run, steps_run = args.n, 1
        sample = (f"{'zhegvd' if cplx else 'dsygvd'} N={args.n} full spectrum, ONE timed solve ({dt:.1f} s), OpenBLAS {threads} "
                  f"threads; N={n_s} warm-up sample: {val_s:.1f} GFLOP/s ({dt_s:.2f} s)")
    else:
        val, dt, n_run, steps_run = val_s, dt_s, n_s, 1
        why = ("--ref-sample-only" if args.ref_sample_only else
               "subset m < N has no ?hegvd equivalent at equal cost" if args.m != args.n else
               f"predicted {predicted:.0f} s for N={args.n} exceeds --ref-budget-s {args.ref_budget_s:.0f}")
        sample = (f"{'zhegvd' if cplx else 'dsygvd'} N={n_s} full spectrum ({dt_s:.2f} s/solve): a SMALLER problem than the GPU arm's "
                  f"N={args.n} ({why}); rate comparison only, OpenBLAS {threads} threads")
    same = n_run == args.n
    line = {
        "impl": "reference", "metric": metric_name(args, n_run), "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": steps_run, "requested_steps": args.steps, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "c128" if cplx else "f64", "data": "synthetic",
        "config": dict(workload_config(args), workload=workload_text(args, n_run, n_run), same_workload_as_gpu_arm=same,
                       parallelism=f"host CPU, {threads} threads"),
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_time_s": dt,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------- synthetic inputs (device)
def make_inputs(torch, n, cplx, family, seed):
    """(a, b): full Hermitian device tensors.  Family R: the reference recipe; family C: conditioned pair."""
    dt = torch.complex128 if cplx else torch.float64
    gen = torch.Generator(device="cuda").manual_seed(seed)     # every rank builds the same problem

    def herm_uniform():
        t = torch.rand((n, n), dtype=torch.float64, device="cuda", generator=gen)
        if cplx:
            t = torch.complex(t, torch.rand((n, n), dtype=torch.float64, device="cuda", generator=gen))
        t = torch.triu(t) + torch.triu(t, 1).conj().T
        if cplx:
            idx = torch.arange(n, device="cuda")
            t[idx, idx] = t[idx, idx].real.to(dt)
        return t

    def sym(x):
        x = ((x + x.conj().T) / 2).contiguous()
        if cplx:
            idx = torch.arange(n, device="cuda")
            x[idx, idx] = x[idx, idx].real.to(dt)
        return x

    if family == "R":
        t = herm_uniform()
        a = sym(t @ t.conj().T)
        del t
    else:
        g = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=gen)
        if cplx:
            g = torch.complex(g, torch.randn((n, n), dtype=torch.float64, device="cuda", generator=gen))
        a = sym(g)
        del g
    t = herm_uniform()
    b = t @ t.conj().T
    del t
    if family == "C":
        b = b / n + torch.eye(n, dtype=dt, device="cuda")
    return a, sym(b)


def parity_metrics(torch, a, b, w, z, m):
    """Gates of the north star on device data (torch matmul, outside every timed region).  a, b: the full Hermitian
    inputs as handed to the solver (tensor[c, r] = element (r, c)); z: (m, n) tensor = eigenvector columns."""
    n = a.shape[0]
    M, Bm, Z = a.T, b.T, z[:m].T
    wv = w[:m]
    # ||A||_2 by power iteration (converges from below: the gate it feeds is the stricter one)
    v = torch.ones(n, dtype=a.dtype, device=a.device)
    v[::2] = -0.5
    nrm = 0.0
    for _ in range(40):
        v = M @ v
        nrm = float(torch.linalg.vector_norm(v))
        v = v / nrm
    BZ = Bm @ Z
    R = M @ Z - BZ * wv.to(a.dtype)[None, :]
    res = torch.linalg.vector_norm(R, dim=0) / (n * EPS * nrm * torch.linalg.vector_norm(Z, dim=0))
    del R
    G = Z.conj().T @ BZ
    G -= torch.eye(m, dtype=a.dtype, device=a.device)
    orth = float(torch.linalg.matrix_norm(G)) / (n * EPS)
    asc = bool((w[1:] >= w[:-1]).all())
    fin = bool(torch.isfinite(w).all()) and bool(torch.isfinite(res).all())
    return {"residual_max": float(res.max()), "b_orth": orth, "w_ascending": asc, "finite": fin, "normA_2": nrm}


# ----------------------------------------------------------------------------------------- cuSOLVER arm
def run_cusolver(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from tools.cusolver_baseline import CusolverHegvdx
    n, m, cplx = args.n, args.m, args.dtype == "z"
    torch.cuda.set_device(0)
    a0, b0 = make_inputs(torch, n, cplx, "R", 1234)
    A, B = torch.empty_like(a0), torch.empty_like(b0)
    plan = CusolverHegvdx(n, 1, m, cplx)

    def step():
        A.copy_(a0); B.copy_(b0)
        plan.solve(A, B)

    for _ in range(max(args.warmup, 1)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    info = int(plan.info.item())
    par = parity_metrics(torch, a0, b0, plan.w, A, m)        # eigenvectors overwrite A
    val = flops_model(n, m, cplx) / (ms * 1e-3) * 1e-9
    line = {"impl": "cusolver", "metric": metric_name(args), "value": val, "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c128" if cplx else "f64", "data": "synthetic", "config": dict(workload_config(args), parallelism="single"),
            "library": f"cusolverDn{'Zhegvdx' if cplx else 'Dsygvdx'} (CUDA 12.9), lwork {plan.lwork.value}", "devInfo": info,
            "parity": {"family_R_last_solve": par}}
    plan.close()
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="eigb200", choices=["eigb200", "reference", "cusolver"])
    ap.add_argument("--n", "--order", dest="n", type=int, default=8192, help="matrix order (--order under torchrun: its parser rejects --n as ambiguous)")
    ap.add_argument("--m", "--wanted", dest="m", type=int, default=None, help="number of eigenpairs il=1..m (default: all)")
    ap.add_argument("--dtype", default="z", choices=["z", "d"])
    ap.add_argument("--ref-n", type=int, default=2048, help="order of the bounded CPU sample")
    ap.add_argument("--ref-budget-s", type=float, default=600.0, help="reference arm: time budget for ONE solve of the real order")
    ap.add_argument("--ref-sample-only", action="store_true", help="reference arm: never run the real order")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run verification")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (orders whose pinned copies do not fit)")
    ap.add_argument("--no-1gpu-compare", action="store_true", help="N > 1: skip the single-GPU solve of the parity pair")
    args = ap.parse_args()
    if args.m is None:
        args.m = args.n
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "cusolver":
        return run_cusolver(args)

    import torch
    import torch.distributed as dist
    from eigensolver_gpu_b200 import api
    from eigensolver_gpu_b200._lib import load
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load()
    if lib.eigb200_init() != 0:
        raise SystemExit("eigb200_init failed: " + lib.eigb200_last_error().decode())

    n, m, cplx = args.n, args.m, args.dtype == "z"
    dt = torch.complex128 if cplx else torch.float64
    es = 16 if cplx else 8
    a0, b0 = make_inputs(torch, n, cplx, "R", 1234)
    torch.cuda.synchronize()
    A = torch.empty_like(a0)
    B = torch.empty_like(b0)
    ws = api.Workspace(n, cplx, host_z=(not args.no_e2e) and rank == 0)
    c0, c1 = 0, n
    if world > 1:
        from eigensolver_gpu_b200 import multi_gpu as MG
        MG.mg_init()                                   # library-owned NCCL communicator behind the C ABI
        c0, c1 = MG.column_ranges(n, world)[rank]
    if not args.no_e2e:
        # pinned host images; with N ranks every rank holds (and uploads) only its 1/N column slice
        a_host = torch.empty((c1 - c0, n), dtype=dt, pin_memory=True)
        b_host = torch.empty((c1 - c0, n), dtype=dt, pin_memory=True)
        a_host.copy_(a0[c0:c1])
        b_host.copy_(b0[c0:c1])
    last = {}

    def solve_device(single=False):
        if world > 1 and not single:
            info, w, z, _ = api.solve_generalized_mg(A, B, 1, m, ws=ws, skip_host_copy=True)
        else:
            info, w, z, _ = api.solve_generalized(A, B, 1, m, ws=ws, skip_host_copy=True)
        if info != 0:
            raise SystemExit("solve failed: " + lib.eigb200_last_error().decode())
        return w, z

    def step_device():
        A.copy_(a0)
        B.copy_(b0)
        last["w"], last["z"] = solve_device()

    up_stream = torch.cuda.Stream()

    def step_e2e():
        if world == 1:
            # B first on the solver's stream; A on a second stream, overlapped with the Cholesky factorization of B
            # (the solver waits for the event before it touches A); D2H of Z overlaps with the final solve with U
            B.copy_(b_host, non_blocking=True)
            up_stream.wait_stream(torch.cuda.current_stream())       # A's previous use is complete
            ev = torch.cuda.Event()
            with torch.cuda.stream(up_stream):
                A.copy_(a_host, non_blocking=True)
                ev.record(up_stream)
            info, w, z, _ = api.solve_generalized(A, B, 1, m, ws=ws, skip_host_copy=False, a_ready_event=ev)
            if info != 0:
                raise SystemExit("solve failed: " + lib.eigb200_last_error().decode())
            return
        # every rank uploads its column slice over its own PCIe link, the slices are assembled over NVLink; the result
        # (Z(:,1:m), w) is read back to the host of rank 0
        A[c0:c1].copy_(a_host, non_blocking=True)
        B[c0:c1].copy_(b_host, non_blocking=True)
        MG.mg_allgather_columns(B)
        MG.mg_allgather_columns(A)
        info, w, z, _ = api.solve_generalized_mg(A, B, 1, m, ws=ws, skip_host_copy=(rank != 0))
        if info != 0:
            raise SystemExit("solve failed: " + lib.eigb200_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ncu_range = [os.environ.get("EIGB_NCU_RANGE") == "1"]     # ncu --profile-from-start off: capture the timed steps only

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        if ncu_range[0]:
            torch.cuda.profiler.start()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        if ncu_range[0]:
            torch.cuda.profiler.stop()
            ncu_range[0] = False                                   # the device-resident steps only, not the e2e ones
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step_device()
    lib.eigb200_prof_enable(1)
    lib.eigb200_prof_reset()
    sampler = ClockSampler(local) if rank == 0 else None
    ms_total = timed(step_device, args.steps)
    clocks = sampler.stop() if sampler else {}
    pms = (C.c_double * 8)()
    pcnt = (C.c_int * 8)()
    plaunch = C.c_longlong(0)
    lib.eigb200_prof_collect(pms, pcnt, C.byref(plaunch))
    lib.eigb200_prof_enable(0)
    ms_step = ms_total / args.steps
    fl = flops_model(n, m, cplx)
    value = fl / (ms_step * 1e-3) * 1e-9            # one problem, solved by all ranks together

    # end-to-end through the reference-facing call with host buffers
    e2e = None
    if not args.no_e2e:
        step_e2e()
        ms_e2e = timed(step_e2e, args.steps) / args.steps
        e2e = {"value": fl / (ms_e2e * 1e-3) * 1e-9, "unit": "GFLOP/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": 2 * n * n * es, "d2h_bytes_per_step": n * m * es + n * 8}
        del a_host, b_host

    # ---- verification of what was just timed (all ranks take part in the distributed solves) ------------------
    parity = None
    if not args.no_parity:
        parity = {"gates": "residual_max < 30 and b_orth < 30 in units of n*eps (north star); family R has cond(B) ~ 1e9 and "
                           "is reported only (LAPACK itself misses the gate there, SURVEY.md section 4)"}
        if rank == 0:
            parity["family_R_last_timed_solve"] = parity_metrics(torch, a0, b0, last["w"], last["z"], m)
        del a0, b0
        last.clear()
        A = B = None                 # (N=32768 complex: 16 GiB each; the generator below needs the room for its temporaries)
        torch.cuda.empty_cache()
        a0, b0 = make_inputs(torch, n, cplx, "C", 4321)
        torch.cuda.empty_cache()
        A = a0.clone(); B = b0.clone()
        w, z = solve_device()
        if rank == 0:
            pc = parity_metrics(torch, a0, b0, w, z, m)
            if world > 1 and not args.no_1gpu_compare:
                wd = w.clone()
        if world > 1 and not args.no_1gpu_compare:
            barrier()
            if rank == 0:
                A.copy_(a0); B.copy_(b0)
                w1, z1 = solve_device(single=True)
                pc["dlambda_vs_1gpu_over_gate"] = float((wd - w1).abs().max()) / (n * EPS * pc["normA_2"])
            barrier()
        if rank == 0:
            pc["pass"] = bool(pc["residual_max"] < 30 and pc["b_orth"] < 30 and pc["w_ascending"] and pc["finite"] and
                              pc.get("dlambda_vs_1gpu_over_gate", 0.0) < 1.0)
            parity["family_C_same_order"] = pc

    if rank != 0:
        if world > 1:
            MG.mg_finalize()
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        peaks = {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    probe = (C.c_double * 4)()
    live = None
    if lib.eigb200_probe_peaks(probe) == 0:
        live = {"dmma_tflops": probe[0], "dfma_tflops": probe[1], "hbm_read_gbs": probe[2], "hbm_copy_gbs": probe[3]}
    cats = ["potrf", "hegst", "hetrd_panel", "hetrd_her2k", "stedc", "backtransform", "trsm", "other"]
    stages_ms = {c: pms[i] / args.steps for i, c in enumerate(cats) if pcnt[i] > 0}
    panel_ms_per_solve = pms[2] / args.steps
    panel_launches = max(pcnt[2] // args.steps, 1)
    hb = hemv_bytes_model(n, cplx) / world          # bytes THIS GPU streams: its share of the block-cyclic trailing matrix
    achieved = hb / (panel_ms_per_solve * 1e-3) * 1e-9 if panel_ms_per_solve > 0 else 0.0
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (one launch of the same workload)
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f).get(f"{prefix(args)}_n{n}")
        if tr and world == 1:
            traffic = float(tr["dram_bytes_read"] + tr["dram_bytes_write"])
            traffic_note = (f"ncu capture of ONE launch ({tr['capture']}): algorithmic bytes of that launch "
                            f"{tr['algorithmic_bytes_of_that_launch']:.3g}; the average launch moves algorithmic_bytes_per_launch")
    except Exception:
        pass
    roofline = {"kernel": "panel_coop_kernel (hetrd panel: symv/hemv tiles + Householder column phases)",
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                "per_gpu": True, "algorithmic_bytes_per_launch": hb / panel_launches, "launches_per_step": panel_launches,
                "avg_launch_ms": panel_ms_per_solve / panel_launches}
    line = {
        "metric": metric_name(args), "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong",            # one fixed problem whatever the number of GPUs
        "vs_baseline": None, "dtype": "c128" if cplx else "f64", "data": "synthetic",
        "config": workload_config(args),
        "clocks": clocks,
        "gpu_launches": int(plaunch.value),
        "stages_ms": stages_ms,
        "roofline": roofline,
        "wall_time_s": ms_step * 1e-3,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and live is not None:
        gemm_ms = sum(stages_ms.get(c, 0.0) for c in ("potrf", "hegst", "hetrd_her2k", "backtransform", "trsm"))
        gfl = gemm_flops_model(n, m, cplx)
        ach = gfl / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else 0.0
        line["roofline_gemm"] = {"kernel": "gemm_kernel (DMMA m8n8k4): potrf + hegst + rank-2k + back-transform + final trsm",
                                 "bound": "tensor", "achieved": ach, "peak": live["dmma_tflops"], "unit": "TFLOP/s",
                                 "frac": ach / live["dmma_tflops"] if live["dmma_tflops"] > 0 else None, "ms_per_step": gemm_ms,
                                 "peak_source": "eigb200_probe_peaks: DMMA m8n8k4 chains measured live on this device "
                                                "(MEASURED_PEAKS.json has no FP64 entry)"}
    if live is not None:
        line["live_peaks"] = live
    if parity is not None:
        line["parity"] = parity
    if not args.no_cpu and world == 1:
        val, dts, threads = cpu_reference_run(min(args.ref_n, n), cplx, 1, 1)
        line["cpu_baseline"] = {"value": val, "unit": "GFLOP/s", "cores": threads, "kind": "port",
                                "sample": f"{'zhegvd' if cplx else 'dsygvd'} N={min(args.ref_n, n)} full spectrum via the oracle's "
                                          f"LAPACK binding (bounded sample; {dts:.1f} s/solve), same flop model; the reference arm "
                                          f"(--impl reference) times the real order once"}
    print(json.dumps(line), flush=True)
    if world > 1:
        MG.mg_finalize()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
