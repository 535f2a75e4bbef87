#!/usr/bin/env python
"""bench.py -- headline benchmark of the dsygvdx_gpu / zhegvdx_gpu hot path (contract: see README/DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 8192] [--dtype z|d] [--m M]

One "step" = one ZHEGVDX solve (N=8192, il=1, iu=8192, complex FP64: BASELINE.json configs[2], the
configuration the metric is quoted on) over synthetic family-R inputs (the reference recipe
A = T T^H, test_driver/test_zhegvdx.F90:28-66).  `value` is achieved GFLOP/s under the nominal flop model
F(N,m) = k [(8/3) N^3 + 3 N^2 m], k = 4 for complex, inputs resident in HBM; `e2e` is the same
metric through the reference-facing call with HOST buffers (H2D of A,B and D2H of Z,w inside the timed
region).  N > 1: one process per GPU (torchrun), ONE problem solved by all ranks (strong scaling):
eigensolver_gpu_b200/multi_gpu.py -- column-split solves/back-transform with NCCL exchanges, replicated
deterministic potrf/hetrd/stedc (DESIGN.md multi-GPU section).
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


def flops_model(n, m, cplx):
    k = 4.0 if cplx else 1.0
    return k * ((8.0 / 3.0) * n ** 3 + 3.0 * n * n * m)


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


# ----------------------------------------------------------------------------------------- CPU comparator
def cpu_reference_run(n, cplx, steps, warmup):
    """Times the reference's own CPU comparator (LAPACK ?hegvd, test_driver/test_zhegvdx.F90:163-182) on the box's
    host cores through the oracle binding.  Returns (GFLOP/s, seconds per step, threads)."""
    from oracle import lapack, matgen
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_s = args.ref_n
    cplx = args.dtype == "z"
    val, dt, threads = cpu_reference_run(n_s, cplx, args.steps, min(args.warmup, 1))
    sample = (f"{'zhegvd' if cplx else 'dsygvd'} N={n_s} full spectrum (bounded sample of the N={args.n} workload, same "
              f"flop model), OpenBLAS {threads} threads")
    line = {
        "impl": "reference", "metric": metric_name(args), "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "c128" if cplx else "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def metric_name(args):
    return f"{'zhegvdx' if args.dtype == 'z' else 'dsygvdx'}_n{args.n}_gflops"


def workload_config(args):
    return {"workload": f"{'ZHEGVDX' if args.dtype == 'z' else 'DSYGVDX'} N={args.n} il=1 iu={args.m} "
                        f"({'complex' if args.dtype == 'z' else 'real'} FP64, ITYPE=1 JOBZ=V RANGE=I UPLO=U)",
            "inputs": "family R (reference recipe A=T*T^H, B=T*T^H), seed 1234 (same problem on every rank)",
            "l2": "inputs (N^2*16 B each) larger than the 126 MB L2; A,B restored from pristine device copies "
                  "inside the timed region (2 D2D copies per step)",
            "parallelism": ("1 problem over %d GPUs: hegst solves, back-transform and final trsm split by columns (NCCL "
                            "exchanges); hetrd trailing matrix 1-D block-cyclic with in-kernel NVLink exchange when "
                            "N >= multi_gpu.CudaStages.dist_hetrd_min_n(world), else replicated; potrf/stedc replicated "
                            "(bitwise deterministic)" % args.gpus)
            if args.gpus > 1 else "single"}


# ----------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="eigb200")
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--m", type=int, default=None)
    ap.add_argument("--dtype", default="z", choices=["z", "d"])
    ap.add_argument("--ref-n", type=int, default=2048, help="order of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.m is None:
        args.m = args.n
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from eigensolver_gpu_b200 import api, stages as S
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
    # synthetic family-R inputs generated on the device from a seeded generator (reference recipe)
    gen = torch.Generator(device="cuda").manual_seed(1234)     # every rank builds the same problem

    def herm_uniform():
        t = torch.rand((n, n), dtype=torch.float64, device="cuda", generator=gen)
        if cplx:
            t = torch.complex(t, torch.rand((n, n), dtype=torch.float64, device="cuda", generator=gen))
        t = torch.triu(t) + torch.triu(t, 1).conj().T
        if cplx:
            idx = torch.arange(n, device="cuda")
            t[idx, idx] = t[idx, idx].real.to(dt)
        return t

    ta = herm_uniform()
    a0 = ta @ ta.conj().T
    del ta
    tb = herm_uniform()
    b0 = tb @ tb.conj().T
    del tb
    a0 = ((a0 + a0.conj().T) / 2).contiguous()
    b0 = ((b0 + b0.conj().T) / 2).contiguous()
    torch.cuda.synchronize()
    A = torch.empty_like(a0)
    B = torch.empty_like(b0)
    ws = api.Workspace(n, cplx, host_z=True)
    a_host = torch.empty((n, n), dtype=dt, pin_memory=True)
    b_host = torch.empty((n, n), dtype=dt, pin_memory=True)
    a_host.copy_(a0)
    b_host.copy_(b0)

    if world > 1:
        from eigensolver_gpu_b200 import multi_gpu as MG
        mg_backend = MG.CudaStages()

    def step_device():
        A.copy_(a0)
        B.copy_(b0)
        if world > 1:
            info, w, z = MG.hegvdx_distributed(A, B, 1, m, backend=mg_backend, gather_z=True)
        else:
            info, w, z, _ = api.solve_generalized(A, B, 1, m, ws=ws, skip_host_copy=True)
        if info != 0:
            raise SystemExit("solve failed: " + lib.eigb200_last_error().decode())

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
        A.copy_(a_host, non_blocking=True)
        B.copy_(b_host, non_blocking=True)
        if world > 1:
            info, w, z = MG.hegvdx_distributed(A, B, 1, m, backend=mg_backend, gather_z=True)
            if rank == 0:
                ws.Z_h[:m].copy_(z)
                ws.w_h.copy_(w)
                torch.cuda.synchronize()
        else:
            info, w, z, _ = api.solve_generalized(A, B, 1, m, ws=ws, skip_host_copy=False)   # D2H of Z(:,1:m), w inside
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
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e_val = fl / (ms_e2e * 1e-3) * 1e-9
    h2d = 2 * n * n * es
    d2h = n * m * es + n * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # parity spot check on the last solve (residual on a few eigenpairs; full gates live in tests/)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        peaks = {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    cats = ["potrf", "hegst", "hetrd_panel", "hetrd_her2k", "stedc", "backtransform", "trsm", "other"]
    stages_ms = {c: pms[i] / args.steps for i, c in enumerate(cats) if pcnt[i] > 0}
    panel_ms_per_solve = pms[2] / args.steps
    panel_launches = max(pcnt[2] // args.steps, 1)
    hb = hemv_bytes_model(n, cplx)
    achieved = hb / (panel_ms_per_solve * 1e-3) * 1e-9 if panel_ms_per_solve > 0 else 0.0
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (one launch of the same workload)
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tr = json.load(f).get(f"{'zhegvdx' if cplx else 'dsygvdx'}_n{n}")
        if tr:
            traffic = float(tr["dram_bytes_read"] + tr["dram_bytes_write"])
            traffic_note = (f"ncu capture of ONE launch ({tr['capture']}): algorithmic bytes of that launch "
                            f"{tr['algorithmic_bytes_of_that_launch']:.3g}; the average launch moves algorithmic_bytes_per_launch")
    except Exception:
        pass
    roofline = {"kernel": "panel_coop_kernel (hetrd panel: symv/hemv tiles + Householder column phases)",
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": hb / panel_launches, "launches_per_step": panel_launches,
                "avg_launch_ms": panel_ms_per_solve / panel_launches}
    line = {
        "metric": metric_name(args), "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong",            # one fixed problem whatever the number of GPUs
        "vs_baseline": None, "dtype": "c128" if cplx else "f64", "data": "synthetic",
        "config": workload_config(args),
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "GFLOP/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": int(plaunch.value),
        "stages_ms": stages_ms,
        "roofline": roofline,
        "wall_time_s": ms_step * 1e-3,
    }
    if not args.no_cpu and world == 1:
        val, dts, threads = cpu_reference_run(args.ref_n, cplx, 1, 1)
        line["cpu_baseline"] = {"value": val, "unit": "GFLOP/s", "cores": threads, "kind": "port",
                                "sample": f"{'zhegvd' if cplx else 'dsygvd'} N={args.ref_n} full spectrum via the oracle's "
                                          f"LAPACK binding (bounded sample; {dts:.1f} s/solve), same flop model"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
