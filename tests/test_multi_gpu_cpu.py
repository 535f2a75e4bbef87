"""world_size-2 gloo test (CPU) of the multi-GPU orchestration: column partition, exchanges and the use of C = C^H
in the distributed reduction to standard form.  The stage backend here is the ORACLE (LAPACK/numpy) -- test
infrastructure standing in for the CUDA stages, which need a GPU; the GPU twin is tests/test_multi_gpu_gpu.py."""
import os
import socket

import numpy as np
import pytest
import scipy.linalg as sla
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lapack, matgen, metrics


class OracleStages:
    def _m(self, t):            # tensor (cols, rows) -> numpy matrix view (rows, cols)
        return t.numpy().T

    def potrf(self, b):
        m = self._m(b)
        u = lapack.potrf(np.array(m))
        m[...] = np.triu(u) + np.tril(m, -1)
        return 0

    def trsm_left(self, trans, u, cols):
        if cols.shape[0] == 0:
            return
        um = np.triu(self._m(u))
        x = sla.solve_triangular(um, self._m(cols), trans="C" if trans == "C" else "N", lower=False)
        self._m(cols)[...] = x

    def hetrd(self, a):
        c, d, e, tau = lapack.hetrd(np.array(self._m(a)))
        self._m(a)[...] = c
        return torch.from_numpy(d), torch.from_numpy(e), torch.from_numpy(tau)

    def stedc(self, d, e):
        w, z, info = lapack.stedc(d.numpy(), e.numpy())
        assert info == 0
        return torch.from_numpy(w), torch.from_numpy(np.ascontiguousarray(z.T))

    def ormtr(self, a, tau, zcols):
        if zcols.shape[0] == 0:
            return
        z = lapack.ormtr("L", "U", "N", np.array(self._m(a)), tau.numpy(), np.array(self._m(zcols)))
        self._m(zcols)[...] = z

    def symmetrize_from_upper(self, a):
        m = metrics.full_from_upper(self._m(a))
        return torch.from_numpy(np.ascontiguousarray(m.T))


def _worker(rank, world, port, cplx, n, il, iu, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eigensolver_gpu_b200 import multi_gpu as MG
    a, b = matgen.family_c(n, cplx, seed=77)
    at = torch.from_numpy(np.ascontiguousarray(np.triu(a).T))
    bt = torch.from_numpy(np.ascontiguousarray(np.triu(b).T))
    info, w, z = MG.hegvdx_distributed(at, bt, il, iu, backend=OracleStages())
    if rank == 0:
        out.put((info, w.numpy(), z.numpy().T.copy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("cplx", [False, True])
def test_distributed_orchestration_world2_gloo(cplx):
    n, il, iu = 150, 3, 77
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cplx, n, il, iu, q)) for r in range(2)]
    for p in procs:
        p.start()
    info, w, z = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert info == 0
    a, b = matgen.family_c(n, cplx, seed=77)
    wr, zr, ur, linfo = lapack.hegvd(a, b)
    assert np.abs(w - wr).max() < n * metrics.EPS * np.linalg.norm(a, 2)
    g = metrics.eig_gates(a, b, w[il - 1:iu], z)
    assert g["residual_max"] < 30 and g["b_orth"] < 30


def test_column_ranges_cover_and_align():
    from eigensolver_gpu_b200.multi_gpu import column_ranges
    for n in (1, 63, 64, 65, 1000, 8192):
        for world in (1, 2, 3, 4, 8):
            r = column_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n
            for (a0, a1), (b0, b1) in zip(r[:-1], r[1:]):
                assert a1 == b0 and a0 <= a1
            for c0, c1 in r[:-1]:
                assert c1 % 64 == 0 or c1 == n


def test_c_side_column_partition_matches_the_python_one():
    """eigb200_mg_column_range (mg.cu) is the partition the C-ABI multi-GPU driver uses; multi_gpu.column_ranges is the one
    the gloo tests above exercise: they must be the same function"""
    import ctypes as C
    from eigensolver_gpu_b200 import multi_gpu as MG
    from eigensolver_gpu_b200._lib import load
    lib = load()
    c0, c1 = C.c_int(), C.c_int()
    for ncols in (1, 63, 64, 65, 700, 2500, 8192, 16384, 32768, 4096 + 17):
        for world in (1, 2, 3, 4, 8):
            ref = MG.column_ranges(ncols, world)
            for r in range(world):
                assert lib.eigb200_mg_column_range(ncols, world, r, C.byref(c0), C.byref(c1)) == 0
                assert (c0.value, c1.value) == ref[r]
            assert ref[0][0] == 0 and ref[-1][1] == ncols
