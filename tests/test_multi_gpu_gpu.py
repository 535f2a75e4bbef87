"""2-GPU test of the distributed driver (skipped unless >= 2 CUDA devices): NCCL ranks spawned from the test."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import lapack, matgen, metrics

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, cplx, n, il, iu, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from eigensolver_gpu_b200 import multi_gpu as MG, stages as S
    a, b = matgen.family_c(n, cplx, seed=5)
    info, w, z = MG.hegvdx_distributed(S.to_dev(np.triu(a)), S.to_dev(np.triu(b)), il, iu)
    if rank == 0:
        out.put((info, S.to_host(w).copy(), np.array(S.to_host(z))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cplx", [False, True])
def test_distributed_solve_two_gpus(cplx):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    n, il, iu = 700, 1, 300
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cplx, n, il, iu, q)) for r in range(2)]
    for p in procs:
        p.start()
    info, w, z = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    a, b = matgen.family_c(n, cplx, seed=5)
    wr, zr, ur, linfo = lapack.hegvd(a, b)
    assert info == 0
    assert np.abs(w - wr).max() < n * metrics.EPS * np.linalg.norm(a, 2)
    g = metrics.eig_gates(a, b, w[il - 1:iu], z)
    assert g["residual_max"] < 30 and g["b_orth"] < 30


def _worker_hetrd(rank, world, port, cplx, n, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from eigensolver_gpu_b200 import multi_gpu as MG, stages as S
    a, _ = matgen.family_c(n, cplx, seed=8)
    be = MG.CudaStages()
    be.dist_hetrd_min_n = lambda world: 0          # force the distributed tridiagonalization
    ad = S.to_dev(np.triu(a))
    d, e, tau = be.hetrd_dist(ad)
    ds = [torch.zeros_like(d) for _ in range(world)]
    dist.all_gather(ds, d)
    same = all(torch.equal(ds[0], x) for x in ds)
    if rank == 0:
        out.put((S.to_host(d).copy(), S.to_host(e).copy(), same))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cplx", [False, True])
def test_distributed_tridiagonalization_two_gpus(cplx):
    """1-D block-cyclic trailing matrix + in-kernel peer exchange: same (d, e) as LAPACK, bitwise equal on all ranks"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    n = 1300
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [ctx.Process(target=_worker_hetrd, args=(r, 2, port, cplx, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    d, e, same = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    a, _ = matgen.family_c(n, cplx, seed=8)
    _, dl, el, _ = lapack.hetrd(a)
    an = np.abs(a).sum(axis=0).max()
    assert same
    assert np.abs(d - dl).max() <= 20 * n * metrics.EPS * an
    assert np.abs(e - el).max() <= 20 * n * metrics.EPS * an
