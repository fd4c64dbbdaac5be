"""Host-side logic of the y-slab pressure projection (csrc/dist.cu, stageApplyProjectionDist) on CPU:
the slab partition, and the communication pattern of one PCG iteration -- one-row halo exchange with both
neighbours, allreduce(sum) of the dot products, allreduce(max) of the residual norm, block preconditioner --
run by two gloo processes in numpy and checked against a serial solve of the same 5-point system."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

fs = importlib.import_module("fluid-sim_b200")


def test_slab_partition_covers_all_rows():
    for ny in (32, 64, 100, 128, 1000, 4096, 16384):
        for world in (1, 2, 4, 8):
            if world > (ny + 31) // 32:
                continue
            rows = [fs.slab_rows(ny, world, r) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == ny
            for (a0, a1), (b0, b1) in zip(rows, rows[1:]):
                assert a1 == b0 and a0 % 32 == 0 and b0 % 32 == 0
            strips = [(b - a + 31) // 32 for a, b in rows]
            assert max(strips) - min(strips) <= 1  # as even as whole strips allow


def _apply_a(s_ext, nx):
    """5-point Poisson operator (Dirichlet outside) on the own rows of a slab given with one halo row each side"""
    s = s_ext[1:-1]
    z = 4.0 * s - s_ext[:-2] - s_ext[2:]
    z[:, 1:] -= s[:, :-1]
    z[:, :-1] -= s[:, 1:]
    return z


def _slab_pcg(rank, world, port, ny, nx, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    rhs_full = rng.standard_normal((ny, nx))
    j0, j1 = fs.slab_rows(ny, world, rank)
    r = rhs_full[j0:j1].copy()
    p = np.zeros_like(r)

    def allreduce(x, op):
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t[0])

    def with_halo(s):
        ext = np.zeros((s.shape[0] + 2, nx))
        ext[1:-1] = s
        reqs = []
        lo, hi = torch.zeros(nx, dtype=torch.float64), torch.zeros(nx, dtype=torch.float64)
        if rank > 0:
            reqs += [dist.isend(torch.from_numpy(s[0].copy()), rank - 1), dist.irecv(lo, rank - 1)]
        if rank < world - 1:
            reqs += [dist.isend(torch.from_numpy(s[-1].copy()), rank + 1), dist.irecv(hi, rank + 1)]
        for q in reqs:
            q.wait()
        ext[0], ext[-1] = lo.numpy(), hi.numpy()
        return ext

    precon = lambda v: v / 4.0  # block-diagonal (here: Jacobi) -- no coupling across slabs, like block-MIC(0)
    z = precon(r)
    s = z.copy()
    sigma = allreduce(float((z * r).sum()), dist.ReduceOp.SUM)
    rhs_norm = allreduce(float(np.abs(r).max()), dist.ReduceOp.MAX)
    iters = 0
    for iters in range(1, 2000):
        z = _apply_a(with_halo(s), nx)
        alpha = sigma / allreduce(float((z * s).sum()), dist.ReduceOp.SUM)
        p += alpha * s
        r -= alpha * z
        if allreduce(float(np.abs(r).max()), dist.ReduceOp.MAX) <= 1e-10 * rhs_norm:
            break
        z = precon(r)
        sigma_new = allreduce(float((z * r).sum()), dist.ReduceOp.SUM)
        s = z + (sigma_new / sigma) * s
        sigma = sigma_new
    gathered = [torch.zeros(fs.slab_rows(ny, world, q)[1] - fs.slab_rows(ny, world, q)[0], nx, dtype=torch.float64)
                for q in range(world)]
    dist.all_gather(gathered, torch.from_numpy(p)) if len({g.shape for g in gathered}) == 1 else None
    if rank == 0:
        out.put((iters, torch.cat(gathered).numpy() if len({g.shape for g in gathered}) == 1 else None, rhs_full))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_slab_pcg_matches_serial_solve():
    ny, nx, world = 64, 24, 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_slab_pcg, args=(r, world, port, ny, nx, out)) for r in range(world)]
    for q in procs:
        q.start()
    iters, p, rhs = out.get(timeout=100)
    for q in procs:
        q.join(timeout=60)
        assert q.exitcode == 0
    # serial reference: the same operator as one dense solve
    n = ny * nx
    a = np.zeros((n, n))
    for j in range(ny):
        for i in range(nx):
            k = j * nx + i
            a[k, k] = 4.0
            if i > 0: a[k, k - 1] = -1.0
            if i < nx - 1: a[k, k + 1] = -1.0
            if j > 0: a[k, k - nx] = -1.0
            if j < ny - 1: a[k, k + nx] = -1.0
    want = np.linalg.solve(a, rhs.ravel()).reshape(ny, nx)
    assert iters < 2000
    assert np.abs(p - want).max() <= 1e-8 * np.abs(want).max()
