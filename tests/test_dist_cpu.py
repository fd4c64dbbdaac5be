"""Host-side logic of the y-slab pressure projection (csrc/dist.cu, csrc/distpeer.cuh, stageApplyProjectionDist) on CPU:
the slab partition, and the communication pattern of one fused PCG iteration as the kernels run it -- ghost rows written by
the neighbours, TWO reduction points per iteration (z.s after applyA; z.r together with |r|_inf inside the forward solve,
where the stop rule is decided before beta), each a slot-per-rank exchange combined in rank order so that every rank holds
bit-identical scalars, p += alpha s deferred to the backward solve (or paid after the loop when it ends in the forward
solve), block preconditioner -- run by two gloo processes in numpy and checked against a serial solve of the same system."""
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

    def combine(total, mx):
        """peerCombine (distpeer.cuh): every rank deposits (sum, max) in slot [rank] of every rank's block, then adds the
        slots up in RANK ORDER -- identical bits on every rank, no reduction tree"""
        slots = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(slots, torch.tensor([total, mx], dtype=torch.float64))
        s_, m_ = 0.0, 0.0
        for q in range(world):
            s_ += float(slots[q][0])
            m_ = max(m_, float(slots[q][1]))
        return s_, m_

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
    sigma, rhs_norm = combine(float((z * r).sum()), float(np.abs(r).max()))  # forward solve, phase 0
    iters, pending_p, alpha = 0, False, 0.0
    scalars = []
    while iters < 2000:
        z = _apply_a(with_halo(s), nx)                       # applyA + z.s        (reduction point 1)
        alpha = sigma / combine(float((z * s).sum()), 0.0)[0]
        r -= alpha * z                                       # forward solve: pre warp
        zz = precon(r)
        sigma_new, rn = combine(float((zz * r).sum()), float(np.abs(r).max()))  # (reduction point 2)
        scalars.append((alpha, sigma_new, rn))
        if rn <= 1e-10 * rhs_norm:                           # stop rule first: iter is not incremented, p is still owed
            pending_p = True
            break
        beta = sigma_new / sigma
        sigma = sigma_new
        iters += 1
        p += alpha * s                                       # backward solve: post warp, on the OLD direction
        s = zz + beta * s
    if pending_p:
        p += alpha * s                                       # pcgFinishKernel
    # every rank must hold bit-identical scalars (the kernels gate on them independently)
    mine = torch.tensor(scalars, dtype=torch.float64).reshape(-1)
    ref0 = mine.clone()
    dist.broadcast(ref0, 0)
    assert torch.equal(mine, ref0), "ranks diverged"
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
