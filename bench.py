#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: Mcell-steps/s of a full FluidSim2D::update() (PIC/FLIP, PCG+MIC(0)
projection included) on the 4096x4096 dam break, state resident in HBM.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl b200|reference]

One JSON line on stdout (rank 0).  Keys follow the driver contract:
  value      whole-job Mcell-steps/s, device-timed (CUDA events on the library's stream), max over ranks
  e2e        the same metric through fsim_step_host with pinned HOST mirrors: per step u,v uploaded and
             u,v,p,cell,phi,particles,particleVels downloaded inside the timed region
  roofline   dominant kernel of the PCG iteration (the one with the largest summed CUDA-event time, one of the two MIC(0)
             triangular solves): algorithmic bytes / CUDA-event duration against the measured HBM copy bandwidth
             (MEASURED_PEAKS.json)
  cpu_baseline  the stock reference (oracle/_ref, kind "reference") or the C restatement (kind "port") timed on
             the host cores on a bounded sample of the same scene
  --impl reference   the reference's own OpenMP + AVX/FMA build on the host cores, SAME configuration, bounded to --cpu-steps
             updates (same_config: true); the cheaper 1024^2 samples are listed beside it
Multi-GPU (N>1, launched by torchrun, one process per GPU): the SAME workload, strong scaling.  The pressure
projection is partitioned into y-slabs; per iteration the one-row halo of the search direction and the PCG scalars travel
through peer memory (NVLink stores from inside the solve / applyA kernels, no collective call in the iteration),
block-MIC(0) across slab boundaries; the other stages run replicated on every rank -- see DESIGN.md section 5.  Before
anything is timed the N-rank projection is checked against the oracle on rank 0 (config.parity_checked).  `--replicas`
runs N independent copies instead (weak scaling, no collective).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "Mcell-steps/s at 4096^2 FLIP (incl. PCG)"
UNIT = "Mcell-steps/s"
# Bytes each PCG kernel has to move per cell it covers (8 B per operand streamed once; stencil neighbours are cache hits;
# labels and precon are folded into coefficient arrays that are zero outside the fluid, and z = M^-1 r is never stored):
#   applyA+dot        R s, Adiag, Ax, Ay; W z                                   = 40
#   fused (default)   forward:  R r, z, Lx, Ly, D; W w, r (r -= alpha z inside)   = 56
#                     backward: R w, Ux, Uy, s, p; W s, p (p += alpha s inside)   = 56   -> 152 B/cell/iteration
#   unfused           axpy R p, s, r, z; W p, r = 48; forward R r, Lx, Ly, D; W w = 40; backward R w, Ux, Uy, s; W s = 40
# (SURVEY.md section 8d's 203 B/cell/iteration is the reference's own pass structure; it is kept for the dense-model figure.)
ALGO_BYTES_FUSED = {0: 40, 2: 56, 3: 56}
ALGO_BYTES_UNFUSED = {0: 40, 1: 48, 2: 40, 3: 40}
KNAMES = {0: "applyA+dot", 1: "axpy+norm", 2: "mic0_forward+dot", 3: "mic0_backward+s_update"}
# other latency-bound kernels of the step, timed the same way (reported, not part of the roofline choice)
XNAMES = {5: "ls_closest_particle_sweep", 6: "ls_eikonal_sweep", 7: "extrapolate_layer_fill", 8: "mic0_factor",
          9: "extrapolate_distance_transform+sort"}


_JSON_FD = None


def claim_stdout():
    """stdout carries the JSON line(s) only: everything else that writes to fd 1 (NCCL prints its version banner there)
    is sent to stderr from here on"""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def _scenes():
    return importlib.import_module("fluid-sim_b200.scenes")


def scene(n):
    return _scenes().dam_break_cells(n)


def scene_params(n):
    # config 2 / headline: dx = 1.28/N, dt scaled to keep the demo's CFL number
    return _scenes().dam_break_params(n)


class ClockSampler(threading.Thread):
    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.stop_flag = device, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(r[0])) for r in self.rows)
        reasons = []
        for idx, name in ((2, "hw_slowdown"), (3, "hw_thermal_slowdown"), (4, "sw_thermal_slowdown"), (5, "sw_power_cap")):
            if any(r[idx].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.rows[0][1])), "reasons": reasons}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_run(n, steps, warmup, threads_all):
    """times FluidSim2D::update() of the reference on the host cores; returns (Mcell-steps/s, kind, cores, desc, iters)"""
    import oracle_lib as ol
    kind_serial = "ref" if ol.available("ref") else "port"
    results = []
    variants = [(kind_serial, 1)]
    if threads_all and ol.available("ref_omp"):
        variants.append(("ref_omp", os.cpu_count() or 1))
    for kind, cores in variants:
        os.environ["OMP_NUM_THREADS"] = str(cores)
        sim = ol.OracleSim(kind, scene(n), mode=ol.PICFLIP, alpha=0.05, **scene_params(n))
        sim.step(warmup)
        t0 = time.perf_counter()
        sim.step(steps)
        dt = time.perf_counter() - t0
        results.append((n * n * steps / dt / 1e6, "reference" if kind.startswith("ref") else "port", cores,
                        "%dx%d PIC/FLIP dam break, %d steps after %d warm-up, %s build" % (
                            n, n, steps, warmup, "OpenMP" if kind == "ref_omp" else "serial -O2 -mavx -mfma"), dt / steps))
        sim.close()
    return max(results, key=lambda r: r[0]), results


def run_reference_arm(args, rank):
    """The reference's own CPU implementation on the box's host cores.  value = the SAME configuration as the GPU arm
    (4096^2 PIC/FLIP dam break) with the OpenMP + AVX/FMA build on every core, bounded to `--cpu-steps` updates from the
    initial state (a step is ~35-90 s; every one of them runs the 200 capped PCG iterations, like the GPU arm's steps).
    The cheaper 1024^2 samples of both builds are reported beside it (cpu_baseline.other_samples)."""
    if rank != 0:
        return
    import oracle_lib as ol
    n = args.size
    kind = "ref_omp" if ol.available("ref_omp") else ("ref" if ol.available("ref") else "port")
    cores = (os.cpu_count() or 1) if kind == "ref_omp" else 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    nsteps = max(1, min(args.cpu_steps, args.steps))
    sim = ol.OracleSim(kind, scene(n), mode=ol.PICFLIP, alpha=0.05, **scene_params(n))
    t0 = time.perf_counter()
    sim.step(nsteps)
    sec = (time.perf_counter() - t0) / nsteps
    sim.close()
    val = n * n / sec / 1e6
    desc = "%dx%d PIC/FLIP dam break (same configuration as the GPU arm), first %d update(s) from the initial state, %s build on %d core(s)" % (
        n, n, nsteps, "OpenMP -O2 -mavx -mfma" if kind == "ref_omp" else "serial -O2 -mavx -mfma", cores)
    others = []
    if args.cpu_size and args.cpu_size != n and not args.no_cpu:
        _, allr = cpu_reference_run(args.cpu_size, 2, 1, True)
        others = [{"value": r[0], "cores": r[2], "sample": r[3]} for r in allr]
    line = {"impl": "reference", "metric": METRIC if n == 4096 else METRIC.replace("4096", str(n)), "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": nsteps, "warmup": 0, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "same_config": True,
            "config": {"workload": "%dx%d PIC/FLIP dam break (flip 0.95, 2x2 ppc), full FluidSim2D::update incl. PCG+MIC(0) (tol 1e-12, cap 200); "
                                   "CPU arm bounded to %d update(s) (requested --steps %d --warmup %d)" % (n, n, nsteps, args.steps, args.warmup)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference" if kind.startswith("ref") else "port",
                             "sample": desc, "other_samples": others},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def splitmix_uniform(count, seed):
    return _scenes().splitmix_uniform(count, seed)


def run_projection_stress(args):
    """BASELINE config 3: 4096x4096 projection only, random interior face velocities, PCG + MIC(0) to a relative residual
    of 1e-6 (cap 10000).  3a closed tank (border SOLID, interior FLUID); 3b free surface (rows j >= 3/4 N EMPTY with
    phi = (j - 3N/4 + 1/2) dx, fluid phi = -dx).  Prints one JSON line per variant: iterations, iter/s, bytes/iteration."""
    import torch
    fs = importlib.import_module("fluid-sim_b200")
    n = args.size
    dx = 1.0 / n
    peak, peak_src = peaks()
    for variant in ("3a closed tank", "3b free surface"):
        cells = np.full((n, n), fs.FS_FLUID, np.uint8)
        cells[0, :] = cells[-1, :] = fs.FS_SOLID
        cells[:, 0] = cells[:, -1] = fs.FS_SOLID
        phi = np.full((n, n), -dx)
        if variant.startswith("3b"):
            top = 3 * n // 4
            cells[top:-1, 1:-1] = fs.FS_EMPTY
            phi[top:, :] = ((np.arange(top, n) - top + 0.5) * dx)[:, None]
        u = splitmix_uniform(n * (n + 1), 0x5EED).reshape(n, n + 1)
        v = splitmix_uniform((n + 1) * n, 0x5EED + 1).reshape(n + 1, n)
        u[:, :2] = 0; u[:, -2:] = 0; v[:2, :] = 0; v[-2:, :] = 0  # faces touching SOLID
        sim = fs.FluidSim2D(cells, dt=dx, dx=dx, pcgTol=1e-6, pcgMaxIters=10000, seedParticles=False, computeStats=False)
        res = []
        for rep in range(max(1, args.warmup) + 1):  # warm-up projections, then the timed one
            sim.set(fs.U, u); sim.set(fs.V, v); sim.set(fs.PHI, phi)
            sim.sync(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            sim.applyProjection(); sim.sync()
            res.append(time.perf_counter() - t0)
        st = sim.stats()
        secs, iters = res[-1], st.pcgIters
        cells_m = int(st.pcgMarchedCells) if st.pcgMarchedCells > 0 else n * n
        line = {"metric": "PCG iter/s at %d^2 projection-only stress (config %s)" % (n, variant), "value": iters / secs, "unit": "iter/s",
                "n_gpus": 1, "iterations": iters, "ms_projection": secs * 1e3, "relative_residual": st.pcgResidual / st.pcgRhsNorm,
                "bytes_per_iteration_algorithmic": 203 * n * n, "cells_marched_per_kernel": cells_m,
                "hbm_frac_algorithmic": 203 * n * n * iters / secs / 1e9 / peak, "peak": peak, "peak_source": peak_src,
                "dtype": "f64", "data": "synthetic (SplitMix64 seed 0x5EED)", "higher_is_better": True}
        emit(line)
        sim.free()


def run_semilagrangian(args):
    """BASELINE config 1 (the reference's own demo scene: 128x128 dam break, semi-Lagrangian advection + PCG, 100 steps
    headless; the stock reference is timed beside it on the host cores) and, with --size 8192, the per-GPU workload of
    config 5.  One JSON line per advection variant: the reference's in-place raster-order semantics (default, exact)
    and the snapshot variant (fsim_options.slDoubleBuffer)."""
    import torch
    fs = importlib.import_module("fluid-sim_b200")
    n = args.size
    steps, warmup = (100, 10) if n <= 256 else (max(1, args.steps), max(1, args.warmup))
    peak, peak_src = peaks()
    cpu = None
    if n <= 512 and not args.no_cpu:
        import oracle_lib as ol
        cpu = []
        variants = [("ref" if ol.available("ref") else "port", 1)]
        if ol.available("ref_omp"):
            variants.append(("ref_omp", os.cpu_count() or 1))
        for kind, cores in variants:
            os.environ["OMP_NUM_THREADS"] = str(cores)
            o = ol.OracleSim(kind, scene(n), mode=ol.SEMILAGRANGIAN, **scene_params(n))
            o.step(warmup)
            t0 = time.perf_counter()
            o.step(steps)
            dt = (time.perf_counter() - t0) / steps
            cpu.append({"value": n * n / dt / 1e6, "unit": UNIT, "ms_per_step": dt * 1e3, "cores": cores,
                        "kind": "reference" if kind.startswith("ref") else "port",
                        "sample": "%dx%d semi-Lagrangian dam break, %d steps after %d warm-up, %s build" % (
                            n, n, steps, warmup, "OpenMP" if kind == "ref_omp" else "serial -O2 -mavx -mfma")})
            o.close()
    for snapshot in (False, True):
        sim = fs.FluidSim2D(scene(n), mode=fs.FS_SEMILAGRANGIAN, slDoubleBuffer=snapshot, **scene_params(n))
        sim.update(warmup); sim.sync(); torch.cuda.synchronize()
        l0 = sim.launch_count
        t0 = time.perf_counter()
        sim.update(steps); sim.sync()
        dt = (time.perf_counter() - t0) / steps
        st = sim.stats()
        iters = st.pcgIters
        step_bytes = n * n * (1208 + 203 * iters)  # SURVEY.md 8d (no particle<->grid transfers in this mode)
        line = {"metric": "Mcell-steps/s at %d^2 semi-Lagrangian + PCG (config %s)" % (n, "1" if n == 128 else "5, one GPU's share" if n == 8192 else "-"),
                "value": n * n / dt / 1e6, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3,
                "higher_is_better": True, "dtype": "f64", "data": "synthetic", "gpu_launches": sim.launch_count - l0,
                "config": {"workload": "%dx%d dam break, semi-Lagrangian advection (%s) + PCG+MIC(0) (tol 1e-12, cap 200)" % (
                               n, n, "snapshot variant, slDoubleBuffer" if snapshot else "the reference's in-place raster order, exact"),
                           "pcg_iters_last_step": iters, "stage_ms_last_step": [float(x) for x in st.stageMs[:st.numStages]],
                           "step_dense_model_frac": step_bytes / dt / 1e9 / peak,
                           "step_dense_model_note": "SURVEY 8d byte model over all cells / step time / peak: work-equivalent, not HBM utilisation",
                           "peak": peak, "peak_source": peak_src}}
        if cpu:
            line["cpu_baseline"] = max(cpu, key=lambda c: c["value"])
            line["cpu_baseline"]["all_builds"] = [dict(c) for c in cpu]
        emit(line)
        sim.free()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--cpu-size", type=int, default=1024, dest="cpu_size",
                    help="grid of the cheap CPU samples reported beside the numbers (cpu_baseline of the GPU arm; other_samples of the reference arm)")
    ap.add_argument("--cpu-steps", type=int, default=2, dest="cpu_steps", help="--impl reference: updates timed at the full size")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", dest="no_parity", help="N>1: skip the oracle check of the slab projection")
    ap.add_argument("--replicas", action="store_true", help="N>1: independent replicas instead of the y-slab projection")
    ap.add_argument("--workload", default="flip", choices=["flip", "projection", "sl"],
                    help="flip: the headline metric; projection: BASELINE config 3 (projection-only stress, PCG to 1e-6); "
                         "sl: semi-Lagrangian + PCG (config 1 with --size 128, config 5's per-GPU share with --size 8192) -- "
                         "extra measurements, one JSON line per variant")
    args = ap.parse_args()
    claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.workload == "projection":
        run_projection_stress(args)
        return
    if args.workload == "sl":
        run_semilagrangian(args)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    fs = importlib.import_module("fluid-sim_b200")

    n = args.size
    cells = scene(n)
    sim = fs.FluidSim2D(cells, mode=fs.FS_PICFLIP, picFlipAlpha=0.05, device=local_rank, **scene_params(n))
    npart = sim.num_particles
    slabs = world > 1 and not args.replicas

    def join(sm):
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(fs.dist_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        sm.dist_init(rank, world, bytes(idt.cpu().tolist()))

    parity = None
    if slabs:
        join(sim)
        if not args.no_parity:
            # the N-rank projection against the ORACLE before anything is timed (the checker, not the product): config 3 in
            # miniature at a stop rule both meet + the distance at the stock cap; carried in config.parity_checked
            from dist_parity import slab_parity_vs_oracle
            parity = slab_parity_vs_oracle(join, rank, world, local_rank)
    jobs = 1 if (slabs or world == 1) else world  # independent simulations in flight
    parallelism_desc = ("y-slab PCG over %d GPUs: halo rows of s and the PCG scalars through peer memory (NVLink stores from inside the "
                        "solve / applyA kernels, no collective call in the iteration), block-MIC(0) across slabs; other stages replicated" % world)

    def barrier():
        if world > 1:
            dist.barrier()
        sim.sync()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- device-resident timing ------------------------------------------------------------------
    sim.update(args.warmup)
    barrier()
    launches0 = sim.launch_count
    barrier()
    t0 = time.perf_counter()
    dev_ms = sim.update_timed(args.steps)  # CUDA events on the library's stream around exactly `steps` updates
    wall = time.perf_counter() - t0
    st = sim.stats()
    stage_ms = [float(x) for x in st.stageMs[:st.numStages]]
    launches = sim.launch_count - launches0
    barrier()
    # per-kernel CUDA-event times on the library's stream: one more step, outside the timed region (two event records
    # around each of ~750 launches cost several ms per step)
    sim.profile_enable(True)
    sim.update(1)
    sim.sync()
    prof = {k: sim.profile_get(k) for k in range(4)}
    xprof = {k: sim.profile_get(k) for k in XNAMES}
    sim.profile_enable(False)
    barrier()
    secs = torch.tensor([dev_ms * 1e-3, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)  # max over ranks
    secs, wall = float(secs[0].item()), float(secs[1].item())
    value = jobs * n * n * args.steps / secs / 1e6

    # ---- end to end through the host-buffer call ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        pin = lambda shape, dt=np.float64: torch.empty(shape, dtype=torch.float64 if dt == np.float64 else torch.uint8).pin_memory()  # noqa: E731
        bufs = {"u": pin((n, n + 1)), "v": pin((n + 1, n)), "p": pin((n, n)), "phi": pin((n, n)),
                "cell": pin((n, n), np.uint8), "particles": pin((npart, 2)), "particleVels": pin((npart, 2))}
        m = fs.FsimHostMirror()
        for k, t in bufs.items():
            setattr(m, k, t.data_ptr())
        sim.step_host(m)  # fills the mirrors (untimed)
        m.u_in, m.v_in = bufs["u"].data_ptr(), bufs["v"].data_ptr()
        h2d = bufs["u"].numel() * 8 + bufs["v"].numel() * 8
        d2h = sum(t.numel() * t.element_size() for t in bufs.values())
        if slabs and rank > 0:
            # the state is replicated: the caller's results come from rank 0; the other ranks only take the caller's u, v
            # (every rank must see the same host edits) and read nothing back
            m = fs.FsimHostMirror()
            m.u_in, m.v_in = bufs["u"].data_ptr(), bufs["v"].data_ptr()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sim.step_host(m)
        e2e_secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_secs, op=dist.ReduceOp.MAX)
        e2e = {"value": jobs * n * n * args.steps / float(e2e_secs.item()) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": h2d * (world if slabs else 1), "d2h_bytes_per_step": d2h}
        if slabs:
            e2e["note"] = "u, v uploaded on every rank (replicated state), every field downloaded on rank 0 only"
    sampler.stop_flag = True

    if rank == 0:
        peak, peak_src = peaks()
        cells_n = n * n
        # cells one PCG kernel launch covers on this rank: the fluid cells' bounding box (whole strips), or the slab
        cells_k = int(st.pcgSolveCells) if st.pcgSolveCells > 0 else (cells_n // world if slabs else cells_n)
        # ... and the triangular solves only march the chunks that hold fluid
        cells_m = int(st.pcgMarchedCells) if st.pcgMarchedCells > 0 else cells_k
        fused = prof[1][1] == 0  # no axpy launches: the axpys ran inside the solves
        algo = ALGO_BYTES_FUSED if fused else ALGO_BYTES_UNFUSED
        kinfo = {}
        for k, (ms, cnt) in prof.items():
            if cnt:
                kinfo[KNAMES[k]] = {"launches": cnt, "avg_ms": ms / cnt, "cells_per_launch": cells_m, "bytes_per_cell": algo[k],
                                    "gbs": algo[k] * cells_m / (ms / cnt * 1e-3) / 1e9}
        for k, (ms, cnt) in xprof.items():
            if cnt:
                kinfo[XNAMES[k]] = {"launches": cnt, "avg_ms": ms / cnt}
        dom = max(prof, key=lambda k: prof[k][0])
        ms, cnt = prof[dom]
        achieved = algo[dom] * cells_m / (ms / cnt * 1e-3) / 1e9 if cnt else 0.0
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tpath) and world == 1 and n == 4096:
            tj = json.load(open(tpath))
            traffic, traffic_src = tj["bytes_per_launch"].get(KNAMES[dom]), tj["source"]
        iters = st.pcgIters
        step_s = secs / args.steps
        # (a) the reference's own pass structure over every cell of the grid (SURVEY.md 8d / BASELINE.md section 4): a
        #     work-equivalent figure, NOT a bandwidth -- the kernels skip what is exactly zero (cells outside the fluid's box)
        dense_bytes = cells_n * (1208 + 203 * iters) + 128 * npart
        # (b) bytes of the cells the kernels actually touch: the PCG iteration on the marched chunks, the projection's
        #     set-up on the fluid box, the rest of the step over the whole grid (an upper bound: the level-set sweeps
        #     early-out on converged cells)
        per_iter = sum(algo.values())
        touched_bytes = cells_m * per_iter * iters + cells_k * 211 + cells_n * 997 + 128 * npart
        pcg_ms = sum(ms_ for ms_, _ in prof.values())
        line = {"metric": METRIC if n == 4096 else METRIC.replace("4096", str(n)), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong" if not args.replicas else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "%dx%d PIC/FLIP dam break (picFlipAlpha 0.05 = flip 0.95, 2x2 particles/cell, %d particles), "
                                       "full FluidSim2D::update incl. PCG+MIC(0) (tol 1e-12, cap 200)" % (n, n, npart),
                           "parallelism": (parallelism_desc if slabs else ("replicas only" if world > 1 else "1 GPU")),
                           "timing": "CUDA events on the library's stream around the %d timed updates (fsim_step_timed), max over ranks; host wall clock of the same region %.1f ms/step" % (args.steps, wall / args.steps * 1e3),
                           "pcg_residual_last_step": st.pcgResidual / st.pcgRhsNorm if st.pcgRhsNorm else None, "l2": "working set %.1f GB >> 126 MB L2" % (
                               25 * cells_n * 8 / 1e9), "pcg_iters_last_step": iters, "pcg_cells_per_launch": cells_k,
                           "pcg_cells_marched": cells_m, "pcg_axpys": "inside the triangular solves" if fused else "separate kernel",
                           "extrapolation": "%d BFS layers after updateVelocity; the first %d (what the particle stages can read) before them, the rest "
                                            "on a second stream beside them and, between the frames of the timed call, beside the next level set "
                                            "(all inside the timed region: the last frame joins it before the closing event)"
                                            % (int(st.extrapolationLayers), int(getattr(st, "extrapolationNearLayers", 0))),
                           "pcg_iter_per_s": iters / (stage_ms[4] * 1e-3) if len(stage_ms) > 4 and stage_ms[4] > 0 else None,
                           "pcg_iteration_ms_kernels": pcg_ms / max(1, prof[0][1]),
                           "pcg_iteration_hbm_frac": (per_iter * cells_m / (pcg_ms / max(1, prof[0][1]) * 1e-3) / 1e9 / peak) if pcg_ms else None,
                           "stage_ms_last_step": stage_ms,
                           "step_dense_model_frac": dense_bytes / step_s / 1e9 / peak,
                           "step_dense_model_note": "SURVEY 8d byte model over ALL %d cells / step time / peak: a work-equivalent figure (the kernels skip exact zeros), not HBM utilisation" % cells_n,
                           "step_hbm_frac_touched": touched_bytes / step_s / 1e9 / peak,
                           "kernels": kinfo},
                "roofline": {"bound": "hbm", "kernel": KNAMES[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                             "algorithmic_bytes_per_launch": algo[dom] * cells_m, "peak_source": peak_src},
                "clocks": sampler.summary(), "gpu_launches": launches}
        if parity is not None:
            line["config"]["parity_checked"] = parity
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu:
            best, allr = cpu_reference_run(args.cpu_size, 2, 1, True)
            line["cpu_baseline"] = {"value": best[0], "unit": UNIT, "cores": best[2], "kind": best[1], "sample": best[3],
                                    "all_builds": [{"value": r[0], "cores": r[2], "sample": r[3]} for r in allr]}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sim.free()


if __name__ == "__main__":
    main()
