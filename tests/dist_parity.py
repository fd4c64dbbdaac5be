"""Multi-GPU parity against the ORACLE (test infrastructure; used by tools/dist_check.py, tests/test_gpu_dist.py and, as the
checker only, by bench.py --gpus N before its timed region).

Every rank calls slab_parity_vs_oracle(); rank 0 runs the reference (oracle/_ref patched build, or the C restatement)
on the same inputs and compares:
  (1) at a stop rule both can meet -- BASELINE config 3b in miniature: free surface (rows j >= 3N/4 EMPTY), SplitMix64 face
      velocities, PCG to a relative residual of 1e-6, cap 10000 -- the N-rank pressure against the reference's, plus the
      iteration delta that block-MIC(0) costs (north_star: "stated and reported separately").  (3b, not the closed tank 3a:
      with solid walls all round the pressure is only defined up to a constant, and a different preconditioner converges
      to a different member of that family.)
  (2) at the STOCK constants (tol 1e-12, cap 200) on the dam break: how far the N-rank fields are from the reference's
      after one update() -- a number, not an assertion: with block-MIC(0) the slabs need more iterations than one GPU, so
      wherever the cap binds the N-rank result is a less converged pressure than the reference's own.
"""
import importlib

import numpy as np

import oracle_lib as ol

fs = importlib.import_module("fluid-sim_b200")
scenes = importlib.import_module("fluid-sim_b200.scenes")


def slab_parity_vs_oracle(join, rank, world, device, n=512, stock_n=512, verbose=False):
    """join(sim) must call sim.dist_init(rank, world, id) with an id shared by all ranks.  Returns a dict (all ranks)."""
    out = {"world": world}
    kind = "ref_patched" if ol.available("ref_patched") else "port"
    # ---- (1) projection only, cap 10000: at the tolerance BASELINE config 3 names (1e-6) and at 1e-10, where "both
    # converged" pins the pressure itself (two solves stopped at 1e-6 may differ by cond(A) * 1e-6 in p) -------------------
    cells, phi, u, v, dx = scenes.projection_stress(n, "3b")
    out["projection"] = []
    for tol in (1e-6, 1e-10):
        sim = fs.FluidSim2D(cells, dt=dx, dx=dx, pcgTol=tol, pcgMaxIters=10000, seedParticles=False, computeStats=False, device=device)
        join(sim)
        sim.set(fs.U, u); sim.set(fs.V, v); sim.set(fs.PHI, phi)
        sim.applyProjection()
        st = sim.stats()
        p = sim.get(fs.P)
        rec = {"size": n, "tol": tol, "iters_slabs": int(st.pcgIters), "relative_residual": float(st.pcgResidual / st.pcgRhsNorm),
               "dist_error": int(getattr(st, "distError", 0))}
        sim.free()
        if rank == 0:
            L = ol.load(kind)
            L.fso_set_pcg(tol, 10000)
            try:
                o = ol.OracleSim(kind, cells, dt=dx, dx=dx)
                o.set(ol.U, u); o.set(ol.V, v); o.set(ol.PHI, phi)
                o.stage(ol.ST_PROJECT)
                rec.update({"iters_reference": int(o.pcg_iters), "p_rel_max_err": ol.rel_max(p, o.get(ol.P)), "oracle": kind})
                o.close()
            finally:
                L.fso_set_pcg(1e-12, 200)
        out["projection"].append(rec)
    # ---- (2) stock constants, one update() of the dam break --------------------------------------------------------
    cells = scenes.dam_break_cells(stock_n)
    kw = dict(dt=0.005, dx=1.28 / stock_n)
    sim = fs.FluidSim2D(cells, mode=fs.FS_PICFLIP, picFlipAlpha=0.05, device=device, **kw)
    join(sim)
    sim.update()
    st = sim.stats()
    got = {f: sim.get(f) for f in (fs.U, fs.V, fs.P, fs.CELL)}
    out["stock_cap"] = {"size": stock_n, "iters_slabs": int(st.pcgIters), "hit_cap": int(st.pcgHitMaxIters),
                        "relative_residual": float(st.pcgResidual / st.pcgRhsNorm)}
    sim.free()
    if rank == 0:
        o = ol.OracleSim(kind, cells, mode=ol.PICFLIP, alpha=0.05, **kw)
        o.step()
        out["stock_cap"].update({"iters_reference": int(o.pcg_iters), "labels_equal": bool(np.array_equal(got[fs.CELL], o.get(ol.CELL))),
                                 "p_rel_max_err": ol.rel_max(got[fs.P], o.get(ol.P)), "u_rel_max_err": ol.rel_max(got[fs.U], o.get(ol.U)),
                                 "v_rel_max_err": ol.rel_max(got[fs.V], o.get(ol.V))})
        o.close()
    if verbose and rank == 0:
        print("slab parity vs oracle:", out, flush=True)
    return out
