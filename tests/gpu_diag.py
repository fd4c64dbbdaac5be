"""Stage-by-stage diagnostic on the GPU (not a pytest file): prints a table of errors of every stage against the
golden fixtures, with the fast wavefront scheduler and with the debug single-CTA scheduler, then short
trajectories.  Usage: python tests/gpu_diag.py [--quick]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle_lib as ol  # noqa: E402
from gpu_common import (ALL_FIELDS, GOLDEN, NAMES, STAGE_TOL, best_oracle, compare_states, copy_state, fs,  # noqa: E402
                        gpu_from_golden, load_snapshot)

report = {}


def stagewise(fixture, debug):
    g = np.load(os.path.join(GOLDEN, fixture))
    sim = gpu_from_golden(g, debugSimpleWavefront=debug)
    order = [int(x) for x in g["order"]]
    for k, st in enumerate(order):
        load_snapshot(sim, g, "s%d" % k)
        t0 = time.time()
        try:
            sim.stage(st)
            sim.sync()
        except Exception as e:  # noqa: BLE001
            print("  stage %d raised: %s" % (st, e))
            report["%s/%d/%d" % (fixture, debug, st)] = "EXC %s" % e
            return
        got = sim.state()
        want = {f: g["s%d_%s" % (k + 1, NAMES[f])] for f in ALL_FIELDS}
        res = compare_states(got, want, STAGE_TOL[st])
        line = " ".join("%s=%s%.1e" % (n, "" if ok else "!", e) for n, e, ok in res)
        stt = sim.stats()
        print("  [%s dbg=%d] stage %d (%.0f ms) iters=%d sweeps=%d layers=%d : %s" %
              (fixture[:4], debug, st, (time.time() - t0) * 1e3, stt.pcgIters, stt.levelSetSweeps,
               stt.extrapolationLayers, line))
        report["%s/%d/%d" % (fixture, debug, st)] = {n: [e, bool(ok)] for n, e, ok in res}
    sim.free()


def trajectory(n, mode, steps, debug=0):
    kind = best_oracle()
    cells = ol.dam_break_cells(n)
    dx = 1.28 / n
    o = ol.OracleSim(kind, cells, dt=0.005 * min(1.0, 128.0 / n), dx=dx, mode=mode, alpha=0.05)
    s = fs.FluidSim2D(cells, dt=o.dt, dx=dx, mode=mode, picFlipAlpha=0.05, debugSimpleWavefront=debug)
    same_seed = np.array_equal(s.get(ol.PARTICLES), o.get(ol.PARTICLES))
    print("  traj %d mode=%d oracle=%s seeded-identically=%s np=%d" % (n, mode, kind, same_seed, s.num_particles))
    for k in range(steps):
        t0 = time.time(); o.step(); t1 = time.time(); s.update(); s.sync(); t2 = time.time()
        got, want = s.state(), o.state()
        res = compare_states(got, want, 1e-6)
        st = s.stats()
        print("   step %2d cpu %.0f ms gpu %.0f ms iters %d/%d: %s" % (
            k, (t1 - t0) * 1e3, (t2 - t1) * 1e3, st.pcgIters, o.pcg_iters,
            " ".join("%s=%s%.1e" % (nm, "" if ok else "!", e) for nm, e, ok in res)))
        report["traj/%d/%d/%d" % (n, mode, k)] = {nm: [e, bool(ok)] for nm, e, ok in res}
    print("   stage ms:", [round(float(x), 3) for x in st.stageMs[:st.numStages]])


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    print(fs.lib().fsim_version().decode())
    for dbg in (1, 0):
        stagewise("flip_stages_40x32.npz", dbg)
    for dbg in (1, 0):
        stagewise("sl_stages_40x32.npz", dbg)
    trajectory(64, ol.PICFLIP, 5)
    trajectory(128, ol.PICFLIP, 3)
    trajectory(64, ol.SEMILAGRANGIAN, 3)
    if not quick:
        trajectory(256, ol.PICFLIP, 2)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(report, open("gpurun_out/diag.json", "w"), indent=1)
