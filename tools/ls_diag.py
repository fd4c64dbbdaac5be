"""Level-set diagnostic (not a test): createWaterLevelSet of the n x n dam break on the GPU against the oracle, from the
initial state (step 0) or after `steps` oracle steps copied to the GPU; prints where phi differs.
    python tools/ls_diag.py n [steps] [dt]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol
fs = importlib.import_module("fluid-sim_b200")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dt = float(sys.argv[3]) if len(sys.argv) > 3 else 0.005 * 128.0 / n
cells = ol.dam_break_cells(n)
dx = 1.28 / n
kind = "ref" if ol.available("ref") else "port"
o = ol.OracleSim(kind, cells, dt=dt, dx=dx, mode=ol.PICFLIP, alpha=0.05)
s = fs.FluidSim2D(cells, dt=dt, dx=dx, mode=fs.FS_PICFLIP, picFlipAlpha=0.05)
if steps:
    o.step(steps)
    for f in (ol.U, ol.V, ol.NEWU, ol.NEWV, ol.P, ol.CELL, ol.PHI):
        s.set(f, o.get(f))
    s.set_particles(o.get(ol.PARTICLES), o.get(ol.PARTICLE_VELS))
o.stage(ol.ST_LEVELSET)
s.profile_enable(True)
s.stage(fs.CREATE_WATER_LEVEL_SET)
print("closest-particle sweeps (ms):", " ".join("%.2f" % t for t in s.profile_list(5)))
print("eikonal sweeps (ms):", " ".join("%.2f" % t for t in s.profile_list(6)))
s.profile_enable(False)
a, b = s.get(fs.PHI), o.get(ol.PHI)
st = s.stats()
print("n %d steps %d env LEGACY=%s: sweeps run %d; labels equal %s" % (n, steps, os.environ.get("FSIM_LS_LEGACY"), st.levelSetSweeps,
      np.array_equal(s.get(fs.CELL), o.get(ol.CELL))))
fin = np.isfinite(b)
print("finite mask equal:", np.array_equal(np.isfinite(a), fin), "max|phi|", np.abs(b[fin]).max())
bad = np.argwhere(a != b)
print("cells that differ: %d of %d" % (len(bad), a.size))
if len(bad):
    jj, ii = bad[:, 0], bad[:, 1]
    print("  rows [%d, %d] cols [%d, %d]" % (jj.min(), jj.max(), ii.min(), ii.max()))
    d = np.abs(a - b)[a != b]
    print("  |diff| max %.3e median %.3e; sign(phi_ref) at diffs: neg %d pos %d" % (np.nanmax(d), np.nanmedian(d), (b[a != b] < 0).sum(), (b[a != b] >= 0).sum()))
    for (j, i) in bad[:12]:
        print("   (i=%d, j=%d) gpu %.17g ref %.17g" % (i, j, a[j, i], b[j, i]))
    hist = np.bincount(jj // 32)
    print("  differing cells per 32-row strip (nonzero):", {int(k): int(v) for k, v in enumerate(hist) if v})
