"""times the semi-Lagrangian mode (exact in-place order and the snapshot variant) at a few sizes"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, oracle_lib as ol
fs = importlib.import_module("fluid-sim_b200")
for n in (128, 1024, 4096):
    for db in (False, True):
        sim = fs.FluidSim2D(ol.dam_break_cells(n), dt=0.005 * 128.0 / n if n > 128 else 0.005, dx=1.28 / n, mode=fs.FS_SEMILAGRANGIAN, slDoubleBuffer=db)
        sim.update(2); sim.sync()
        t0 = time.perf_counter(); sim.update(3); sim.sync(); dt = (time.perf_counter() - t0) / 3
        st = sim.stats()
        print("n=%d doubleBuffer=%s: %.2f ms/step  stages %s iters %d" % (n, db, dt * 1e3, [round(x, 2) for x in st.stageMs[:st.numStages]], st.pcgIters), flush=True)
        sim.free()
