import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, oracle_lib as ol
fs = importlib.import_module("fluid-sim_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sim = fs.FluidSim2D(ol.dam_break_cells(n), dt=0.005 * 128.0 / n, dx=1.28 / n, mode=fs.FS_PICFLIP, picFlipAlpha=0.05)
sim.update(2); sim.sync()
sim.profile_enable(True)
sim.update(1); sim.sync()
st = sim.stats()
print("layers", st.extrapolationLayers, "stage ms", [round(x, 2) for x in st.stageMs[:st.numStages]])
for k in (7, 9):
    print(k, sim.profile_get(k))
d = sim.get(fs.U); print("u finite", np.isfinite(d).all())
