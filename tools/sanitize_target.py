"""Small run of every public entry point for compute-sanitizer (tools/gpu_sanitize.sh)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
fs = importlib.import_module("fluid-sim_b200")
scenes = importlib.import_module("fluid-sim_b200.scenes")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
for mode in (fs.FS_PICFLIP, fs.FS_SEMILAGRANGIAN):
    sim = fs.FluidSim2D(scenes.dam_break_cells(n, n - 24), mode=mode, picFlipAlpha=0.05, dt=0.005, dx=1.28 / n, pcgMaxIters=40)
    sim.update(3)
    print(mode, "diag", sim._diag(), "iters", sim.stats().pcgIters)
    r = sim.render_buffers()
    print("  render", {k: v.shape for k, v in r.items()})
    sim.save_checkpoint("/tmp/sanitize.ckp")
    b = fs.FluidSim2D.load_checkpoint("/tmp/sanitize.ckp", pcgMaxIters=40)
    b.update(1)
    b.free(); sim.free()
sim = fs.FluidSim2D(scenes.dam_break_cells(96), mode=fs.FS_PICFLIP, dt=0.005, dx=1.28 / 96, reserved=[0, 0, 0, 0, 0, 0, 0, 1])
sim.update(2); sim.free()
print("done")
