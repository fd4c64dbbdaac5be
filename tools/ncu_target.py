"""Workload for the ncu captures (tools/gpu_r2.sh): the 4096^2 PIC/FLIP dam break with the PCG capped at a few iterations, so
that one step launches every kernel class a handful of times.  Not a benchmark."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fs = importlib.import_module("fluid-sim_b200")
scenes = importlib.import_module("fluid-sim_b200.scenes")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sim = fs.FluidSim2D(scenes.dam_break_cells(n), mode=fs.FS_PICFLIP, picFlipAlpha=0.05, pcgMaxIters=cap, **scenes.dam_break_params(n))
import torch  # noqa: E402  (NVTX range for ncu --nvtx-include "capture/")

sim.update(steps - 1)
sim.sync()
torch.cuda.nvtx.range_push("capture")
sim.update(1)
sim.sync()
torch.cuda.nvtx.range_pop()
print("launches", sim.launch_count, "iters", sim.stats().pcgIters)
sim.free()
