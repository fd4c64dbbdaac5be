"""the headless driver (fluid-sim_b200/bin/fsim_run), its perf.csv / conservation.csv in the
reference's formats (src/PerformanceCounter.cpp:58-74, src/FluidSim2D.cpp:738-750) and the state checkpoints."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from gpu_common import ALL_FIELDS, NAMES, best_oracle, fs

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FSIM_RUN = os.path.join(ROOT, "fluid-sim_b200", "bin", "fsim_run")


@pytest.mark.parametrize("mode", [fs.FS_PICFLIP, fs.FS_SEMILAGRANGIAN])
def test_checkpoint_roundtrip_continues_bit_for_bit(tmp_path, mode):
    n = 96
    cells = ol.dam_break_cells(n)
    a = fs.FluidSim2D(cells, dt=0.005, dx=1.28 / n, mode=mode, picFlipAlpha=0.05)
    a.update(3)
    path = str(tmp_path / "state.ckp")
    a.save_checkpoint(path)
    expect_bytes = 88 + 8 * (2 * (n + 1) * n * 2 + 2 * n * n) + n * n + 32 * a.num_particles
    assert os.path.getsize(path) == expect_bytes
    a.update(3)
    b = fs.FluidSim2D.load_checkpoint(path)
    assert b.num_particles == a.num_particles and (b.sizeX, b.sizeY, b.mode) == (n, n, mode)
    b.update(3)
    for f in ALL_FIELDS:
        assert np.array_equal(a.get(f), b.get(f)), NAMES[f]
    assert a.stats().currentTime == b.stats().currentTime
    a.free(); b.free()
    with open(path, "r+b") as fh:  # a damaged file is refused, not half-loaded
        fh.truncate(expect_bytes - 100)
    with pytest.raises(fs.FsimError):
        fs.FluidSim2D.load_checkpoint(path)


def _read_conservation(path):
    lines = open(path).read().splitlines()
    assert lines[0] == "Total Volume, Total Energy, Total Energy (Particle) "
    return np.array([[float(x) for x in ln.split(",")] for ln in lines[1:]])


def test_headless_driver_writes_the_reference_csv_formats(tmp_path):
    assert os.path.exists(FSIM_RUN), "fluid-sim_b200/bin/fsim_run is not built"
    n, steps = 128, 40
    out = str(tmp_path)
    ck = os.path.join(out, "s.ckp")
    r = subprocess.run([FSIM_RUN, "--size", str(n), "--steps", str(steps), "--mode", "flip", "--out", out, "--save", ck],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    cons = _read_conservation(os.path.join(out, "conservation.csv"))
    assert cons.shape == (steps, 3)
    # the same scene in the oracle (demo scene: dx 0.01, dt 0.005, alpha 0.05)
    o = ol.OracleSim(best_oracle(), ol.dam_break_cells(n), dt=0.005, dx=0.01, mode=ol.PICFLIP, alpha=0.05)
    want = []
    for _ in range(steps):
        o.step()
        want.append([o.stat(0), o.stat(1), o.stat(2)])
    want = np.array(want)
    assert np.allclose(cons, want, rtol=1e-6, atol=2e-6), np.abs(cons - want).max()
    # perf.csv: rows only after the 30-frame ring has filled, 8 stages (PIC/FLIP), every value followed by a comma
    rows = open(os.path.join(out, "perf.csv")).read().splitlines()
    assert len(rows) == steps - 30
    for ln in rows:
        assert ln.endswith(",")
        vals = [float(x) for x in ln[:-1].split(",")]
        assert len(vals) == 8 and all(v >= 0.0 for v in vals) and sum(vals) > 0.0
    # resume from the checkpoint: the next 5 steps continue the oracle's trajectory
    out2 = str(tmp_path / "resumed")
    os.makedirs(out2)
    r = subprocess.run([FSIM_RUN, "--load", ck, "--steps", "5", "--out", out2, "--quiet"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    cons2 = _read_conservation(os.path.join(out2, "conservation.csv"))
    want2 = []
    for _ in range(5):
        o.step()
        want2.append([o.stat(0), o.stat(1), o.stat(2)])
    assert np.allclose(cons2, np.array(want2), rtol=1e-6, atol=2e-6)
    assert open(os.path.join(out2, "perf.csv")).read() == ""  # fewer than 30 frames: no rows, like the reference
    o.close()
