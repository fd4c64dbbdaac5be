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


def test_renderer_staging_arrays_from_device_state():
    """fsim_render_fill = FluidRenderer2D::updateBuffers (reference demo/FluidRenderer2D.cpp:435-486) from the
    device-resident state, against the numpy restatement in oracle/render_oracle.py driven by the reference's own velInterp:
    list lengths and cell locations exact (raster order), interpolated and sigmoid values to float rounding"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import render_oracle
    n = 96
    cells = ol.dam_break_cells(n)
    kw = dict(dt=0.005, dx=1.28 / n)
    o = ol.OracleSim(best_oracle(), cells, mode=ol.PICFLIP, alpha=0.05, **kw)
    s = fs.FluidSim2D(cells, mode=fs.FS_PICFLIP, picFlipAlpha=0.05, **kw)
    for _ in range(6):
        o.step(); s.update()
    want = render_oracle.update_buffers(o)
    got = s.render_buffers()
    for k in ("water", "solid", "pressureCells"):
        assert got[k].shape == want[k].shape, (k, got[k].shape, want[k].shape)
        assert np.array_equal(got[k], want[k]), k
    assert len(got["water"]) == int((s.get(ol.CELL) == ol.FLUID).sum()) > 0
    for k in ("cellVels", "particleVelLines", "pressureValues", "phiValues"):
        assert got[k].shape == want[k].shape, (k, got[k].shape, want[k].shape)
        fin = np.isfinite(want[k])
        assert np.array_equal(fin, np.isfinite(got[k])), k
        err = np.abs(got[k][fin] - want[k][fin]).max() / max(1e-30, np.abs(want[k][fin]).max())
        assert err <= 2e-6, (k, err)
    # an undersized list is refused, not overrun
    io = fs.FsimRenderStaging()
    small = np.zeros((4, 2), np.float32)
    io.waterCells, io.waterCap = small.ctypes.data, 4
    rc = fs.lib().fsim_render_fill(s._h, __import__("ctypes").byref(io))
    assert rc != 0 and io.nWater == len(want["water"])
    s.free(); o.close()
