"""The drop-in C++ shim (fluid-sim_b200/shim/FluidSim2D_b200.cpp, compiled against the reference's own headers into
oracle/_ref/shim_demo where /root/reference exists) makes the same calls the reference's demo makes --
FluidSim2D::create, update() per frame -- and must leave the same public fields behind as the stock reference."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
DEMO = os.path.join(ol.ORACLE_DIR, "_ref", "shim_demo")


@pytest.mark.parametrize("mode", [ol.PICFLIP, ol.SEMILAGRANGIAN])
def test_shim_matches_reference(mode):
    if not (os.path.exists(DEMO) and ol.available("ref")):
        pytest.skip("shim_demo / reference library were not prebuilt (needs /root/reference at build time)")
    n, steps = 64, 4
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "state.bin")
        r = subprocess.run([DEMO, str(n), str(mode), str(steps), out], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr + r.stdout
        raw = open(out, "rb").read()
    npart = int(np.frombuffer(raw[:8], dtype=np.int64)[0])
    off = 8

    def take(count, dtype):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a

    u = take((n + 1) * n, np.float64).reshape(n, n + 1)
    v = take(n * (n + 1), np.float64).reshape(n + 1, n)
    p = take(n * n, np.float64).reshape(n, n)
    phi = take(n * n, np.float64).reshape(n, n)
    cell = take(n * n, np.uint8).reshape(n, n)
    pos = take(2 * npart, np.float64).reshape(npart, 2)
    vel = take(2 * npart, np.float64).reshape(npart, 2)

    o = ol.OracleSim("ref", ol.dam_break_cells(n), dt=0.005, dx=1.28 / n, mode=mode, alpha=0.05)
    o.step(steps)
    assert npart == o.num_particles
    assert np.array_equal(cell, o.get(ol.CELL))
    for got, f in ((u, ol.U), (v, ol.V), (p, ol.P), (phi, ol.PHI), (pos, ol.PARTICLES), (vel, ol.PARTICLE_VELS)):
        assert ol.rel_max(got, o.get(f)) <= 1e-6, f
