"""CPU tests: pin the C restatement (oracle/fsim_oracle.c) against (a) the golden fixtures dumped from
the stock reference build and (b) the stock reference itself when oracle/_ref/ is present."""
import os

import numpy as np
import pytest

import oracle_lib as ol

NAMES = {ol.U: "u", ol.V: "v", ol.NEWU: "newu", ol.NEWV: "newv", ol.P: "p", ol.CELL: "cell", ol.PHI: "phi",
         ol.PARTICLES: "pos", ol.PARTICLE_VELS: "vel"}
GRID_FIELDS = (ol.U, ol.V, ol.NEWU, ol.NEWV, ol.P, ol.CELL, ol.PHI)


def load_snapshot(sim, g, tag):
    for f in GRID_FIELDS:
        sim.set(f, g["%s_%s" % (tag, NAMES[f])])
    sim.set_particles(g["%s_pos" % tag], g["%s_vel" % tag])


def make_sim(kind, g, cells=None):
    nx, ny, ppc, mode = [int(x) for x in g["meta"]]
    dt, dx, rho, gx, gy, alpha = [float(x) for x in g["params"]]
    return ol.OracleSim(kind, g["cells0"] if cells is None else cells, dt=dt, dx=dx, rho=rho, gravity=(gx, gy),
                        mode=mode, alpha=alpha, ppc_sqrt=ppc)


@pytest.mark.parametrize("fixture,float_tol", [("flip_stages_40x32.npz", 0.0), ("sl_stages_40x32.npz", 1e-12)])
def test_port_stagewise_vs_golden(golden_dir, fixture, float_tol):
    g = np.load(os.path.join(golden_dir, fixture))
    sim = make_sim("port", g)
    order = [int(x) for x in g["order"]]
    for k, st in enumerate(order):
        # every stage starts from the reference's own state, so errors cannot accumulate or cancel
        load_snapshot(sim, g, "s%d" % k)
        sim.stage(st)
        got = sim.state()
        for f, a in got.items():
            want = g["s%d_%s" % (k + 1, NAMES[f])]
            if f == ol.CELL or float_tol == 0.0:
                assert np.array_equal(a, want), "stage %d field %s differs" % (st, NAMES[f])
            else:
                assert ol.rel_max(a, want) <= float_tol, (st, NAMES[f], ol.rel_max(a, want))
        if st == ol.ST_LEVELSET:
            stats = np.array([sim.stat(0), sim.stat(1), sim.stat(2)])
            assert np.allclose(stats, g["s%d_stats" % (k + 1)], rtol=1e-13, atol=0)


@pytest.mark.parametrize("fixture,float_tol", [("flip_traj_64.npz", 0.0), ("sl_traj_64.npz", 1e-11)])
def test_port_trajectory_vs_golden(golden_dir, fixture, float_tol):
    g = np.load(os.path.join(golden_dir, fixture))
    sim = make_sim("port", g)
    # seeding uses glibc rand() with the implicit seed 1 (reference src/FluidSim2D.cpp:19-21, 52-64)
    assert np.array_equal(sim.get(ol.PARTICLES), g["pos0"])
    done = 0
    for upto in (1, 5, 25):
        sim.step(upto - done)
        done = upto
        assert np.array_equal(sim.get(ol.CELL), g["t%d_cell" % upto])
        for f in (ol.U, ol.V, ol.P, ol.PHI, ol.PARTICLES, ol.PARTICLE_VELS):
            want = g["t%d_%s" % (upto, NAMES[f])]
            got = sim.get(f)
            if float_tol == 0.0:
                assert np.array_equal(got, want), (upto, NAMES[f])
            else:
                assert ol.rel_max(got, want) <= float_tol, (upto, NAMES[f], ol.rel_max(got, want))
        # particle cell indices (int)(x/dx) must be exact
        dx = float(g["params"][1])
        assert np.array_equal((sim.get(ol.PARTICLES) / dx).astype(np.int32),
                              (g["t%d_pos" % upto] / dx).astype(np.int32))


def test_port_projection_internals_consistent(golden_dir):
    """A p = rhs on fluid cells to PCG tolerance, using the port's own A (sanity of the restated solver)."""
    g = np.load(os.path.join(golden_dir, "flip_stages_40x32.npz"))
    sim = make_sim("port", g)
    load_snapshot(sim, g, "s4")  # state right before applyProjection
    sim.stage(ol.ST_PROJECT)
    ad, ax, ay, rhs, p = (sim.get(f) for f in (ol.ADIAG, ol.AX, ol.AY, ol.RHS, ol.P))
    cell = sim.get(ol.CELL)
    ap = ad * p
    ap[:, 1:] += ax[:, :-1] * p[:, :-1]
    ap[:, :-1] += ax[:, :-1] * p[:, 1:]
    ap[1:, :] += ay[:-1, :] * p[:-1, :]
    ap[:-1, :] += ay[:-1, :] * p[1:, :]
    fl = cell == ol.FLUID
    assert np.abs((ap - rhs)[fl]).max() <= 1e-10 * np.abs(rhs).max()
    assert 0 < sim.pcg_iters < 200


needs_ref = pytest.mark.skipif(not ol.available("ref"), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
@pytest.mark.parametrize("mode", [ol.PICFLIP, ol.SEMILAGRANGIAN])
def test_port_vs_reference_live(mode):
    """Different scene from the fixtures: 48x36 with an interior solid block and a column of water."""
    nx, ny = 48, 36
    cells = np.zeros((ny, nx), np.uint8)
    cells[0, :] = cells[-1, :] = ol.SOLID
    cells[:, 0] = cells[:, -1] = ol.SOLID
    cells[1:24, 1:14] = ol.FLUID
    cells[1:6, 20:26] = ol.SOLID
    dx = 1.0 / nx
    a = ol.OracleSim("ref", cells, dt=0.004, dx=dx, mode=mode, alpha=0.05)
    b = ol.OracleSim("port", cells, dt=0.004, dx=dx, mode=mode, alpha=0.05)
    assert np.array_equal(a.get(ol.PARTICLES), b.get(ol.PARTICLES))
    for _ in range(15):
        a.step()
        b.step()
        assert np.array_equal(a.get(ol.CELL), b.get(ol.CELL))
    tol = 0.0 if mode == ol.PICFLIP else 1e-10
    for f in (ol.U, ol.V, ol.P, ol.PHI, ol.PARTICLES, ol.PARTICLE_VELS):
        assert ol.rel_max(b.get(f), a.get(f)) <= tol, NAMES[f]


@needs_ref
def test_reference_patched_build_is_stock_by_default():
    if not ol.available("ref_patched"):
        pytest.skip("patched build absent")
    cells = ol.dam_break_cells(32)
    a = ol.OracleSim("ref", cells, dt=0.005, dx=0.04)
    b = ol.OracleSim("ref_patched", cells, dt=0.005, dx=0.04)
    ol.load("ref_patched").fso_set_pcg(1e-12, 200)
    a.step(8)
    b.step(8)
    for f in (ol.U, ol.V, ol.P, ol.CELL, ol.PHI, ol.PARTICLES, ol.PARTICLE_VELS):
        assert np.array_equal(a.get(f), b.get(f))
    assert 0 < b.pcg_iters <= 200


@needs_ref
def test_port_pcg_iteration_count_matches_patched_reference():
    if not ol.available("ref_patched"):
        pytest.skip("patched build absent")
    cells = ol.dam_break_cells(64)
    a = ol.OracleSim("ref_patched", cells, dt=0.005, dx=0.02)
    b = ol.OracleSim("port", cells, dt=0.005, dx=0.02)
    for _ in range(5):
        a.step()
        b.step()
        assert a.pcg_iters == b.pcg_iters


@needs_ref
def test_patched_reference_exports_projection_locals():
    """applyProjection's locals Adiag/Ax/Ay/rhs/precon (reference src/FluidSim2D.cpp:253-258, 334, 366) copied out of
    the patched reference build equal the port's, bit for bit, and satisfy A p = rhs with the reference's own p"""
    if not ol.available("ref_patched"):
        pytest.skip("patched build absent")
    cells = ol.dam_break_cells(64)
    a = ol.OracleSim("ref_patched", cells, dt=0.005, dx=0.02)
    b = ol.OracleSim("port", cells, dt=0.005, dx=0.02)
    a.step(3)
    b.step(3)
    for f in (ol.ADIAG, ol.AX, ol.AY, ol.RHS, ol.PRECON):
        assert np.array_equal(a.get(f), b.get(f)), f
    ad, ax, ay, rhs, p = (a.get(f) for f in (ol.ADIAG, ol.AX, ol.AY, ol.RHS, ol.P))
    ap = ad * p
    ap[:, 1:] += ax[:, :-1] * p[:, :-1]
    ap[:, :-1] += ax[:, :-1] * p[:, 1:]
    ap[1:, :] += ay[:-1, :] * p[:-1, :]
    ap[:-1, :] += ay[:-1, :] * p[1:, :]
    fl = a.get(ol.CELL) == ol.FLUID
    # (the labels are relabelled by the NEXT step's level set only, so `fl` is the set the projection used)
    assert np.abs((ap - rhs)[fl]).max() <= 1e-9 * np.abs(rhs).max()


@needs_ref
def test_port_diagnostics_match_reference_methods():
    """avgPressure / avgPressureInFluid / maxVelocity (reference src/FluidSim2D.cpp:607-638): the port's raster-order loops
    against the reference's own methods, bit for bit"""
    cells = ol.dam_break_cells(48)
    a = ol.OracleSim("ref", cells, dt=0.005, dx=0.02)
    b = ol.OracleSim("port", cells, dt=0.005, dx=0.02)
    a.step(6)
    b.step(6)
    for which in (4, 5, 6):
        assert a.stat(which) == b.stat(which), which
    assert a.stat(6) > 0.0


@needs_ref
def test_renderer_staging_restatement_port_vs_reference():
    """oracle/render_oracle.py (numpy restatement of FluidRenderer2D::updateBuffers, demo/FluidRenderer2D.cpp:435-486) gives the same
    arrays whether it is driven by the reference build (its own mac.velInterp) or by the C port"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import render_oracle
    cells = ol.dam_break_cells(40)
    a = ol.OracleSim("ref", cells, dt=0.005, dx=0.02)
    b = ol.OracleSim("port", cells, dt=0.005, dx=0.02)
    a.step(5)
    b.step(5)
    ra, rb = render_oracle.update_buffers(a), render_oracle.update_buffers(b)
    for k in ra:
        assert np.array_equal(ra[k], rb[k], equal_nan=True), k
    ncell = int((a.get(ol.CELL) == ol.FLUID).sum())
    assert len(ra["water"]) == ncell and len(ra["solid"]) == 2 * 40 + 2 * 38
    assert len(ra["cellVels"]) == 2 * 39 * 39 and len(ra["particleVelLines"]) == 2 * a.num_particles
    assert ((ra["phiValues"] >= 0) & (ra["phiValues"] <= 1)).all()
