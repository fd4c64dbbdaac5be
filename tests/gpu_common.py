"""Helpers shared by the GPU parity tests and tests/gpu_diag.py."""
import importlib
import os

import numpy as np

import oracle_lib as ol

fs = importlib.import_module("fluid-sim_b200")

NAMES = {ol.U: "u", ol.V: "v", ol.NEWU: "newu", ol.NEWV: "newv", ol.P: "p", ol.CELL: "cell", ol.PHI: "phi",
         ol.PARTICLES: "pos", ol.PARTICLE_VELS: "vel"}
GRID_FIELDS = (ol.U, ol.V, ol.NEWU, ol.NEWV, ol.P, ol.CELL, ol.PHI)
ALL_FIELDS = GRID_FIELDS + (ol.PARTICLES, ol.PARTICLE_VELS)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# stage-wise tolerances (max|a-b| / max|b|) from identical inputs; labels and phi must be bit-exact.
# The north_star bar is 1e-4 per step; these are what FP64 reordering actually needs.
STAGE_TOL = {
    ol.ST_LEVELSET: 0.0, ol.ST_P2G: 1e-11, ol.ST_SL_ADVECT: 1e-11, ol.ST_GRAVITY: 1e-13, ol.ST_SOLID_LS: 0.0,
    ol.ST_PROJECT: 1e-7, ol.ST_UPDATE_VEL: 1e-11, ol.ST_G2P: 1e-11, ol.ST_ADVECT: 1e-11,
}


def gpu_from_golden(g, **kw):
    nx, ny, ppc, mode = [int(x) for x in g["meta"]]
    dt, dx, rho, gx, gy, alpha = [float(x) for x in g["params"]]
    return fs.FluidSim2D(g["cells0"], dt=dt, dx=dx, rho=rho, gravity=(gx, gy), mode=mode, picFlipAlpha=alpha,
                         particlesPerCellSqrt=ppc, **kw)


def oracle_from_golden(kind, g):
    nx, ny, ppc, mode = [int(x) for x in g["meta"]]
    dt, dx, rho, gx, gy, alpha = [float(x) for x in g["params"]]
    return ol.OracleSim(kind, g["cells0"], dt=dt, dx=dx, rho=rho, gravity=(gx, gy), mode=mode, alpha=alpha,
                        ppc_sqrt=ppc)


def load_snapshot(sim, g, tag):
    for f in GRID_FIELDS:
        sim.set(f, g["%s_%s" % (tag, NAMES[f])])
    sim.set_particles(g["%s_pos" % tag], g["%s_vel" % tag])


def copy_state(dst, src_state):
    for f in GRID_FIELDS:
        dst.set(f, src_state[f])
    dst.set_particles(src_state[ol.PARTICLES], src_state[ol.PARTICLE_VELS])


def compare_states(got, want, tol):
    """returns list of (field, err, ok)"""
    out = []
    for f in want:
        if f == ol.CELL:
            bad = int((got[f] != want[f]).sum())
            out.append((NAMES[f], float(bad), bad == 0))
        elif tol == 0.0:
            eq = np.array_equal(got[f], want[f])
            out.append((NAMES[f], 0.0 if eq else ol.rel_max(got[f], want[f]), eq))
        else:
            e = ol.rel_max(got[f], want[f])
            out.append((NAMES[f], e, e <= tol))
    return out


def best_oracle():
    """the real reference when its prebuilt library travelled with the repo, else the pinned C restatement"""
    return "ref" if ol.available("ref") else "port"


def random_scene(nx, ny, seed, solid_block=True):
    rng = np.random.default_rng(seed)
    c = np.zeros((ny, nx), np.uint8)
    c[0, :] = c[-1, :] = ol.SOLID
    c[:, 0] = c[:, -1] = ol.SOLID
    w, h = int(nx * rng.uniform(0.25, 0.6)), int(ny * rng.uniform(0.3, 0.8))
    x0 = int(rng.integers(1, nx - w - 1))
    c[1:1 + h, x0:x0 + w] = ol.FLUID
    if solid_block:
        bx, by = int(rng.integers(2, nx - 8)), int(rng.integers(1, max(2, ny // 3)))
        c[by:by + 3, bx:bx + 5] = ol.SOLID
    return c
