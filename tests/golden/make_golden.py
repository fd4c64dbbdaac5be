"""Regenerates the golden fixtures in this directory from the STOCK reference build
(oracle/_ref/libfsim_ref.so = lasagnaphil/fluid-sim @ 29962de compiled from /root/reference by
oracle/Makefile).  Run in the build container only (the GPU box has no /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

Fixtures (all little-endian float64 / uint8, numpy .npz):
  flip_stages_40x32.npz  full public state before a step and after each of the 8 PIC/FLIP stages
                         (FluidSim2D::runFrame, reference src/FluidSim2D.cpp:116-133), taken after 10
                         warm-up steps of a 40x32 dam break (non-square on purpose).
  sl_stages_40x32.npz    the same for the 7 semi-Lagrangian stages (:100-115).
  flip_traj_64.npz       64x64 demo dam break (demo/App.cpp:147-160), state after 1, 5 and 25 steps,
                         plus the initial particles the reference seeded with glibc rand() (seed 1).
  sl_traj_64.npz         semi-Lagrangian counterpart (state after 1, 5, 25 steps).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

NAMES = {ol.U: "u", ol.V: "v", ol.NEWU: "newu", ol.NEWV: "newv", ol.P: "p", ol.CELL: "cell", ol.PHI: "phi",
         ol.PARTICLES: "pos", ol.PARTICLE_VELS: "vel"}
FLIP_ORDER = [ol.ST_LEVELSET, ol.ST_P2G, ol.ST_GRAVITY, ol.ST_SOLID_LS, ol.ST_PROJECT, ol.ST_UPDATE_VEL, ol.ST_G2P,
              ol.ST_ADVECT]
SL_ORDER = [ol.ST_LEVELSET, ol.ST_SL_ADVECT, ol.ST_GRAVITY, ol.ST_SOLID_LS, ol.ST_PROJECT, ol.ST_UPDATE_VEL,
            ol.ST_ADVECT]


def snap(sim, out, tag):
    for f, a in sim.state().items():
        out["%s_%s" % (tag, NAMES[f])] = a
    out["%s_stats" % tag] = np.array([sim.stat(0), sim.stat(1), sim.stat(2)])


def stages(mode, order, path):
    nx, ny = 40, 32
    cells = ol.dam_break_cells(nx, ny)
    dx = 1.28 / nx
    sim = ol.OracleSim("ref", cells, dt=0.005, dx=dx, mode=mode, alpha=0.05)
    out = {"meta": np.array([nx, ny, 2, mode], dtype=np.int64), "params": np.array([0.005, dx, 997.0, 0.0, -9.81, 0.05]),
           "order": np.array(order, dtype=np.int64), "cells0": cells}
    sim.step(10)
    snap(sim, out, "s0")
    for k, st in enumerate(order):
        sim.stage(st)
        snap(sim, out, "s%d" % (k + 1))
    np.savez_compressed(path, **out)
    sim.close()


def traj(mode, path):
    n = 64
    cells = ol.dam_break_cells(n)
    dx = 1.28 / n
    sim = ol.OracleSim("ref", cells, dt=0.005, dx=dx, mode=mode, alpha=0.05)
    out = {"meta": np.array([n, n, 2, mode], dtype=np.int64), "params": np.array([0.005, dx, 997.0, 0.0, -9.81, 0.05]),
           "cells0": cells, "pos0": sim.get(ol.PARTICLES)}
    done = 0
    for upto in (1, 5, 25):
        sim.step(upto - done)
        done = upto
        for f in (ol.U, ol.V, ol.P, ol.CELL, ol.PHI, ol.PARTICLES, ol.PARTICLE_VELS):
            out["t%d_%s" % (upto, NAMES[f])] = sim.get(f)
    np.savez_compressed(path, **out)
    sim.close()


if __name__ == "__main__":
    stages(ol.PICFLIP, FLIP_ORDER, os.path.join(HERE, "flip_stages_40x32.npz"))
    stages(ol.SEMILAGRANGIAN, SL_ORDER, os.path.join(HERE, "sl_stages_40x32.npz"))
    traj(ol.PICFLIP, os.path.join(HERE, "flip_traj_64.npz"))
    traj(ol.SEMILAGRANGIAN, os.path.join(HERE, "sl_traj_64.npz"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
