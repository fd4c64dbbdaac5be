"""TEST INFRASTRUCTURE ONLY.  numpy restatement of FluidRenderer2D::updateBuffers (reference demo/FluidRenderer2D.cpp:435-486),
the per-frame staging arrays of the reference's renderer, computed from an oracle simulation (tests/oracle_lib.OracleSim).
The renderer itself needs SDL2 + OpenGL and cannot be compiled here; everything numeric it calls -- mac.velInterp -- is the
reference's own code behind fso_vel_interp.  Parity is pinned only as far as that: the float conversions and the sigmoid are
restated from the cited lines."""
import numpy as np

U, V, P, CELL, PHI, PARTICLES = 0, 1, 4, 5, 6, 7
FLUID, SOLID = 1, 2


def sigmoid_f32(x):
    """aml::sigmoid<float> (deps/altmath/src/math_utils.h:76-79): (T)1 / ((T)1 + exp(-x)) on a float argument"""
    x = np.asarray(x, dtype=np.float32)
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(np.float32)


def update_buffers(o):
    nx, ny, dx, dt = o.nx, o.ny, o.dx, o.dt
    cell, p, phi, pos = o.get(CELL), o.get(P), o.get(PHI), o.get(PARTICLES)
    out = {}
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")  # raster order of FluidSim2D::iterate: j outer, i inner
    loc = np.stack([(ii * dx).astype(np.float32), (jj * dx).astype(np.float32)], axis=-1)  # :442, :445 (float)(i * dx)
    out["water"] = loc[cell == FLUID]
    out["solid"] = loc[cell == SOLID]
    # :449-459
    j2, i2 = np.meshgrid(np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
    cpos = np.stack([((i2 + 0.5) * dx).ravel(), ((j2 + 0.5) * dx).ravel()], axis=-1)
    vel = o.vel_interp(cpos)
    cv = np.zeros((len(cpos) * 2, 2), np.float32)
    cv[0::2] = cpos.astype(np.float32)
    cv[1::2] = (cpos + vel * dt).astype(np.float32)
    out["cellVels"] = cv
    # :460-470
    nz = p != 0
    out["pressureCells"] = loc[nz]
    out["pressureValues"] = sigmoid_f32(np.float32(0.01) * p[nz].astype(np.float32))
    # :471-479 (float arithmetic)
    pv = o.vel_interp(pos)
    lines = np.zeros((2 * len(pos), 2), np.float32)
    lines[0::2] = pos.astype(np.float32)
    lines[1::2] = pos.astype(np.float32) + np.float32(dt) * pv.astype(np.float32)
    out["particleVelLines"] = lines
    # :480-485
    out["phiValues"] = sigmoid_f32((100.0 * phi).astype(np.float32)).ravel()
    return out
