"""ctypes wrapper shared by the reference harness (oracle/_ref/*.so) and the C restatement
(oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's CPU legs and
__graft_entry__.smoke(); never by the product package.

Both libraries export the same `fso_*` C ABI (oracle/fsim_oracle.h).
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

# field ids (oracle/fsim_oracle.h, include/fsim.h)
U, V, NEWU, NEWV, P, CELL, PHI, PARTICLES, PARTICLE_VELS = range(9)
ADIAG, AX, AY, RHS, PRECON = 9, 10, 11, 12, 13
# stage ids = FluidSim2D::StageType (reference include/FluidSim2D.h:93-97)
ST_LEVELSET, ST_P2G, ST_SL_ADVECT, ST_GRAVITY, ST_SOLID_LS, ST_PROJECT, ST_UPDATE_VEL, ST_G2P, ST_ADVECT = range(1, 10)
SEMILAGRANGIAN, PICFLIP = 0, 1
EMPTY, FLUID, SOLID = 0, 1, 2

PATHS = {
    "port": os.path.join(ORACLE_DIR, "liboracle.so"),
    "ref": os.path.join(ORACLE_DIR, "_ref", "libfsim_ref.so"),
    "ref_patched": os.path.join(ORACLE_DIR, "_ref", "libfsim_ref_patched.so"),
    "ref_omp": os.path.join(ORACLE_DIR, "_ref", "libfsim_ref_omp.so"),
}


def build(targets=("port",)):
    """Compile the oracle libraries (the checker, not the product)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR] + list(targets), check=True)


def available(kind):
    return os.path.exists(PATHS[kind])


_libs = {}


def load(kind):
    if kind in _libs:
        return _libs[kind]
    path = PATHS[kind]
    if not os.path.exists(path):
        if kind == "port":
            build(("port",))
        else:
            raise FileNotFoundError(path)
    L = ctypes.CDLL(path)
    vp, ci, cd, cl = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_long
    L.fso_kind.restype = ctypes.c_char_p
    L.fso_create.restype = vp
    L.fso_create.argtypes = [ci, ci, ci, cd, cd, cd, cd, cd, ci, cd, vp]
    L.fso_destroy.argtypes = [vp]
    L.fso_num_particles.restype = cl
    L.fso_num_particles.argtypes = [vp]
    L.fso_get.argtypes = [vp, ci, vp]
    L.fso_set.argtypes = [vp, ci, vp]
    L.fso_set_particles.argtypes = [vp, cl, vp, vp]
    L.fso_stage.argtypes = [vp, ci]
    L.fso_step.argtypes = [vp, ci]
    L.fso_set_params.argtypes = [vp, cd, cd, cd, cd]
    L.fso_stat.restype = cd
    L.fso_stat.argtypes = [vp, ci]
    L.fso_stage_times.argtypes = [vp, vp, ci]
    L.fso_vel_interp.argtypes = [vp, cl, vp, vp]
    L.fso_set_pcg.argtypes = [cd, ci]
    L.fso_last_pcg_iters.argtypes = [vp]
    L.fso_set_sl_double_buffer.argtypes = [ci]
    _libs[kind] = L
    return L


class OracleSim:
    """One simulation instance behind the fso_* ABI (reference or port)."""

    def __init__(self, kind, cells, dt, dx, rho=997.0, gravity=(0.0, -9.81), mode=PICFLIP, alpha=0.05, ppc_sqrt=2):
        self.L = load(kind)
        self.kind = kind
        cells = np.ascontiguousarray(cells, dtype=np.uint8)
        self.ny, self.nx = cells.shape
        self.h = self.L.fso_create(self.nx, self.ny, ppc_sqrt, dt, dx, rho, gravity[0], gravity[1], mode, alpha,
                                   cells.ctypes.data)
        self.dx, self.dt = dx, dt

    def close(self):
        if self.h:
            self.L.fso_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_particles(self):
        return int(self.L.fso_num_particles(self.h))

    def shape(self, field):
        nx, ny = self.nx, self.ny
        if field in (U, NEWU):
            return (ny, nx + 1), np.float64
        if field in (V, NEWV):
            return (ny + 1, nx), np.float64
        if field == CELL:
            return (ny, nx), np.uint8
        if field in (PARTICLES, PARTICLE_VELS):
            return (self.num_particles, 2), np.float64
        return (ny, nx), np.float64

    def get(self, field):
        shp, dt = self.shape(field)
        a = np.zeros(shp, dtype=dt)
        if a.size and self.L.fso_get(self.h, field, a.ctypes.data) != 0:
            raise ValueError("field %d not available in %s" % (field, self.kind))
        return a

    def set(self, field, arr):
        shp, dt = self.shape(field)
        a = np.ascontiguousarray(arr, dtype=dt)
        assert a.shape == shp, (a.shape, shp)
        if a.size and self.L.fso_set(self.h, field, a.ctypes.data) != 0:
            raise ValueError("field %d not settable in %s" % (field, self.kind))

    def set_particles(self, pos, vel):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        vel = np.ascontiguousarray(vel, dtype=np.float64)
        assert pos.shape == vel.shape and pos.shape[1] == 2
        self.L.fso_set_particles(self.h, pos.shape[0], pos.ctypes.data, vel.ctypes.data)

    def stage(self, st):
        assert self.L.fso_stage(self.h, st) == 0

    def step(self, n=1):
        self.L.fso_step(self.h, n)

    def set_params(self, gx, gy, alpha, dt):
        self.L.fso_set_params(self.h, gx, gy, alpha, dt)

    def vel_interp(self, pos):
        """MACGrid2D::velInterp (reference include/MACGrid2D.h:96-98) at positions [n, 2]"""
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        out = np.zeros_like(pos)
        if len(pos):
            self.L.fso_vel_interp(self.h, len(pos), pos.ctypes.data, out.ctypes.data)
        return out

    def stat(self, which):
        return float(self.L.fso_stat(self.h, which))

    def stage_times(self):
        out = np.zeros(8, dtype=np.float32)
        n = self.L.fso_stage_times(self.h, out.ctypes.data, 8)
        return out[:n]

    @property
    def pcg_iters(self):
        return int(self.L.fso_last_pcg_iters(self.h))

    def state(self, fields=(U, V, NEWU, NEWV, P, CELL, PHI, PARTICLES, PARTICLE_VELS)):
        return {f: self.get(f) for f in fields}


def dam_break_cells(n, ny=None):
    """Scene of reference demo/App.cpp:147-160 in row-major cell[j, i]: solid border,
    FLUID where i + j < 3N/4, EMPTY elsewhere."""
    ny = n if ny is None else ny
    j, i = np.meshgrid(np.arange(ny), np.arange(n), indexing="ij")
    c = np.where(i + j < ny * 3 // 4, FLUID, EMPTY).astype(np.uint8)
    c[0, :] = SOLID
    c[-1, :] = SOLID
    c[:, 0] = SOLID
    c[:, -1] = SOLID
    return c


def rel_max(a, b):
    """max |a-b| / max |b| -- the parity metric of SURVEY.md section 8(c)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    fin = np.isfinite(b)
    if not np.array_equal(fin, np.isfinite(a)):
        return float("inf")
    if not np.array_equal(a[~fin], b[~fin]):
        return float("inf")
    den = np.abs(b[fin]).max() if fin.any() else 0.0
    num = np.abs(a[fin] - b[fin]).max() if fin.any() else 0.0
    return float(num / den) if den > 0 else float(num)
