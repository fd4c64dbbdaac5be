"""fluid-sim_b200: B200-native (sm_100a) hot path of lasagnaphil/fluid-sim behind the reference's own
FluidSim2D interface.

This module is the thin Python mirror used by tests/ and bench.py; the product is the C ABI in
include/fsim.h (lib/libfsim_b200.so) and the C++14 drop-in translation unit shim/FluidSim2D_b200.cpp of this
package (compiled against the reference's own include/FluidSim2D.h in place of its src/FluidSim2D.cpp).
Method names follow the reference's public methods (reference include/FluidSim2D.h:116-150).

There is no CPU fallback: importing works anywhere (so symbol checks can run), but creating a simulation
without the CUDA library or without a GPU raises.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfsim_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fsim.h")

# field ids (include/fsim.h)
U, V, NEWU, NEWV, P, CELL, PHI, PARTICLES, PARTICLE_VELS = range(9)
ADIAG, AX, AY, RHS, PRECON = 9, 10, 11, 12, 13
# stage ids = FluidSim2D::StageType (reference include/FluidSim2D.h:93-97)
(CREATE_WATER_LEVEL_SET, TRANSFER_VELOCITY_TO_GRID, APPLY_SEMI_LAGRANGIAN_ADVECTION, APPLY_GRAVITY,
 CREATE_SOLID_LEVEL_SET, APPLY_PROJECTION, UPDATE_VELOCITY, UPDATE_PARTICLE_VELOCITIES, APPLY_ADVECTION) = range(1, 10)
FS_SEMILAGRANGIAN, FS_PICFLIP = 0, 1
FS_EMPTY, FS_FLUID, FS_SOLID = 0, 1, 2

EXPORTS = [
    "fsim_default_options", "fsim_create", "fsim_destroy", "fsim_step", "fsim_stage", "fsim_sync",
    "fsim_num_particles", "fsim_upload", "fsim_download", "fsim_set_particles", "fsim_set_params", "fsim_set_pcg",
    "fsim_get_stats", "fsim_step_host", "fsim_launch_count", "fsim_profile_enable", "fsim_profile_get", "fsim_last_error", "fsim_version",
    "fsim_dist_unique_id", "fsim_dist_init", "fsim_host_register", "fsim_host_unregister",
    "fsim_step_timed", "fsim_diagnostics", "fsim_checkpoint_save", "fsim_checkpoint_load", "fsim_profile_list", "fsim_render_fill",
]


class FsimConfig(ctypes.Structure):
    _fields_ = [("sizeX", ctypes.c_int), ("sizeY", ctypes.c_int), ("particlesPerCellSqrt", ctypes.c_int),
                ("dt", ctypes.c_double), ("dx", ctypes.c_double), ("rho", ctypes.c_double),
                ("gravityX", ctypes.c_double), ("gravityY", ctypes.c_double), ("mode", ctypes.c_int),
                ("picFlipAlpha", ctypes.c_double), ("initialValues", ctypes.c_void_p)]


class FsimOptions(ctypes.Structure):
    _fields_ = [("pcgTol", ctypes.c_double), ("pcgMaxIters", ctypes.c_int), ("device", ctypes.c_int),
                ("seedParticles", ctypes.c_int), ("computeStats", ctypes.c_int), ("slDoubleBuffer", ctypes.c_int),
                ("debugSimpleWavefront", ctypes.c_int), ("reserved", ctypes.c_int * 8)]


class FsimStats(ctypes.Structure):
    _fields_ = [("waterVolume", ctypes.c_double), ("totalEnergy", ctypes.c_double),
                ("particleTotalEnergy", ctypes.c_double), ("currentTime", ctypes.c_double),
                ("pcgIters", ctypes.c_int), ("pcgHitMaxIters", ctypes.c_int), ("pcgResidual", ctypes.c_double),
                ("pcgRhsNorm", ctypes.c_double), ("cflMax", ctypes.c_double), ("nanPositions", ctypes.c_int),
                ("levelSetSweeps", ctypes.c_int), ("extrapolationLayers", ctypes.c_int),
                ("stageMs", ctypes.c_float * 8), ("numStages", ctypes.c_int), ("pcgSolveCells", ctypes.c_longlong), ("pcgMarchedCells", ctypes.c_longlong),
                ("distError", ctypes.c_int), ("extrapolationNearLayers", ctypes.c_int)]


class FsimHostMirror(ctypes.Structure):
    _fields_ = [("u_in", ctypes.c_void_p), ("v_in", ctypes.c_void_p), ("u", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("p", ctypes.c_void_p), ("cell", ctypes.c_void_p), ("phi", ctypes.c_void_p),
                ("particles", ctypes.c_void_p), ("particleVels", ctypes.c_void_p)]


class FsimRenderStaging(ctypes.Structure):
    _fields_ = [("waterCells", ctypes.c_void_p), ("waterCap", ctypes.c_size_t), ("solidCells", ctypes.c_void_p), ("solidCap", ctypes.c_size_t),
                ("cellVels", ctypes.c_void_p), ("pressureCells", ctypes.c_void_p), ("pressureValues", ctypes.c_void_p),
                ("pressureCap", ctypes.c_size_t), ("particleVelLines", ctypes.c_void_p), ("phiValues", ctypes.c_void_p),
                ("nWater", ctypes.c_size_t), ("nSolid", ctypes.c_size_t), ("nPressure", ctypes.c_size_t)]


class FsimError(RuntimeError):
    pass


def build(verbose=False):
    """Compile lib/libfsim_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-s", "-C", _HERE, "-j8"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """Load the C ABI. Raises if the CUDA library has not been built: there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FsimError("%s is missing: run `make -C fluid-sim_b200` (or __graft_entry__.build()); "
                        "this package has no CPU path" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cd, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_size_t
    L.fsim_default_options.argtypes = [ctypes.POINTER(FsimOptions)]
    L.fsim_default_options.restype = None
    L.fsim_create.argtypes = [ctypes.POINTER(FsimConfig), ctypes.POINTER(FsimOptions), ctypes.POINTER(vp)]
    L.fsim_destroy.argtypes = [vp]
    L.fsim_step.argtypes = [vp, ci]
    L.fsim_step_timed.argtypes = [vp, ci, ctypes.POINTER(cd)]
    L.fsim_diagnostics.argtypes = [vp, ctypes.POINTER(cd), ctypes.POINTER(cd), ctypes.POINTER(cd)]
    L.fsim_render_fill.argtypes = [vp, ctypes.POINTER(FsimRenderStaging)]
    L.fsim_checkpoint_save.argtypes = [vp, ctypes.c_char_p]
    L.fsim_checkpoint_load.argtypes = [ctypes.c_char_p, ctypes.POINTER(FsimOptions), ctypes.POINTER(vp)]
    L.fsim_stage.argtypes = [vp, ci]
    L.fsim_sync.argtypes = [vp]
    L.fsim_num_particles.argtypes = [vp, ctypes.POINTER(sz)]
    L.fsim_upload.argtypes = [vp, ci, vp, sz]
    L.fsim_download.argtypes = [vp, ci, vp, sz]
    L.fsim_set_particles.argtypes = [vp, sz, vp, vp]
    L.fsim_set_params.argtypes = [vp, cd, cd, cd, cd]
    L.fsim_set_pcg.argtypes = [vp, cd, ci]
    L.fsim_get_stats.argtypes = [vp, ctypes.POINTER(FsimStats)]
    L.fsim_step_host.argtypes = [vp, ctypes.POINTER(FsimHostMirror)]
    L.fsim_host_register.argtypes = [vp, vp, sz]
    L.fsim_host_unregister.argtypes = [vp, vp]
    L.fsim_launch_count.argtypes = [vp, ctypes.POINTER(ctypes.c_ulonglong)]
    L.fsim_profile_enable.argtypes = [vp, ci]
    L.fsim_profile_get.argtypes = [vp, ci, ctypes.POINTER(cd), ctypes.POINTER(ci)]
    L.fsim_profile_list.argtypes = [vp, ci, ctypes.POINTER(cd), ci, ctypes.POINTER(ci)]
    L.fsim_last_error.restype = ctypes.c_char_p
    L.fsim_version.restype = ctypes.c_char_p
    L.fsim_dist_unique_id.argtypes = [vp]
    L.fsim_dist_init.argtypes = [vp, ci, ci, vp]
    _lib = L
    return L


def dist_unique_id():
    """128-byte NCCL id (call on one rank, ship to the others)."""
    buf = (ctypes.c_ubyte * 128)()
    _check(lib().fsim_dist_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
    return bytes(buf)


def slab_rows(ny, world, rank):
    """Rows [j0, j1) of `rank` when the strips of 32 rows of a `ny`-row range are dealt to `world` ranks in
    contiguous blocks as even as possible (mirrors distSlabOf in csrc/dist.cu; the library applies it to the
    strips of the fluid cells' bounding box every step)."""
    ns = (ny + 31) // 32
    base, extra = divmod(ns, world)
    s0 = rank * base + min(rank, extra)
    n = base + (1 if rank < extra else 0)
    return min(ny, 32 * s0), min(ny, 32 * (s0 + n))


def _check(rc):
    if rc != 0:
        raise FsimError("fsim error %d: %s" % (rc, lib().fsim_last_error().decode()))


class FluidSim2D:
    """Mirror of the reference's struct FluidSim2D (include/FluidSim2D.h:65-176) over the C ABI."""

    def __init__(self, cells, dt, dx, rho=997.0, gravity=(0.0, -9.81), mode=FS_PICFLIP, picFlipAlpha=0.05,
                 particlesPerCellSqrt=2, pcgTol=1e-12, pcgMaxIters=200, device=0, seedParticles=True,
                 computeStats=True, slDoubleBuffer=False, debugSimpleWavefront=False, reserved=None):
        L = lib()
        cells = np.ascontiguousarray(cells, dtype=np.uint8)
        self.sizeY, self.sizeX = cells.shape
        self.dt, self.dx, self.mode = dt, dx, mode
        cfg = FsimConfig(self.sizeX, self.sizeY, particlesPerCellSqrt, dt, dx, rho, gravity[0], gravity[1], mode,
                         picFlipAlpha, cells.ctypes.data)
        opt = FsimOptions()
        L.fsim_default_options(ctypes.byref(opt))
        opt.pcgTol, opt.pcgMaxIters, opt.device = pcgTol, pcgMaxIters, device
        opt.seedParticles, opt.computeStats = int(seedParticles), int(computeStats)
        opt.slDoubleBuffer, opt.debugSimpleWavefront = int(slDoubleBuffer), int(debugSimpleWavefront)
        for k, v in enumerate(reserved or ()):  # schedule switches for A/B checks
            opt.reserved[k] = int(v)
        self._h = ctypes.c_void_p()
        _check(L.fsim_create(ctypes.byref(cfg), ctypes.byref(opt), ctypes.byref(self._h)))

    # -- checkpoints ----------------------------------------------------------------------------
    def save_checkpoint(self, path):
        _check(lib().fsim_checkpoint_save(self._h, os.fsencode(path)))

    @classmethod
    def load_checkpoint(cls, path, pcgTol=1e-12, pcgMaxIters=200, device=0, reserved=None):
        L = lib()
        opt = FsimOptions()
        L.fsim_default_options(ctypes.byref(opt))
        opt.pcgTol, opt.pcgMaxIters, opt.device = pcgTol, pcgMaxIters, device
        for k, v in enumerate(reserved or ()):
            opt.reserved[k] = int(v)
        self = cls.__new__(cls)
        self._h = ctypes.c_void_p()
        _check(L.fsim_checkpoint_load(os.fsencode(path), ctypes.byref(opt), ctypes.byref(self._h)))
        hdr = np.fromfile(path, dtype=np.int32, count=6)
        self.sizeX, self.sizeY, self.mode = int(hdr[2]), int(hdr[3]), int(hdr[5])
        d = np.fromfile(path, dtype=np.float64, count=5, offset=24)
        self.dt, self.dx = float(d[0]), float(d[1])
        return self

    # -- lifetime -------------------------------------------------------------------------------
    def free(self):
        if getattr(self, "_h", None):
            lib().fsim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # -- reference methods ----------------------------------------------------------------------
    def update(self, n=1):
        _check(lib().fsim_step(self._h, n))

    runFrame = update

    def update_timed(self, n=1):
        """n updates; returns their device time in ms (CUDA events on the library's stream)"""
        ms = ctypes.c_double()
        _check(lib().fsim_step_timed(self._h, n, ctypes.byref(ms)))
        return float(ms.value)

    def _diag(self):
        a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        _check(lib().fsim_diagnostics(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return float(a.value), float(b.value), float(c.value)

    def avgPressure(self): return self._diag()[0]
    def avgPressureInFluid(self): return self._diag()[1]
    def maxVelocity(self): return self._diag()[2]

    def stage(self, st):
        _check(lib().fsim_stage(self._h, st))

    def createWaterLevelSet(self): self.stage(CREATE_WATER_LEVEL_SET)
    def transferVelocityToGrid(self): self.stage(TRANSFER_VELOCITY_TO_GRID)
    def applySemiLagrangianAdvection(self): self.stage(APPLY_SEMI_LAGRANGIAN_ADVECTION)
    def applyGravity(self): self.stage(APPLY_GRAVITY)
    def createSolidLevelSet(self): self.stage(CREATE_SOLID_LEVEL_SET)
    def applyProjection(self): self.stage(APPLY_PROJECTION)
    def updateVelocity(self): self.stage(UPDATE_VELOCITY)
    def updateParticleVelocities(self): self.stage(UPDATE_PARTICLE_VELOCITIES)
    def applyAdvection(self): self.stage(APPLY_ADVECTION)

    def sync(self):
        _check(lib().fsim_sync(self._h))

    # -- data -----------------------------------------------------------------------------------
    @property
    def num_particles(self):
        n = ctypes.c_size_t()
        _check(lib().fsim_num_particles(self._h, ctypes.byref(n)))
        return int(n.value)

    def shape(self, field):
        nx, ny = self.sizeX, self.sizeY
        if field in (U, NEWU):
            return (ny, nx + 1), np.float64
        if field in (V, NEWV):
            return (ny + 1, nx), np.float64
        if field == CELL:
            return (ny, nx), np.uint8
        if field in (PARTICLES, PARTICLE_VELS):
            return (self.num_particles, 2), np.float64
        return (ny, nx), np.float64

    def get(self, field, out=None):
        shp, dt = self.shape(field)
        a = np.empty(shp, dtype=dt) if out is None else out
        assert a.shape == shp and a.dtype == dt and a.flags.c_contiguous
        _check(lib().fsim_download(self._h, field, a.ctypes.data, a.nbytes))
        return a

    def set(self, field, arr):
        shp, dt = self.shape(field)
        a = np.ascontiguousarray(arr, dtype=dt)
        assert a.shape == shp, (a.shape, shp)
        _check(lib().fsim_upload(self._h, field, a.ctypes.data, a.nbytes))

    def set_particles(self, pos, vel=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        vptr = None
        if vel is not None:
            vel = np.ascontiguousarray(vel, dtype=np.float64)
            assert vel.shape == pos.shape
            vptr = vel.ctypes.data
        _check(lib().fsim_set_particles(self._h, pos.shape[0], pos.ctypes.data, vptr))

    def set_params(self, gravity, picFlipAlpha, dt):
        self.dt = dt
        _check(lib().fsim_set_params(self._h, gravity[0], gravity[1], picFlipAlpha, dt))

    def set_pcg(self, tol, maxIters):
        _check(lib().fsim_set_pcg(self._h, tol, maxIters))

    def dist_init(self, rank, world, unique_id):
        """Join the y-slab pressure projection of `world` processes (one per GPU, include/fsim.h fsim_dist_init).
        `unique_id` is the 128-byte id rank 0 obtained from dist_unique_id()."""
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        _check(lib().fsim_dist_init(self._h, rank, world, ctypes.cast(buf, ctypes.c_void_p)))

    def stats(self):
        st = FsimStats()
        _check(lib().fsim_get_stats(self._h, ctypes.byref(st)))
        return st

    def step_host(self, mirror):
        _check(lib().fsim_step_host(self._h, ctypes.byref(mirror)))

    def host_register(self, arr):
        """page-locks a numpy array the caller keeps alive (see fsim_host_register)"""
        _check(lib().fsim_host_register(self._h, arr.ctypes.data, arr.nbytes))

    def host_unregister(self, arr):
        _check(lib().fsim_host_unregister(self._h, arr.ctypes.data))

    @property
    def launch_count(self):
        n = ctypes.c_ulonglong()
        _check(lib().fsim_launch_count(self._h, ctypes.byref(n)))
        return int(n.value)

    def profile_enable(self, on=True):
        _check(lib().fsim_profile_enable(self._h, int(on)))

    def profile_get(self, klass):
        ms, n = ctypes.c_double(), ctypes.c_int()
        _check(lib().fsim_profile_get(self._h, klass, ctypes.byref(ms), ctypes.byref(n)))
        return float(ms.value), int(n.value)

    def render_buffers(self):
        """FluidRenderer2D::updateBuffers (demo/FluidRenderer2D.cpp:435-486) from the device state: dict of numpy arrays"""
        nx, ny, n = self.sizeX, self.sizeY, self.num_particles
        out = {"water": np.zeros((nx * ny, 2), np.float32), "solid": np.zeros((nx * ny, 2), np.float32),
               "cellVels": np.zeros(((nx - 1) * (ny - 1) * 2, 2), np.float32), "pressureCells": np.zeros((nx * ny, 2), np.float32),
               "pressureValues": np.zeros(nx * ny, np.float32), "particleVelLines": np.zeros((2 * n, 2), np.float32),
               "phiValues": np.zeros(nx * ny, np.float32)}
        io = FsimRenderStaging()
        io.waterCells, io.waterCap = out["water"].ctypes.data, nx * ny
        io.solidCells, io.solidCap = out["solid"].ctypes.data, nx * ny
        io.cellVels = out["cellVels"].ctypes.data
        io.pressureCells, io.pressureValues, io.pressureCap = out["pressureCells"].ctypes.data, out["pressureValues"].ctypes.data, nx * ny
        io.particleVelLines = out["particleVelLines"].ctypes.data if n else None
        io.phiValues = out["phiValues"].ctypes.data
        _check(lib().fsim_render_fill(self._h, ctypes.byref(io)))
        out["water"] = out["water"][:io.nWater]; out["solid"] = out["solid"][:io.nSolid]
        out["pressureCells"] = out["pressureCells"][:io.nPressure]; out["pressureValues"] = out["pressureValues"][:io.nPressure]
        return out

    def profile_list(self, klass, cap=512):
        buf, n = (ctypes.c_double * cap)(), ctypes.c_int()
        _check(lib().fsim_profile_list(self._h, klass, buf, cap, ctypes.byref(n)))
        return [float(buf[k]) for k in range(min(cap, n.value))]

    def state(self, fields=(U, V, NEWU, NEWV, P, CELL, PHI, PARTICLES, PARTICLE_VELS)):
        return {f: self.get(f) for f in fields}
