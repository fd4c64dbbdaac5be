// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI harness around the UNMODIFIED reference solver (lasagnaphil/fluid-sim @ 29962de).
// It is compiled by oracle/Makefile together with the reference's own sources *where they
// lie* under /root/reference (nothing is copied into this repository); the output goes to
// oracle/_ref/ (git-ignored).  The harness only drives the reference through its public
// surface, struct FluidSim2D (include/FluidSim2D.h:65-176): create/free/update, the nine
// public stage methods, and the public data members.
//
// The same `fso_*` entry points are exported by the C restatement (oracle/fsim_oracle.c)
// so tests can run either implementation behind one ctypes wrapper (tests/oracle_lib.py).
//
// Two builds exist (see oracle/Makefile):
//   libfsim_ref.so          stock sources, nothing touched.
//   libfsim_ref_patched.so  FluidSim2D.cpp streamed through `sed` into the compiler with
//                           (1) PCG tolerance / iteration cap lifted into globals whose
//                           defaults are the reference's literals (FluidSim2D.cpp:429,453),
//                           (2) the final PCG iteration count exported, and (3) an optional
//                           double-buffered semi-Lagrangian advection.
//                           With the defaults it is arithmetic-identical to the stock build
//                           (tests/test_oracle.py checks this bit for bit).

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include <log.h>
#include "FluidSim2D.h"

#ifdef FSIM_REF_PATCHED
// defined here, referenced by the sed-patched FluidSim2D.cpp
double g_fsim_ref_tol = 1e-12;
int g_fsim_ref_max_iters = 200;
int g_fsim_ref_last_iters = -1;
int g_fsim_ref_sl_db = 0;
// locals of applyProjection (src/FluidSim2D.cpp:253-258, 334, 366), copied out by a call the sed patch inserts
// right before the PCG loop (:424); fields 9..13 of fso_get
#include <vector>
static std::vector<double> g_proj[5];
void fsim_ref_export_projection(const double* Adiag, const double* Ax, const double* Ay, const double* rhs,
                                const double* precon, size_t n) {
    const double* src[5] = {Adiag, Ax, Ay, rhs, precon};
    for (int k = 0; k < 5; k++) g_proj[k].assign(src[k], src[k] + n);
}
#endif

namespace {

enum Field {
    F_U = 0, F_V = 1, F_NEWU = 2, F_NEWV = 3, F_P = 4, F_CELL = 5, F_PHI = 6,
    F_PARTICLES = 7, F_PARTICLE_VELS = 8
};

struct Harness {
    FluidSim2D sim;
};

size_t fieldBytes(FluidSim2D& s, int field, void** ptr) {
    switch (field) {
        case F_U: *ptr = s.mac.u.data; return sizeof(double) * (size_t)s.mac.u.NX * s.mac.u.NY;
        case F_V: *ptr = s.mac.v.data; return sizeof(double) * (size_t)s.mac.v.NX * s.mac.v.NY;
        case F_NEWU: *ptr = s.newMac.u.data; return sizeof(double) * (size_t)s.newMac.u.NX * s.newMac.u.NY;
        case F_NEWV: *ptr = s.newMac.v.data; return sizeof(double) * (size_t)s.newMac.v.NX * s.newMac.v.NY;
        case F_P: *ptr = s.p.data; return sizeof(double) * (size_t)s.p.NX * s.p.NY;
        case F_CELL: *ptr = s.cell.data; return (size_t)s.cell.NX * s.cell.NY;
        case F_PHI: *ptr = s.waterLevelSet.phi.data; return sizeof(double) * (size_t)s.sizeX * s.sizeY;
        case F_PARTICLES: *ptr = s.particles.data; return sizeof(vec2d) * s.particles.size;
        case F_PARTICLE_VELS: *ptr = s.particleVels.data; return sizeof(vec2d) * s.particleVels.size;
        default: *ptr = nullptr; return 0;
    }
}

}  // namespace

extern "C" {

const char* fso_kind() {
#ifdef FSIM_REF_PATCHED
    return "reference-patched";
#else
    return "reference-stock";
#endif
}

void* fso_create(int sizeX, int sizeY, int ppcSqrt, double dt, double dx, double rho,
                 double gx, double gy, int mode, double alpha, const uint8_t* cells) {
    log_set_quiet(1);
    srand(1);  // glibc's implicit seed, made explicit so every create() sees the same jitter
    FluidSim2DConfig cfg = {};
    cfg.sizeX = sizeX;
    cfg.sizeY = sizeY;
    cfg.particlesPerCellSqrt = ppcSqrt;
    cfg.dt = dt;
    cfg.dx = dx;
    cfg.rho = rho;
    cfg.gravityX = gx;
    cfg.gravityY = gy;
    cfg.mode = mode == 0 ? FS_SEMILAGRANGIAN : FS_PICFLIP;
    cfg.picFlipAlpha = alpha;
    cfg.initialValues = (FluidCellType*)cells;
    Harness* h = new Harness;
    h->sim = FluidSim2D::create(cfg);
    return h;
}

void fso_destroy(void* hv) {
    Harness* h = (Harness*)hv;
    h->sim.free();
    delete h;
}

long fso_num_particles(void* hv) { return (long)((Harness*)hv)->sim.particles.size; }

int fso_get(void* hv, int field, void* dst) {
#ifdef FSIM_REF_PATCHED
    if (field >= 9 && field <= 13) {  // projection internals of the last applyProjection
        FluidSim2D& s = ((Harness*)hv)->sim;
        const std::vector<double>& v = g_proj[field - 9];
        if (v.size() != (size_t)s.sizeX * s.sizeY) return -1;
        memcpy(dst, v.data(), v.size() * sizeof(double));
        return 0;
    }
#endif
    void* p;
    size_t n = fieldBytes(((Harness*)hv)->sim, field, &p);
    if (!p) return -1;
    memcpy(dst, p, n);
    return 0;
}

int fso_set(void* hv, int field, const void* src) {
    void* p;
    size_t n = fieldBytes(((Harness*)hv)->sim, field, &p);
    if (!p) return -1;
    memcpy(p, src, n);
    return 0;
}

// Replace the particle set (positions and velocities, n entries each of 2 doubles).
int fso_set_particles(void* hv, long n, const double* pos, const double* vel) {
    FluidSim2D& s = ((Harness*)hv)->sim;
    s.particles.free();
    s.particleVels.free();
    s.particles = Vec<vec2d>::emptyWithSize((size_t)n);
    s.particleVels = Vec<vec2d>::emptyWithSize((size_t)n);
    memcpy(s.particles.data, pos, sizeof(vec2d) * (size_t)n);
    memcpy(s.particleVels.data, vel, sizeof(vec2d) * (size_t)n);
    return 0;
}

// stage ids follow FluidSim2D::StageType (include/FluidSim2D.h:93-97)
int fso_stage(void* hv, int stage) {
    FluidSim2D& s = ((Harness*)hv)->sim;
    switch (stage) {
        case 1: s.createWaterLevelSet(); break;
        case 2: s.transferVelocityToGrid(); break;
        case 3: s.applySemiLagrangianAdvection(); break;
        case 4: s.applyGravity(); break;
        case 5: s.createSolidLevelSet(); break;
        case 6: s.applyProjection(); break;
        case 7: s.updateVelocity(); break;
        case 8: s.updateParticleVelocities(); break;
        case 9: s.applyAdvection(); break;
        default: return -1;
    }
    return 0;
}

int fso_step(void* hv, int n) {
    FluidSim2D& s = ((Harness*)hv)->sim;
    for (int k = 0; k < n; k++) s.update();
    return 0;
}

void fso_set_params(void* hv, double gx, double gy, double alpha, double dt) {
    FluidSim2D& s = ((Harness*)hv)->sim;
    s.gravity = vec2d{gx, gy};
    s.picFlipAlpha = alpha;
    s.dt = dt;
}

// which: 0 waterVolume, 1 totalEnergy, 2 particleTotalEnergy, 3 currentTime,
//        4 avgPressure(), 5 avgPressureInFluid(), 6 maxVelocity()  (the reference's own methods, src/FluidSim2D.cpp:607-638)
double fso_stat(void* hv, int which) {
    FluidSim2D& s = ((Harness*)hv)->sim;
    switch (which) {
        case 0: return s.waterVolume;
        case 1: return s.totalEnergy;
        case 2: return s.particleTotalEnergy;
        case 3: return s.currentTime;
        case 4: return s.avgPressure();
        case 5: return s.avgPressureInFluid();
        case 6: return s.maxVelocity();
        default: return 0.0;
    }
}

// per-stage wall-clock (ms) of the most recent frame, from the reference's own
// PerformanceCounter ring (src/PerformanceCounter.cpp:16-50)
int fso_stage_times(void* hv, float* out, int maxStages) {
    FluidSim2D& s = ((Harness*)hv)->sim;
    int frame = (s.perfCounter.currentFrame + PerformanceCounter::SampleCount - 1) % PerformanceCounter::SampleCount;
    int n = (int)s.perfCounter.samples.size;
    if (n > maxStages) n = maxStages;
    for (int i = 0; i < n; i++) out[i] = s.perfCounter.samples[i][frame];
    return n;
}

// mac.velInterp (include/MACGrid2D.h:96-98) at n positions {x, y}; out receives {vx, vy} -- the reference's own sampler, for
// checks of anything that interpolates the grid velocity (renderer staging arrays, diagnostics)
int fso_vel_interp(void* hv, long n, const double* pos, double* out) {
    FluidSim2D& s = ((Harness*)hv)->sim;
    for (long k = 0; k < n; k++) {
        vec2d v = s.mac.velInterp(vec2d{pos[2 * k], pos[2 * k + 1]});
        out[2 * k] = v.x;
        out[2 * k + 1] = v.y;
    }
    return 0;
}

int fso_set_pcg(double tol, int maxIters) {
#ifdef FSIM_REF_PATCHED
    g_fsim_ref_tol = tol;
    g_fsim_ref_max_iters = maxIters;
    return 0;
#else
    (void)tol; (void)maxIters;
    return -1;  // stock build: literals are compiled in
#endif
}

int fso_last_pcg_iters(void* hv) {
    (void)hv;
#ifdef FSIM_REF_PATCHED
    return g_fsim_ref_last_iters;
#else
    return -1;
#endif
}

int fso_set_sl_double_buffer(int enable) {
#ifdef FSIM_REF_PATCHED
    g_fsim_ref_sl_db = enable;
    return 0;
#else
    (void)enable;
    return -1;
#endif
}

}  // extern "C"
