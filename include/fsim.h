/* fsim.h -- C ABI of the B200-native fluid-sim hot path (libfsim_b200.so).
 *
 * The reference (lasagnaphil/fluid-sim @ 29962de) has no FFI layer: its boundary is the C++14 struct
 * FluidSim2D (reference include/FluidSim2D.h:65-176).  This header is what a binding for that struct
 * calls; fluid-sim_b200/shim/FluidSim2D_b200.cpp is the drop-in C++14 translation unit built on it (compiled against
 * the reference's own include/FluidSim2D.h, replacing src/FluidSim2D.cpp), and INTEGRATION.md shows the
 * maintainer-side wiring.  Plain pointers and sizes only; no CUDA or torch types.
 *
 * Conventions (all from the reference):
 *   - grids are dense row-major, a(i,j) = data[j*NX + i] (include/Array2D.h:43,87)
 *   - u is (sizeX+1) x sizeY, v is sizeX x (sizeY+1) (include/MACGrid2D.h:19-25)
 *   - cell labels: 0 EMPTY, 1 FLUID, 2 SOLID, one byte each (include/FluidSim2D.h:44-46)
 *   - particles / particle velocities are arrays of {x, y} doubles (deps/altmath/src/vec2.h:11-28)
 *   - the domain border must be SOLID
 * Every entry point returns 0 on success, a negative FSIM_E_* code otherwise; fsim_last_error() gives
 * the message for the calling thread.  A handle is not thread-safe; distinct handles are independent.
 * There is no CPU fallback: without a CUDA device fsim_create fails with FSIM_E_CUDA.
 */
#ifndef FSIM_H
#define FSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fsim_sim* fsim_handle;

enum { FSIM_OK = 0, FSIM_E_INVALID = -1, FSIM_E_CUDA = -2, FSIM_E_NOMEM = -3, FSIM_E_STATE = -4 };

/* FluidSimMode (include/FluidSim2D.h:40-42) */
enum { FSIM_SEMILAGRANGIAN = 0, FSIM_PICFLIP = 1 };

/* Field ids for upload/download.  0-8 are the public data members of FluidSim2D
 * (include/FluidSim2D.h:71-78); 9-13 are locals of applyProjection (src/FluidSim2D.cpp:253-258,334,366)
 * exposed read-only for stage-wise parity tests. */
enum {
    FSIM_U = 0, FSIM_V = 1, FSIM_NEWU = 2, FSIM_NEWV = 3, FSIM_P = 4, FSIM_CELL = 5, FSIM_PHI = 6,
    FSIM_PARTICLES = 7, FSIM_PARTICLE_VELS = 8,
    FSIM_ADIAG = 9, FSIM_AX = 10, FSIM_AY = 11, FSIM_RHS = 12, FSIM_PRECON = 13
};

/* Stage ids = FluidSim2D::StageType (include/FluidSim2D.h:93-97); each maps to the public method of
 * the same name (include/FluidSim2D.h:124-140). */
enum {
    FSIM_STAGE_CREATE_WATER_LEVEL_SET = 1,    /* src/FluidSim2D.cpp:653 */
    FSIM_STAGE_TRANSFER_VELOCITY_TO_GRID = 2, /* :144 */
    FSIM_STAGE_APPLY_SEMI_LAGRANGIAN_ADVECTION = 3, /* :206 */
    FSIM_STAGE_APPLY_GRAVITY = 4,             /* :237 */
    FSIM_STAGE_CREATE_SOLID_LEVEL_SET = 5,    /* :734 (empty in the reference) */
    FSIM_STAGE_APPLY_PROJECTION = 6,          /* :252 */
    FSIM_STAGE_UPDATE_VELOCITY = 7,           /* :469 */
    FSIM_STAGE_UPDATE_PARTICLE_VELOCITIES = 8,/* :552 */
    FSIM_STAGE_APPLY_ADVECTION = 9            /* :570 */
};

/* Mirrors FluidSim2DConfig (include/FluidSim2D.h:48-63). initialValues is borrowed for the call. */
typedef struct {
    int sizeX, sizeY, particlesPerCellSqrt;
    double dt, dx, rho, gravityX, gravityY;
    int mode;
    double picFlipAlpha;
    const uint8_t* initialValues; /* [sizeY*sizeX], index j*sizeX+i */
} fsim_config;

/* Extension parameters; zero-initialise and call fsim_default_options for the reference's constants. */
typedef struct {
    double pcgTol;      /* 1e-12  (src/FluidSim2D.cpp:453) */
    int pcgMaxIters;    /* 200    (src/FluidSim2D.cpp:429) */
    int device;         /* CUDA device ordinal, default 0 */
    int seedParticles;  /* 1: seed like FluidSim2D::create (glibc rand(), src/FluidSim2D.cpp:52-64); 0: none */
    int computeStats;   /* 1: volume/energy sums every step like src/FluidSim2D.cpp:709-731 */
    int slDoubleBuffer; /* 0: exact in-place raster order of :206-235 (default); 1: snapshot variant */
    int debugSimpleWavefront; /* 1: run every wavefront stage with the slow single-CTA scheduler (debug) */
    int reserved[8];    /* 0 = default; schedule switches for A/B checks (same arithmetic), indexed by FSIM_OPT_* */
} fsim_options;

/* indices into fsim_options.reserved: every switch changes the schedule, never the arithmetic */
enum {
    FSIM_OPT_SD_SIGMA = 0,          /* lane skew (2 or 3) of the one-row-per-lane solve kernel (FSIM_SD_RPL=1 builds only) */
    FSIM_OPT_NO_FLUID_BOX = 1,      /* 1: the projection covers the whole grid, not the FLUID cells' bounding box */
    FSIM_OPT_NO_STRIP_RANGES = 2,   /* 1: the triangular solves march every chunk of every strip */
    FSIM_OPT_NO_SECOND_STREAM = 3,  /* 1: transferVelocityToGrid runs after, not beside, createWaterLevelSet */
    FSIM_OPT_SERIAL_MIRRORS = 4,    /* 1: fsim_step_host copies the mirrors in order on the one stream */
    FSIM_OPT_SL_SELF_VALIDATING = 5,/* 1: in-place semi-Lagrangian kernel that polls the NEW values themselves; 2: progress
                                       counters through global memory even where the grid fits the shared-memory kernel */
    FSIM_OPT_LATE_EXTRAP_PREP = 6,  /* 1: updateVelocity's extrapolation structure is built inside stage 7 */
    FSIM_OPT_FUSED_AXPY = 7         /* 1: p += alpha s, r -= alpha z inside the triangular solves instead of a kernel of their own
                                       (same bits; measured slower on B200) */
};

/* Per-step diagnostics (FluidSim2D::waterVolume/totalEnergy/particleTotalEnergy, include/FluidSim2D.h:106-114;
 * PCG loop state, src/FluidSim2D.cpp:429-466; CFL diagnostic :572-585; NaN check :598-601). */
typedef struct {
    double waterVolume, totalEnergy, particleTotalEnergy, currentTime;
    int pcgIters;          /* value of `iter` when the PCG loop ended */
    int pcgHitMaxIters;    /* the reference logs "Maximum iteration limit exceeded!" */
    double pcgResidual;    /* last |r|_inf */
    double pcgRhsNorm;     /* |rhs|_inf */
    double cflMax;         /* max (vx+vy)*dt/dx over particles */
    int nanPositions;      /* particles whose position was NaN before advection */
    int levelSetSweeps;    /* directional sweeps actually executed (<= 32, early exit at a fixed point) */
    int extrapolationLayers; /* BFS layers of the last extrapolation */
    float stageMs[8];      /* CUDA-event time of each stage of the last fsim_step (PerformanceCounter analogue) */
    int numStages;
    long long pcgSolveCells; /* cells the last projection's PCG covered: the bounding box of the FLUID cells (whole strips
                                of 32 rows), or the rank's slab in the multi-GPU mode */
    long long pcgMarchedCells; /* cells (layout slots) the two triangular solves actually march: per strip only the
                                  32-step chunks that hold fluid */
    int distError;         /* multi-GPU: 1 if a wait on a peer rank timed out during the last projection (the solve stopped) */
    int extrapolationNearLayers; /* update() only: BFS layers of the extrapolation after updateVelocity that were filled before the
                                  particle stages (the rest ran beside them); 0 if the fill was not split in the last frame */
} fsim_stats;

void fsim_default_options(fsim_options* opt);
int fsim_create(const fsim_config* cfg, const fsim_options* opt /* may be NULL */, fsim_handle* out);
int fsim_destroy(fsim_handle h);

/* FluidSim2D::update() n times (src/FluidSim2D.cpp:140-142); asynchronous w.r.t. the host until a
 * download, fsim_get_stats or fsim_sync. */
int fsim_step(fsim_handle h, int nsteps);
/* The same, bracketed by CUDA events on the simulation's own stream; returns when the steps are done.
 * deviceMs = device time of the nsteps updates (bench.py's timed region). */
int fsim_step_timed(fsim_handle h, int nsteps, double* deviceMs);
/* One public stage method (for stage-wise parity). */
int fsim_stage(fsim_handle h, int stage);
int fsim_sync(fsim_handle h);

/* Multi-GPU (one process per GPU on one NVLink/NVSwitch node): y-slab partition of the pressure projection (the PCG of
 * src/FluidSim2D.cpp:423-466).  Per iteration the one-row halo of the search direction and the PCG scalars travel
 * through peer memory (CUDA IPC mappings set up here; the kernels store into their peers' HBM directly, no collective
 * call inside the iteration); across slab boundaries the preconditioner is block-MIC(0).  Every rank holds the full
 * replicated state and runs the other stages redundantly, so all ranks stay bit-identical.
 * fsim_dist_unique_id fills 128 bytes on one rank (ncclGetUniqueId; NCCL carries the handle exchange at init and the
 * once-per-step exchange of pressure rows); the caller ships them to the other ranks.  At most 16 ranks. */
int fsim_dist_unique_id(void* out128);
int fsim_dist_init(fsim_handle h, int rank, int world, const void* uniqueId128);

int fsim_num_particles(fsim_handle h, size_t* n);
/* bytes must equal the dense size of the field; host memory may be pageable or pinned. */
int fsim_upload(fsim_handle h, int field, const void* src, size_t bytes);
int fsim_download(fsim_handle h, int field, void* dst, size_t bytes);
/* Replace the particle set (n entries of {x,y} each). */
int fsim_set_particles(fsim_handle h, size_t n, const double* pos, const double* vel);
/* gravity, picFlipAlpha and dt are re-read by the reference every update(); callers write them as fields. */
int fsim_set_params(fsim_handle h, double gravityX, double gravityY, double picFlipAlpha, double dt);
int fsim_set_pcg(fsim_handle h, double tol, int maxIters);
int fsim_get_stats(fsim_handle h, fsim_stats* out);
/* FluidSim2D::avgPressure / avgPressureInFluid / maxVelocity (src/FluidSim2D.cpp:607-638) as one fused reduction over the
 * device-resident p, labels and mac; any pointer may be NULL.  avgPressureInFluid is NaN without FLUID cells, as in the
 * reference (0/0). */
int fsim_diagnostics(fsim_handle h, double* avgPressure, double* avgPressureInFluid, double* maxVelocity);

/* The staging arrays FluidRenderer2D::updateBuffers rebuilds from the state every frame (demo/FluidRenderer2D.cpp:435-486),
 * filled from the device-resident state: a caller that draws the simulation then needs no per-frame download of the whole
 * state.  Every output pointer may be NULL (skipped); vec2f lists are {x, y} floats; the variable-length lists keep the
 * reference's raster order (FluidSim2D::iterate, j outer, i inner) and report their lengths; *Cap are capacities in entries. */
typedef struct {
    float* waterCells;      size_t waterCap;     /* :436-448  {(float)(i*dx), (float)(j*dx)} of every FLUID cell */
    float* solidCells;      size_t solidCap;     /*           ... of every SOLID cell */
    float* cellVels;                             /* :449-459  2 vec2f per cell of the (sizeX-1) x (sizeY-1) block: centre, centre + dt*velInterp */
    float* pressureCells;   float* pressureValues; size_t pressureCap; /* :460-470  cells with p != 0: location, sigmoid(0.01f * p) */
    float* particleVelLines;                     /* :471-479  2 vec2f per particle: position, position + dt*velInterp(position) */
    float* phiValues;                            /* :480-485  sigmoid(100 * phi) per cell, [sizeY*sizeX] */
    size_t nWater, nSolid, nPressure;            /* out: lengths of the three lists */
} fsim_render_staging;
int fsim_render_fill(fsim_handle h, fsim_render_staging* io);

/* State checkpoint (everything update() carries over: mac, newMac, p, cell, phi, particles, particleVels and dt, gravity,
 * picFlipAlpha, currentTime) as one binary file, format in csrc/checkpoint.cu.  The reference has no checkpointing
 * (SURVEY.md section 5); fsim_checkpoint_load creates a new handle (opt may be NULL) that continues bit for bit. */
int fsim_checkpoint_save(fsim_handle h, const char* path);
int fsim_checkpoint_load(const char* path, const fsim_options* opt, fsim_handle* out);

/* End-to-end step with HOST buffers, as a caller that owns host mirrors uses it (the reference's renderer
 * writes mac.u/v and reads every public field each frame, demo/FluidRenderer2D.cpp:305-308, 436-485):
 * uploads u and v (if non-NULL), runs one update(), downloads the fields whose pointers are non-NULL. */
typedef struct {
    const double* u_in; const double* v_in;
    double* u; double* v; double* p; uint8_t* cell; double* phi; double* particles; double* particleVels;
} fsim_host_mirror;
int fsim_step_host(fsim_handle h, const fsim_host_mirror* io);

/* Page-locks host memory the caller owns (the reference allocates its public arrays with malloc, Array2D.h:63-72), so
 * that fsim_step_host copies it at full PCIe rate and beside the stages: with every mirror buffer page-locked the
 * upload overlaps the level set and phi, labels, p and the particle velocities leave as soon as no later stage writes
 * them; with pageable buffers the copies run in order on the one stream.  Unregister before freeing the memory. */
int fsim_host_register(fsim_handle h, void* ptr, size_t bytes);
int fsim_host_unregister(fsim_handle h, void* ptr);

/* Per-kernel device timing of the PCG inner loop with CUDA events on the launching stream (bench.py's
 * roofline).  Classes: 0 applyA+dot, 1 axpy+norm, 2 forward solve, 3 backward solve+dot, 4 s-update.
 * fsim_profile_get synchronises, returns the summed duration and launch count since the last enable. */
int fsim_profile_enable(fsim_handle h, int on);
int fsim_profile_get(fsim_handle h, int klass, double* totalMs, int* launches);
/* ... and the individual durations of that class in launch order (at most cap are stored; *count = how many there were).
 * Further classes: 5 closest-particle sweep, 6 eikonal sweep, 7 extrapolation layer fill, 8 MIC(0) factor, 9 distance transform. */
int fsim_profile_list(fsim_handle h, int klass, double* ms, int cap, int* count);

/* Number of CUDA kernels launched by this handle so far (bench.py's gpu_launches). */
int fsim_launch_count(fsim_handle h, unsigned long long* n);

const char* fsim_last_error(void);
const char* fsim_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FSIM_H */
