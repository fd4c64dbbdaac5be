// C ABI (include/fsim.h): handle lifetime, dense <-> frame transfers, stage dispatch, per-stage timing.
// Mirrors FluidSim2D::create/free/update/runFrame (reference src/FluidSim2D.cpp:23-142).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <utility>

#include "sampling.cuh"
#include "sim.h"
#include "wavefront.cuh"

int pcgSetParams(Sim* s, double tol, int maxIters);

static thread_local char g_err[512] = "";

void fsim_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* fsim_last_error(void) { return g_err; }
extern "C" const char* fsim_version(void) { return "fsim_b200 0.1 (sm_100a)"; }

extern "C" void fsim_default_options(fsim_options* o) {
    memset(o, 0, sizeof(*o));
    o->pcgTol = 1e-12;
    o->pcgMaxIters = 200;
    o->device = 0;
    o->seedParticles = 1;
    o->computeStats = 1;
    o->slDoubleBuffer = 0;
    o->debugSimpleWavefront = 0;
}

namespace {

__global__ void fillU64Kernel(unsigned long long* p, size_t n, unsigned long long v) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) p[k] = v;
}

// avgPressure / avgPressureInFluid / maxVelocity (reference src/FluidSim2D.cpp:607-638) in one pass: per block the sum of p
// over all cells, the sum and count over FLUID cells and the largest |velInterp((i, j) dx)| over FLUID cells; the last
// block adds the partials up in block order (deterministic).  out[0..3] = sum p, sum p in fluid, fluid count, max |vel|.
__global__ void __launch_bounds__(256) diagnosticsKernel(const double* __restrict__ p, const uint8_t* __restrict__ cell, GridView g,
                                                         double* partials, unsigned int* counter, double* out) {
    __shared__ double red[32];
    __shared__ bool isLast;
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), j = blockIdx.y * 8 + (threadIdx.x >> 5);
    double sp = 0.0, spf = 0.0, cnt = 0.0, vmax = 0.0;
    if (i < g.nx && j < g.ny) {
        const long long o = (long long)j * g.pitch + i;
        sp = p[o];
        if (cell[o] == FSIM_CELL_FLUID) {
            spf = sp; cnt = 1.0;
            const double x = (double)i * g.dx, y = (double)j * g.dx;
            const double vx = sampleU<false>(g, x, y), vy = sampleV<false>(g, x, y);
            vmax = sqrt(vx * vx + vy * vy);
        }
    }
    const unsigned int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
    double v[4] = {sp, spf, cnt, vmax};
    for (int q = 0; q < 4; ++q) {
        const double r = q < 3 ? blockReduce<false>(v[q], red) : blockReduce<true>(v[q], red);
        if (threadIdx.x == 0) partials[(size_t)q * nblocks + bid] = r;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        isLast = atomicAdd(counter, 1u) == nblocks - 1;
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    for (int q = 0; q < 4; ++q) {
        double acc = 0.0;
        // block order within a thread, thread order in the tree: the same grouping every run
        for (unsigned int k = threadIdx.x; k < nblocks; k += blockDim.x) {
            const double pv = __ldcg(&partials[(size_t)q * nblocks + k]);
            acc = q < 3 ? acc + pv : fmax(acc, pv);
        }
        acc = q < 3 ? blockReduce<false>(acc, red) : blockReduce<true>(acc, red);
        if (threadIdx.x == 0) out[q] = acc;
    }
    if (threadIdx.x == 0) *counter = 0;
}

template <class T>
int allocFrame(Sim* s, T** out) {
    void* raw = nullptr;
    size_t bytes = s->fr.elems * sizeof(T);
    CUDA_TRY(cudaMalloc(&raw, bytes));
    s->rawAllocs.push_back(raw);
    CUDA_TRY(cudaMemsetAsync(raw, 0, bytes, s->stream));
    *out = reinterpret_cast<T*>(raw) + s->fr.org;
    return FSIM_OK;
}

template <class T>
int allocLinear(Sim* s, T** out, size_t n) {
    void* raw = nullptr;
    CUDA_TRY(cudaMalloc(&raw, (n ? n : 1) * sizeof(T)));
    s->rawAllocs.push_back(raw);
    CUDA_TRY(cudaMemsetAsync(raw, 0, (n ? n : 1) * sizeof(T), s->stream));
    *out = reinterpret_cast<T*>(raw);
    return FSIM_OK;
}

int allocParticles(Sim* s, size_t n) {
    if (n <= s->npCap) return FSIM_OK;
    size_t cap = n + n / 8 + 1024;
    // particle arrays are the only ones that can grow (fsim_set_particles); free the old ones
    void* olds[] = {s->pos, s->vel, s->pcell, s->sortedIdx};
    for (void* p : olds) {
        if (!p) continue;
        for (auto& r : s->rawAllocs) if (r == p) r = nullptr;
        cudaFree(p);
    }
    int rc;
    if ((rc = allocLinear(s, &s->pos, cap))) return rc;
    if ((rc = allocLinear(s, &s->vel, cap))) return rc;
    if ((rc = allocLinear(s, &s->pcell, cap))) return rc;
    if ((rc = allocLinear(s, &s->sortedIdx, cap))) return rc;
    s->npCap = cap;
    return FSIM_OK;
}

struct FieldInfo {
    void* base;   // device pointer at (0,0), or linear
    int NX, NY;   // dense extents (0 for linear)
    size_t elem;  // bytes per element
    size_t linearCount;
};

int fieldInfo(Sim* s, int field, FieldInfo* fi) {
    int nx = s->nx, ny = s->ny;
    fi->linearCount = 0;
    switch (field) {
        case FSIM_U: *fi = {s->u, nx + 1, ny, 8, 0}; return 0;
        case FSIM_V: *fi = {s->v, nx, ny + 1, 8, 0}; return 0;
        case FSIM_NEWU: *fi = {s->nu, nx + 1, ny, 8, 0}; return 0;
        case FSIM_NEWV: *fi = {s->nv, nx, ny + 1, 8, 0}; return 0;
        case FSIM_P: *fi = {s->p, nx, ny, 8, 0}; return 0;
        case FSIM_CELL: *fi = {s->cell, nx, ny, 1, 0}; return 0;
        case FSIM_PHI: *fi = {s->phi, nx, ny, 8, 0}; return 0;
        case FSIM_ADIAG: *fi = {s->Adiag, nx, ny, 8, 0}; return 0;
        case FSIM_AX: *fi = {s->Ax, nx, ny, 8, 0}; return 0;
        case FSIM_AY: *fi = {s->Ay, nx, ny, 8, 0}; return 0;
        case FSIM_RHS: *fi = {s->rhs, nx, ny, 8, 0}; return 0;
        case FSIM_PRECON: *fi = {s->pc, nx, ny, 8, 0}; return 0;
        case FSIM_PARTICLES: *fi = {s->pos, 0, 0, 16, s->np}; return 0;
        case FSIM_PARTICLE_VELS: *fi = {s->vel, 0, 0, 16, s->np}; return 0;
        default: return -1;
    }
}

// glibc-compatible particle seeding of FluidSim2D::create (src/FluidSim2D.cpp:19-21, 49-69): raster order over
// FLUID cells, k = 0..ppc-1, jitter from rand() with the implicit seed 1, `dist` held in float.
void seedParticles(const fsim_config* c, std::vector<double>& pos) {
    int s = c->particlesPerCellSqrt, ppc = s * s;
    float dist = 1.0f / s;
    srand(1);
    for (int j = 0; j < c->sizeY; j++)
        for (int i = 0; i < c->sizeX; i++)
            if (c->initialValues[(size_t)j * c->sizeX + i] == FSIM_CELL_FLUID)
                for (int k = 0; k < ppc; k++) {
                    int x = k % s, y = k / s;
                    double rx = ((double)rand() / (double)RAND_MAX) * dist;
                    double px = ((double)i + dist * x + rx) * c->dx;
                    double ry = ((double)rand() / (double)RAND_MAX) * dist;
                    double py = ((double)j + dist * y + ry) * c->dx;
                    pos.push_back(px);
                    pos.push_back(py);
                }
}

int runStage(Sim* s, int stage) {
    switch (stage) {
        case FSIM_STAGE_CREATE_WATER_LEVEL_SET: return stageCreateWaterLevelSet(s);
        case FSIM_STAGE_TRANSFER_VELOCITY_TO_GRID: return stageTransferVelocityToGrid(s);
        case FSIM_STAGE_APPLY_SEMI_LAGRANGIAN_ADVECTION: return stageApplySemiLagrangianAdvection(s);
        case FSIM_STAGE_APPLY_GRAVITY: return stageApplyGravity(s);
        case FSIM_STAGE_CREATE_SOLID_LEVEL_SET: return FSIM_OK;  // empty in the reference (:734-736)
        case FSIM_STAGE_APPLY_PROJECTION: return stageApplyProjection(s);
        case FSIM_STAGE_UPDATE_VELOCITY: return stageUpdateVelocity(s);
        case FSIM_STAGE_UPDATE_PARTICLE_VELOCITIES: return stageUpdateParticleVelocities(s);
        case FSIM_STAGE_APPLY_ADVECTION: return stageApplyAdvection(s);
        default: fsim_set_error("unknown stage %d", stage); return FSIM_E_INVALID;
    }
}

// ---- host mirrors of fsim_step_host ---------------------------------------------------------------------------------
enum { M_U = 1, M_V = 2, M_P = 4, M_CELL = 8, M_PHI = 16, M_POS = 32, M_VEL = 64, M_ALL = 127 };

// Queues the downloads in `which` that have not been issued yet, behind everything enqueued on s->stream so far: on the
// copy stream when the mirrors are pinned (they then run beside the following stages), else on s->stream itself.
int mirrorDownload(Sim* s, unsigned which) {
    const fsim_host_mirror* io = s->mirror;
    if (!io) return FSIM_OK;
    which &= ~s->mirrorDone;
    if (!which) return FSIM_OK;
    cudaStream_t cs = s->stream;
    if (s->mirrorOverlap) {
        CUDA_TRY(cudaEventRecord(s->evMirror, s->stream));
        CUDA_TRY(cudaStreamWaitEvent(s->copyStream, s->evMirror, 0));
        cs = s->copyStream;
    }
    const Frame& f = s->fr;
    const int nx = s->nx, ny = s->ny;
    if ((which & M_U) && io->u) CUDA_TRY(cudaMemcpy2DAsync(io->u, (nx + 1) * 8, s->u, f.pitch * 8, (nx + 1) * 8, ny, cudaMemcpyDeviceToHost, cs));
    if ((which & M_V) && io->v) CUDA_TRY(cudaMemcpy2DAsync(io->v, nx * 8, s->v, f.pitch * 8, nx * 8, ny + 1, cudaMemcpyDeviceToHost, cs));
    if ((which & M_P) && io->p) CUDA_TRY(cudaMemcpy2DAsync(io->p, nx * 8, s->p, f.pitch * 8, nx * 8, ny, cudaMemcpyDeviceToHost, cs));
    if ((which & M_PHI) && io->phi) CUDA_TRY(cudaMemcpy2DAsync(io->phi, nx * 8, s->phi, f.pitch * 8, nx * 8, ny, cudaMemcpyDeviceToHost, cs));
    if ((which & M_CELL) && io->cell) CUDA_TRY(cudaMemcpy2DAsync(io->cell, nx, s->cell, f.pitch, nx, ny, cudaMemcpyDeviceToHost, cs));
    if ((which & M_POS) && io->particles && s->np) CUDA_TRY(cudaMemcpyAsync(io->particles, s->pos, s->np * 16, cudaMemcpyDeviceToHost, cs));
    if ((which & M_VEL) && io->particleVels && s->np) CUDA_TRY(cudaMemcpyAsync(io->particleVels, s->vel, s->np * 16, cudaMemcpyDeviceToHost, cs));
    s->mirrorDone |= which;
    return FSIM_OK;
}

// Fields no later stage of the frame writes: phi and the labels after the level set, p after the projection, the
// particle velocities after the grid-to-particle transfer, the grid velocities after the last copy into mac
// (src/FluidSim2D.cpp:547-549 in semi-Lagrangian mode, :566 in PIC/FLIP mode).
int mirrorAfterStage(Sim* s, int stage) {
    if (!s->mirror || !s->mirrorOverlap) return FSIM_OK;
    switch (stage) {
        case FSIM_STAGE_CREATE_WATER_LEVEL_SET: return mirrorDownload(s, M_PHI | M_CELL);
        case FSIM_STAGE_APPLY_PROJECTION: return mirrorDownload(s, M_P);
        // (split extrapolation: the far layers of u, v are still being filled on the second stream -- after the join)
        case FSIM_STAGE_UPDATE_VELOCITY: return s->mode == FSIM_SEMILAGRANGIAN && !s->farPending ? mirrorDownload(s, M_U | M_V) : FSIM_OK;
        case FSIM_STAGE_UPDATE_PARTICLE_VELOCITIES: return mirrorDownload(s, s->farPending ? M_VEL : (M_U | M_V | M_VEL));
        default: return FSIM_OK;
    }
}

// runFrame (src/FluidSim2D.cpp:94-138)
int runFrame(Sim* s) {
    static const int sl[] = {1, 3, 4, 5, 6, 7, 9};
    static const int pf[] = {1, 2, 4, 5, 6, 7, 8, 9};
    const int* order = s->mode == FSIM_SEMILAGRANGIAN ? sl : pf;
    int n = s->mode == FSIM_SEMILAGRANGIAN ? 7 : 8;
    CUDA_TRY(cudaEventRecord(s->stageEv[0], s->stream));
    s->extrapReady = false; s->prepPending = false;
    static const bool noSplit = getenv("FSIM_NO_SPLIT_FILL") && atoi(getenv("FSIM_NO_SPLIT_FILL")) != 0;
    s->splitFill = !noSplit && s->stream2 != nullptr && s->opt.reserved[3] != 1;
    s->lastSplit = false;
    // Every way out of the frame leaves the switch off (stage-wise callers never split) and the far fill joined -- except
    // between the frames of one fsim_step / fsim_step_timed call (deferFarJoin): there the next frame's level set, which
    // does not touch the grid velocities until its statistics kernel, runs beside the rest of the fill; that stage joins
    // it where it joins fsim_step_host's upload, and the next particle-to-grid transfer queues behind it on the second
    // stream anyway.
    struct SplitGuard {
        Sim* s;
        bool ok;
        ~SplitGuard() { if (!(ok && s->deferFarJoin)) joinFarFill(s); s->splitFill = false; }
    } splitGuard{s, false};
    int first = 0;
    if (s->mode == FSIM_PICFLIP && s->opt.reserved[3] != 1) {
        // createWaterLevelSet and transferVelocityToGrid only share the particle sort: the level set reads the particle
        // positions and (for the statistics) the OLD grid velocities, the transfer writes the new ones.  The transfer runs
        // on the second stream into the newMac buffers (free until updateVelocity), the two pointer pairs are swapped
        // at the join; both stages are latency-bound and their CTAs fit the SMs side by side.
        int rc = sortParticlesByCell(s);
        if (rc) return rc;
        s->skipSort = true;
        CUDA_TRY(cudaEventRecord(s->evFork, s->stream));
        CUDA_TRY(cudaStreamWaitEvent(s->stream2, s->evFork, 0));
        cudaStream_t mainStream = s->stream;
        std::swap(s->u, s->nu); std::swap(s->v, s->nv);
        s->stream = s->stream2;
        rc = runStage(s, FSIM_STAGE_TRANSFER_VELOCITY_TO_GRID);
        s->stream = mainStream;
        std::swap(s->u, s->nu); std::swap(s->v, s->nv);
        if (rc) { s->skipSort = false; return rc; }
        CUDA_TRY(cudaEventRecord(s->evJoin, s->stream2));
        rc = runStage(s, FSIM_STAGE_CREATE_WATER_LEVEL_SET);
        s->skipSort = false;
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(s->stageEv[1], s->stream));
        if ((rc = mirrorAfterStage(s, FSIM_STAGE_CREATE_WATER_LEVEL_SET))) return rc;
        CUDA_TRY(cudaStreamWaitEvent(s->stream, s->evJoin, 0));
        std::swap(s->u, s->nu); std::swap(s->v, s->nv);  // mac = what the transfer produced
        CUDA_TRY(cudaEventRecord(s->stageEv[2], s->stream));
        first = 2;
        s->prepPending = s->opt.reserved[6] != 1;
    }
    const int forkAfter = s->mode == FSIM_SEMILAGRANGIAN ? FSIM_STAGE_CREATE_WATER_LEVEL_SET : FSIM_STAGE_TRANSFER_VELOCITY_TO_GRID;
    for (int k = first; k < n; ++k) {
        if (order[k] == FSIM_STAGE_UPDATE_VELOCITY) {
            if (s->prepPending) { int rp = forkExtrapolationPrepare(s); if (rp) return rp; }  // (the projection did not start it)
            if (s->extrapReady) CUDA_TRY(cudaStreamWaitEvent(s->stream, s->evPrep, 0));
        }
        int rc = runStage(s, order[k]);
        if (rc) { s->extrapReady = false; s->prepPending = false; return rc; }
        if (order[k] == forkAfter) s->prepPending = s->opt.reserved[6] != 1;
        if (k == n - 1 && !s->deferFarJoin && (rc = joinFarFill(s))) return rc;  // the frame ends when the far layers are in
        CUDA_TRY(cudaEventRecord(s->stageEv[k + 1], s->stream));
        if ((rc = joinUpload(s))) return rc;  // no-op unless the stage returned without reading u, v
        if ((rc = mirrorAfterStage(s, order[k]))) return rc;
    }
    s->numStages = n;
    s->currentTime += s->dt;
    splitGuard.ok = true;
    return FSIM_OK;
}

}  // namespace

// The labels are final once the level set is done (and, in PIC/FLIP mode, the extrapolation buffers are free once the
// particle-to-grid transfer is): from here the structure of updateVelocity's extrapolation is built on the second
// stream beside the projection.  runFrame only marks it pending; stageApplyProjection starts it once its set-up and
// the first batch of iterations are queued, so that it runs beside the latency-bound triangular solves rather than
// beside the bandwidth-bound assembly / factorisation (which it would only slow down).
int forkExtrapolationPrepare(Sim* s) {
    s->prepPending = false;
    CUDA_TRY(cudaEventRecord(s->evFork, s->stream));
    CUDA_TRY(cudaStreamWaitEvent(s->stream2, s->evFork, 0));
    cudaStream_t mainStream = s->stream;
    s->stream = s->stream2;
    int rc = prepareVelocityExtrapolation(s);
    s->stream = mainStream;
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(s->evPrep, s->stream2));
    s->extrapReady = true;
    return FSIM_OK;
}

int joinFarFill(Sim* s) {
    if (!s->farPending) return FSIM_OK;
    s->farPending = false;
    CUDA_TRY(cudaStreamWaitEvent(s->stream, s->evFar, 0));
    return FSIM_OK;
}

int joinUpload(Sim* s) {
    if (!s->uploadPending) return FSIM_OK;
    s->uploadPending = false;
    CUDA_TRY(cudaStreamWaitEvent(s->stream, s->evUpload, 0));
    return FSIM_OK;
}

int fillHandSentinel(Sim* s) {
    fillU64Kernel<<<296, 256, 0, s->stream>>>(s->hand, s->handWords, wf::SENT);
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}

extern "C" int fsim_create(const fsim_config* cfg, const fsim_options* optIn, fsim_handle* out) {
    if (!cfg || !out || !cfg->initialValues) { fsim_set_error("null argument"); return FSIM_E_INVALID; }
    if (cfg->sizeX < 4 || cfg->sizeY < 4 || cfg->sizeX > 32768 || cfg->sizeY > 32768 || cfg->particlesPerCellSqrt < 1 ||
        !(cfg->dx > 0) || !(cfg->dt > 0) || !(cfg->rho > 0)) {
        fsim_set_error("invalid configuration");
        return FSIM_E_INVALID;
    }
    // the reference reads cell(i-1,j) etc. without bounds checks: the border must be SOLID
    for (int i = 0; i < cfg->sizeX; ++i)
        if (cfg->initialValues[i] != FSIM_CELL_SOLID || cfg->initialValues[(size_t)(cfg->sizeY - 1) * cfg->sizeX + i] != FSIM_CELL_SOLID) {
            fsim_set_error("domain border must be SOLID");
            return FSIM_E_INVALID;
        }
    for (int j = 0; j < cfg->sizeY; ++j)
        if (cfg->initialValues[(size_t)j * cfg->sizeX] != FSIM_CELL_SOLID || cfg->initialValues[(size_t)j * cfg->sizeX + cfg->sizeX - 1] != FSIM_CELL_SOLID) {
            fsim_set_error("domain border must be SOLID");
            return FSIM_E_INVALID;
        }
    fsim_options opt;
    if (optIn) opt = *optIn; else fsim_default_options(&opt);
    if (cfg->mode != FSIM_SEMILAGRANGIAN && cfg->mode != FSIM_PICFLIP) { fsim_set_error("unknown mode %d", cfg->mode); return FSIM_E_INVALID; }
    if (!(opt.pcgTol > 0) || opt.pcgMaxIters < 1) { fsim_set_error("bad PCG parameters (tol %g, maxIters %d)", opt.pcgTol, opt.pcgMaxIters); return FSIM_E_INVALID; }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        fsim_set_error("no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(ce));
        return FSIM_E_CUDA;
    }
    if (opt.device < 0 || opt.device >= ndev) { fsim_set_error("bad device ordinal %d", opt.device); return FSIM_E_INVALID; }
    CUDA_TRY(cudaSetDevice(opt.device));

    Sim* s = new Sim();
    s->nx = cfg->sizeX; s->ny = cfg->sizeY; s->ppcSqrt = cfg->particlesPerCellSqrt; s->mode = cfg->mode;
    s->dt = cfg->dt; s->dx = cfg->dx; s->dr = 0.9 * cfg->dx; s->rho = cfg->rho; s->gx = cfg->gravityX; s->gy = cfg->gravityY;
    s->alpha = cfg->picFlipAlpha; s->currentTime = 0.0;
    s->opt = opt; s->device = opt.device;
    s->fr = makeFrame(s->nx, s->ny);
    s->np = 0; s->npCap = 0; s->pos = nullptr; s->vel = nullptr; s->pcell = nullptr; s->sortedIdx = nullptr;
    s->launches = 0; s->numStages = 0; s->statsValid = false; s->dbgState = nullptr; s->profile = false; s->profUsed = 0;
    memset(s->stageMs, 0, sizeof(s->stageMs));
    int rc = FSIM_OK;
#define TRY(x) do { if ((rc = (x)) != FSIM_OK) { fsim_destroy(reinterpret_cast<fsim_handle>(s)); return rc; } } while (0)
#define CTRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { fsim_set_error("%s -> %s", #x, cudaGetErrorString(_e)); fsim_destroy(reinterpret_cast<fsim_handle>(s)); return FSIM_E_CUDA; } } while (0)
    {
        // the main stream carries the latency-critical wavefront kernels: its CTAs go first when the second stream's
        // side work (particle-to-grid transfer, extrapolation structure) competes for the SMs
        int prLow = 0, prHigh = 0;
        CTRY(cudaDeviceGetStreamPriorityRange(&prLow, &prHigh));
        CTRY(cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, prHigh));
        CTRY(cudaStreamCreateWithPriority(&s->stream2, cudaStreamNonBlocking, prLow));
    }
    CTRY(cudaEventCreateWithFlags(&s->evFork, cudaEventDisableTiming));
    CTRY(cudaEventCreateWithFlags(&s->evJoin, cudaEventDisableTiming));
    CTRY(cudaEventCreateWithFlags(&s->evPrep, cudaEventDisableTiming));
    CTRY(cudaEventCreateWithFlags(&s->evNear, cudaEventDisableTiming));
    CTRY(cudaEventCreateWithFlags(&s->evFar, cudaEventDisableTiming));
    s->extrapReady = false;
    CTRY(cudaStreamCreateWithFlags(&s->copyStream, cudaStreamNonBlocking));
    CTRY(cudaStreamCreateWithFlags(&s->axpyStream, cudaStreamNonBlocking));
    CTRY(cudaEventCreateWithFlags(&s->evAxpyA, cudaEventDisableTiming));
    CTRY(cudaEventCreateWithFlags(&s->evAxpyP, cudaEventDisableTiming));
    CTRY(cudaEventCreateWithFlags(&s->evUpload, cudaEventDisableTiming));
    CTRY(cudaEventCreateWithFlags(&s->evMirror, cudaEventDisableTiming));
    for (int k = 0; k < 8; ++k) CTRY(cudaEventCreateWithFlags(&s->evChunk[k], cudaEventDisableTiming));
    s->mirror = nullptr; s->mirrorOverlap = false; s->uploadPending = false; s->mirrorDone = 0;
    s->skipSort = false;
    double** dbl[] = {&s->u, &s->v, &s->nu, &s->nv, &s->p, &s->phi, &s->phiTmp, &s->Adiag, &s->Ax, &s->Ay, &s->rhs, &s->fmask,
                      &s->pc, &s->D, &s->Ux, &s->Uy, &s->Lx, &s->Ly, &s->r, &s->z, &s->s, &s->t, &s->lsPx, &s->lsPy, &s->lsId};
    for (double** p : dbl) TRY(allocFrame(s, p));
    s->slU = nullptr; s->slV = nullptr;
    if (s->mode == FSIM_SEMILAGRANGIAN) { TRY(allocFrame(s, &s->slU)); TRY(allocFrame(s, &s->slV)); }
    TRY(allocFrame(s, &s->cell)); TRY(allocFrame(s, &s->unkU)); TRY(allocFrame(s, &s->unkV));
    {
        int* raw;
        TRY(allocLinear(s, &raw, s->fr.elems)); s->distU = raw;
        TRY(allocLinear(s, &raw, s->fr.elems)); s->distV = raw;
        TRY(allocLinear(s, &raw, 2 * s->fr.elems)); s->distTmp = raw;
    }
    TRY(allocLinear(s, &s->layerCellsU, s->fr.elems));
    TRY(allocLinear(s, &s->layerCellsV, s->fr.elems));
    TRY(allocLinear(s, &s->layerMaskU, s->fr.elems));
    TRY(allocLinear(s, &s->layerMaskV, s->fr.elems));
    TRY(allocLinear(s, &s->layerConsU, s->fr.elems));
    TRY(allocLinear(s, &s->layerConsV, s->fr.elems));
    s->maxLayers = s->nx + s->ny + 8;
    TRY(allocLinear(s, &s->layerStartU, (size_t)(s->maxLayers + 2) * 2));
    TRY(allocLinear(s, &s->layerStartV, (size_t)(s->maxLayers + 2) * 2));
    {
        // PCG layout: two rows per lane with skew 1 (solveKernelR) by default; FSIM_SD_RPL=1 selects one row per lane
        // with skew 2 (solveKernel), FSIM_SD_SIGMA its skew
        int rpl = 2, sigma = 1;
        if (const char* e = getenv("FSIM_SD_RPL")) { if (atoi(e) == 1) rpl = 1; }
        if (rpl == 1) {
            sigma = opt.reserved[0] >= 2 && opt.reserved[0] <= 3 ? opt.reserved[0] : 2;
            if (const char* e = getenv("FSIM_SD_SIGMA")) { int v = atoi(e); if (v >= 2 && v <= 3) sigma = v; }  // tuning knob
        } else {
            // two rows per lane: lane skew 1 (the neighbour lane's value arrives by shuffle ON the dependent chain) or 2/3
            // (the shuffle gets a step or two of slack: the chain of a step is the two dependent DFMA only)
            if (opt.reserved[0] >= 1 && opt.reserved[0] <= 3) sigma = opt.reserved[0];
            if (const char* e = getenv("FSIM_SD_SIGMA")) { int v = atoi(e); if (v >= 1 && v <= 3) sigma = v; }  // tuning knob
        }
        s->sdg = sd::makeGeom(s->nx, s->ny, sigma, rpl);
        s->swg = sd::makeGeom(s->nx, s->ny, 1);
        size_t sdElems = s->sdg.elems;  // the arrays also serve the sweeps (skew 1, one row per lane) as scratch
        if (s->swg.elems > sdElems) sdElems = s->swg.elems;
        sdElems += (size_t)3 * s->sdg.Sp * 32 * rpl;  // a y-slab adds a halo strip on each side
        double** sdArr[] = {&s->sAd, &s->sAx, &s->sAy, &s->sLx, &s->sLy, &s->sD, &s->sUx, &s->sUy, &s->sR, &s->sP, &s->sS, &s->sZ, &s->sT};
        for (double** p : sdArr) TRY(allocLinear(s, p, sdElems));
        s->sdHandWords = sd::handWords(s->sdg);
        TRY(allocLinear(s, &s->sdHand, s->sdHandWords));
        TRY(allocLinear(s, &s->sdRange, (size_t)2 * (s->ny / 32 + 4)));
        fillU64Kernel<<<296, 256, 0, s->stream>>>(s->sdHand, s->sdHandWords, sd::SENT);
        LAUNCH_COUNT(s);
        s->swPlaneWords = sd::handWords(s->swg);
        TRY(allocLinear(s, &s->swHand, 3 * s->swPlaneWords));
        fillU64Kernel<<<296, 256, 0, s->stream>>>(s->swHand, 3 * s->swPlaneWords, sd::SENT);
        LAUNCH_COUNT(s);
    }
    {
        const size_t tiles = (size_t)(s->nx / 32 + 2) * (s->ny / 32 + 2);
        for (int k = 0; k < 2; ++k) {
            TRY(allocLinear(s, &s->lsTileNeg[k], tiles));
            TRY(allocLinear(s, &s->lsTileStamp[k], tiles));
        }
    }
    size_t ncells = (size_t)s->nx * s->ny;
    TRY(allocLinear(s, &s->cellStart, ncells + 1));
    TRY(allocLinear(s, &s->cellCursor, ncells));
    TRY(allocLinear(s, &s->scanTmp, ncells / 2048 + 2));
    TRY(allocLinear(s, &s->ctl, 1));
    size_t relabelBlocks = (size_t)((s->nx + 31) / 32) * ((s->ny + 7) / 8);
    size_t aaBlocks = (size_t)(s->sdg.Sp / 8 + 1) * s->sdg.nstrips;
    TRY(allocLinear(s, &s->partials, (4 * relabelBlocks > aaBlocks ? 4 * relabelBlocks : aaBlocks) + 4096));  // (diagnosticsKernel: 4 planes)
    TRY(allocLinear(s, &s->counters, 16));
    TRY(allocLinear(s, &s->wfTicket, 4));
    s->wfFinished = s->wfTicket + 1;
    TRY(allocLinear(s, &s->slProgress, (size_t)s->fr.H + 64));
    s->handWords = (size_t)4 * (s->fr.H / 32) * s->fr.W;
    TRY(allocLinear(s, &s->hand, s->handWords));
    TRY(fillHandSentinel(s));
    if (opt.debugSimpleWavefront) TRY(allocLinear(s, &s->dbgState, (size_t)4 * s->fr.W * s->fr.H));
    CTRY(cudaMallocHost(&s->hctl, sizeof(DevCtl)));
    CTRY(cudaMallocHost(&s->hPcgFlags, 8 * sizeof(int) + 4 * sizeof(double)));
    s->hBox = s->hPcgFlags + 4;
    s->hDiag = reinterpret_cast<double*>(s->hPcgFlags + 8);
    TRY(allocLinear(s, &s->dDiag, 4));
    CTRY(cudaEventCreate(&s->evT0));
    CTRY(cudaEventCreate(&s->evT1));
    s->lastSolveCells = 0;
    for (int k = 0; k < 2; ++k) CTRY(cudaEventCreateWithFlags(&s->pollEv[k], cudaEventDisableTiming));
    for (int k = 0; k < 10; ++k) CTRY(cudaEventCreate(&s->stageEv[k]));
    TRY(pcgSetParams(s, opt.pcgTol, opt.pcgMaxIters));

    // cell labels
    CTRY(cudaMemcpy2DAsync(s->cell, s->fr.pitch, cfg->initialValues, s->nx, s->nx, s->ny, cudaMemcpyHostToDevice, s->stream));
    // particles
    if (opt.seedParticles) {
        std::vector<double> pos;
        seedParticles(cfg, pos);
        size_t n = pos.size() / 2;
        TRY(allocParticles(s, n));
        s->np = n;
        if (n) {
            CTRY(cudaMemcpyAsync(s->pos, pos.data(), n * 16, cudaMemcpyHostToDevice, s->stream));
            CTRY(cudaMemsetAsync(s->vel, 0, n * 16, s->stream));
        }
        CTRY(cudaStreamSynchronize(s->stream));  // `pos` goes out of scope
    } else {
        TRY(allocParticles(s, 1024));
    }
    CTRY(cudaStreamSynchronize(s->stream));
#undef TRY
#undef CTRY
    *out = reinterpret_cast<fsim_handle>(s);
    return FSIM_OK;
}

extern "C" int fsim_destroy(fsim_handle h) {
    if (!h) return FSIM_OK;
    Sim* s = reinterpret_cast<Sim*>(h);
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->stream2) { cudaStreamSynchronize(s->stream2); cudaStreamDestroy(s->stream2); }
    if (s->copyStream) { cudaStreamSynchronize(s->copyStream); cudaStreamDestroy(s->copyStream); }
    if (s->axpyStream) { cudaStreamSynchronize(s->axpyStream); cudaStreamDestroy(s->axpyStream); }
    if (s->evAxpyA) cudaEventDestroy(s->evAxpyA);
    if (s->evAxpyP) cudaEventDestroy(s->evAxpyP);
    if (s->evUpload) cudaEventDestroy(s->evUpload);
    if (s->evMirror) cudaEventDestroy(s->evMirror);
    for (int k = 0; k < 8; ++k) if (s->evChunk[k]) cudaEventDestroy(s->evChunk[k]);
    if (s->evFork) cudaEventDestroy(s->evFork);
    if (s->evPrep) cudaEventDestroy(s->evPrep);
    if (s->evNear) cudaEventDestroy(s->evNear);
    if (s->evFar) cudaEventDestroy(s->evFar);
    if (s->evJoin) cudaEventDestroy(s->evJoin);
    if (s->evT0) cudaEventDestroy(s->evT0);
    if (s->evT1) cudaEventDestroy(s->evT1);
    distDestroy(s);
    for (void* p : s->rawAllocs) if (p) cudaFree(p);
    if (s->hctl) cudaFreeHost(s->hctl);
    if (s->hPcgFlags) cudaFreeHost(s->hPcgFlags);
    for (int k = 0; k < 2; ++k) if (s->pollEv[k]) cudaEventDestroy(s->pollEv[k]);
    for (int k = 0; k < 10; ++k) if (s->stageEv[k]) cudaEventDestroy(s->stageEv[k]);
    for (auto e : s->profEv) if (e) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return FSIM_OK;
}

#define HANDLE(h)                                                              \
    if (!(h)) { fsim_set_error("null handle"); return FSIM_E_INVALID; }        \
    Sim* s = reinterpret_cast<Sim*>(h);                                        \
    CUDA_TRY(cudaSetDevice(s->device));

extern "C" int fsim_step(fsim_handle h, int nsteps) {
    HANDLE(h);
    for (int k = 0; k < nsteps; ++k) {
        s->deferFarJoin = k + 1 < nsteps;  // (runFrame: the last frame of the call ends with everything joined)
        int rc = runFrame(s);
        s->deferFarJoin = false;
        if (rc) return rc;
    }
    return FSIM_OK;
}

// n updates bracketed by CUDA events on the simulation's main stream (every side stream joins it before a frame ends)
extern "C" int fsim_step_timed(fsim_handle h, int nsteps, double* deviceMs) {
    HANDLE(h);
    CUDA_TRY(cudaEventRecord(s->evT0, s->stream));
    for (int k = 0; k < nsteps; ++k) {
        s->deferFarJoin = k + 1 < nsteps;
        int rc = runFrame(s);
        s->deferFarJoin = false;
        if (rc) return rc;
    }
    CUDA_TRY(cudaEventRecord(s->evT1, s->stream));
    CUDA_TRY(cudaEventSynchronize(s->evT1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, s->evT0, s->evT1));
    if (deviceMs) *deviceMs = ms;
    return FSIM_OK;
}

// FluidSim2D::avgPressure / avgPressureInFluid / maxVelocity (src/FluidSim2D.cpp:607-638) from the device-resident state
extern "C" int fsim_diagnostics(fsim_handle h, double* avgPressure, double* avgPressureInFluid, double* maxVelocity) {
    HANDLE(h);
    GridView g{s->u, s->v, s->nx, s->ny, s->fr.pitch, s->dx};
    dim3 grd((s->nx + 31) / 32, (s->ny + 7) / 8);
    diagnosticsKernel<<<grd, 256, 0, s->stream>>>(s->p, s->cell, g, s->partials, &s->counters[8], s->dDiag);
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(s->hDiag, s->dDiag, 4 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    // (the reference divides by the int product sizeX*sizeY and, in the fluid, by a count that may be zero: NaN then)
    if (avgPressure) *avgPressure = s->hDiag[0] / (double)(s->nx * s->ny);
    if (avgPressureInFluid) *avgPressureInFluid = s->hDiag[1] / s->hDiag[2];
    if (maxVelocity) *maxVelocity = s->hDiag[3];
    return FSIM_OK;
}

extern "C" int fsim_dist_unique_id(void* out128) { return distGetUniqueId(out128); }

extern "C" int fsim_dist_init(fsim_handle h, int rank, int world, const void* uniqueId128) {
    HANDLE(h);
    return distInit(s, rank, world, uniqueId128);
}

extern "C" int fsim_stage(fsim_handle h, int stage) {
    HANDLE(h);
    return runStage(s, stage);
}

extern "C" int fsim_sync(fsim_handle h) {
    HANDLE(h);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return FSIM_OK;
}

extern "C" int fsim_num_particles(fsim_handle h, size_t* n) {
    HANDLE(h);
    *n = s->np;
    return FSIM_OK;
}

extern "C" int fsim_upload(fsim_handle h, int field, const void* src, size_t bytes) {
    HANDLE(h);
    FieldInfo fi;
    if (fieldInfo(s, field, &fi) || field >= FSIM_ADIAG) { fsim_set_error("field %d is not uploadable", field); return FSIM_E_INVALID; }
    if (fi.NX == 0) {
        if (bytes != fi.linearCount * fi.elem) { fsim_set_error("size mismatch for field %d", field); return FSIM_E_INVALID; }
        if (bytes) CUDA_TRY(cudaMemcpyAsync(fi.base, src, bytes, cudaMemcpyHostToDevice, s->stream));
    } else {
        if (bytes != (size_t)fi.NX * fi.NY * fi.elem) { fsim_set_error("size mismatch for field %d", field); return FSIM_E_INVALID; }
        CUDA_TRY(cudaMemcpy2DAsync(fi.base, s->fr.pitch * fi.elem, src, fi.NX * fi.elem, fi.NX * fi.elem, fi.NY,
                                   cudaMemcpyHostToDevice, s->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));  // src may be pageable and reused by the caller
    return FSIM_OK;
}

extern "C" int fsim_download(fsim_handle h, int field, void* dst, size_t bytes) {
    HANDLE(h);
    FieldInfo fi;
    if (fieldInfo(s, field, &fi)) { fsim_set_error("unknown field %d", field); return FSIM_E_INVALID; }
    if (fi.NX == 0) {
        if (bytes != fi.linearCount * fi.elem) { fsim_set_error("size mismatch for field %d", field); return FSIM_E_INVALID; }
        if (bytes) CUDA_TRY(cudaMemcpyAsync(dst, fi.base, bytes, cudaMemcpyDeviceToHost, s->stream));
    } else {
        if (bytes != (size_t)fi.NX * fi.NY * fi.elem) { fsim_set_error("size mismatch for field %d", field); return FSIM_E_INVALID; }
        CUDA_TRY(cudaMemcpy2DAsync(dst, fi.NX * fi.elem, fi.base, s->fr.pitch * fi.elem, fi.NX * fi.elem, fi.NY,
                                   cudaMemcpyDeviceToHost, s->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return FSIM_OK;
}

extern "C" int fsim_set_particles(fsim_handle h, size_t n, const double* pos, const double* vel) {
    HANDLE(h);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    int rc = allocParticles(s, n);
    if (rc) return rc;
    s->np = n;
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(s->pos, pos, n * 16, cudaMemcpyHostToDevice, s->stream));
        if (vel) CUDA_TRY(cudaMemcpyAsync(s->vel, vel, n * 16, cudaMemcpyHostToDevice, s->stream));
        else CUDA_TRY(cudaMemsetAsync(s->vel, 0, n * 16, s->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return FSIM_OK;
}

extern "C" int fsim_set_params(fsim_handle h, double gx, double gy, double alpha, double dt) {
    HANDLE(h);
    s->gx = gx; s->gy = gy; s->alpha = alpha; s->dt = dt;
    return FSIM_OK;
}

extern "C" int fsim_set_pcg(fsim_handle h, double tol, int maxIters) {
    HANDLE(h);
    if (!(tol > 0) || maxIters < 1) { fsim_set_error("bad PCG parameters"); return FSIM_E_INVALID; }
    s->opt.pcgTol = tol; s->opt.pcgMaxIters = maxIters;
    return pcgSetParams(s, tol, maxIters);
}

extern "C" int fsim_get_stats(fsim_handle h, fsim_stats* out) {
    HANDLE(h);
    CUDA_TRY(cudaMemcpyAsync(s->hctl, s->ctl, sizeof(DevCtl), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    memset(out, 0, sizeof(*out));
    const DevCtl& c = *s->hctl;
    out->waterVolume = c.fluidCells * (s->dx * s->dx);
    out->totalEnergy = c.gridEnergy;
    out->particleTotalEnergy = c.particleEnergy;
    out->currentTime = s->currentTime;
    out->pcgIters = c.iter;
    out->pcgHitMaxIters = c.hitMax;
    out->pcgResidual = c.rnorm;
    out->pcgRhsNorm = c.rhsNorm;
    out->cflMax = c.cflMax;
    out->nanPositions = c.nanCount;
    out->levelSetSweeps = c.sweepsRun;
    out->extrapolationLayers = c.maxLayer[0] > c.maxLayer[1] ? c.maxLayer[0] : c.maxLayer[1];
    out->numStages = s->numStages;
    out->pcgSolveCells = s->lastSolveCells;
    out->pcgMarchedCells = (long long)c.marchedSlots;
    out->distError = c.distError;
    out->extrapolationNearLayers = s->lastSplit ? c.nearLayers : 0;
    for (int k = 0; k < s->numStages && k < 8; ++k) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s->stageEv[k], s->stageEv[k + 1]) == cudaSuccess) out->stageMs[k] = ms;
    }
    cudaGetLastError();
    return FSIM_OK;
}

// true if every buffer of the mirror is page-locked: only then do async copies on another stream return at once
static bool mirrorPinned(const fsim_host_mirror* io) {
    const void* ptrs[] = {io->u_in, io->v_in, io->u, io->v, io->p, io->cell, io->phi, io->particles, io->particleVels};
    for (const void* q : ptrs) {
        if (!q) continue;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, q) != cudaSuccess) { cudaGetLastError(); return false; }
        if (a.type != cudaMemoryTypeHost) return false;
    }
    return true;
}

extern "C" int fsim_step_host(fsim_handle h, const fsim_host_mirror* io) {
    HANDLE(h);
    if (!io) { fsim_set_error("null io"); return FSIM_E_INVALID; }
    const Frame& f = s->fr;
    const int nx = s->nx, ny = s->ny;
    // pinned mirrors: the copies that do not depend on the end of the frame leave the critical path (copy stream);
    // pageable mirrors (or fsim_options.reserved[4] = 1): everything in order on the one stream
    s->mirrorOverlap = s->opt.reserved[4] != 1 && mirrorPinned(io);
    cudaStream_t up = s->stream;
    if (s->mirrorOverlap && (io->u_in || io->v_in)) {
        CUDA_TRY(cudaEventRecord(s->evMirror, s->stream));  // behind whatever the caller queued before
        CUDA_TRY(cudaStreamWaitEvent(s->copyStream, s->evMirror, 0));
        up = s->copyStream;
    }
    if (io->u_in) CUDA_TRY(cudaMemcpy2DAsync(s->u, f.pitch * 8, io->u_in, (nx + 1) * 8, (nx + 1) * 8, ny, cudaMemcpyHostToDevice, up));
    if (io->v_in) CUDA_TRY(cudaMemcpy2DAsync(s->v, f.pitch * 8, io->v_in, nx * 8, nx * 8, ny + 1, cudaMemcpyHostToDevice, up));
    if (up != s->stream) {
        CUDA_TRY(cudaEventRecord(s->evUpload, up));
        s->uploadPending = true;  // joined by the first stage that reads u, v (the level set's statistics)
    }
    s->mirror = io;
    s->mirrorDone = 0;
    int rc = runFrame(s);
    if (!rc) rc = joinUpload(s);
    if (!rc) rc = mirrorDownload(s, M_ALL);
    s->mirror = nullptr;
    s->uploadPending = false;
    cudaError_t e1 = cudaStreamSynchronize(s->stream);
    cudaError_t e2 = s->mirrorOverlap ? cudaStreamSynchronize(s->copyStream) : cudaSuccess;
    if (rc) return rc;
    CUDA_TRY(e1);
    CUDA_TRY(e2);
    return FSIM_OK;
}

extern "C" int fsim_host_register(fsim_handle h, void* ptr, size_t bytes) {
    HANDLE(h);
    if (!ptr || !bytes) { fsim_set_error("null buffer"); return FSIM_E_INVALID; }
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();  // not sticky: the caller may carry on with pageable memory
        fsim_set_error("cudaHostRegister(%zu bytes) -> %s", bytes, cudaGetErrorString(e));
        return FSIM_E_CUDA;
    }
    return FSIM_OK;
}

extern "C" int fsim_host_unregister(fsim_handle h, void* ptr) {
    HANDLE(h);
    if (!ptr) { fsim_set_error("null buffer"); return FSIM_E_INVALID; }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fsim_set_error("cudaHostUnregister -> %s", cudaGetErrorString(e));
        return FSIM_E_CUDA;
    }
    return FSIM_OK;
}

extern "C" int fsim_profile_enable(fsim_handle h, int on) {
    HANDLE(h);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (on && s->profEv.empty()) {
        s->profEv.resize(2 * 8192);
        for (auto& e : s->profEv) CUDA_TRY(cudaEventCreate(&e));
    }
    s->profile = on != 0;
    s->profUsed = 0;
    s->profClass.clear();
    return FSIM_OK;
}

extern "C" int fsim_profile_get(fsim_handle h, int klass, double* totalMs, int* launches) {
    HANDLE(h);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    double tot = 0.0;
    int n = 0;
    for (size_t k = 0; k < s->profUsed / 2; ++k) {
        if (s->profClass[k] != klass) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s->profEv[2 * k], s->profEv[2 * k + 1]) == cudaSuccess) { tot += ms; ++n; }
    }
    cudaGetLastError();
    if (totalMs) *totalMs = tot;
    if (launches) *launches = n;
    return FSIM_OK;
}

// the individual launch durations of one class, in launch order (diagnostics: which sweep of a stage costs what)
extern "C" int fsim_profile_list(fsim_handle h, int klass, double* ms, int cap, int* count) {
    HANDLE(h);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    int n = 0;
    for (size_t k = 0; k < s->profUsed / 2; ++k) {
        if (s->profClass[k] != klass) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, s->profEv[2 * k], s->profEv[2 * k + 1]) != cudaSuccess) continue;
        if (n < cap && ms) ms[n] = t;
        ++n;
    }
    cudaGetLastError();
    if (count) *count = n;
    return FSIM_OK;
}

extern "C" int fsim_launch_count(fsim_handle h, unsigned long long* n) {
    HANDLE(h);
    *n = s->launches;
    return FSIM_OK;
}
