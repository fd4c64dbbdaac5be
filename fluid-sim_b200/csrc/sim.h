// Host-side state of one simulation handle and the stage entry points implemented in the .cu files.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/fsim.h"
#include "common.cuh"
#include "sdwave.cuh"

// Device-resident control block: every scalar the step needs lives here so that no stage has to
// synchronise with the host (the PCG loop polls `pcgDone` asynchronously, one batch behind).
struct DevCtl {
    // PCG (src/FluidSim2D.cpp:423-466)
    double sigma, zs, alpha, beta, rnorm, rhsNorm, tol;
    int iter, maxIters, pcgDone, hitMax;
    // level set (construct, redistance): lsChanged[kind][t] = sweep t changed a cell; lsGate[kind][t] = a later sweep can
    // still change something (construct: sweep t changed a cell -- its visit is the same function in every direction, so one
    // clean sweep is a fixed point; redistance: one of the sweeps t-2..t did -- a direction's sweep is idempotent, so it is a
    // no-op when the three sweeps since its last run changed nothing).  Gated-off sweeps are exact no-ops.
    int lsChanged[2][16];
    int lsGate[2][16];
    int sweepsRun;
    int lsMirror[2];  // x-mirror flag of the SD copy the sweeps of each kind last ran on (gated-off layout switches keep the old one)
    // extrapolation
    int anyKnown[2];
    int maxLayer[2];
    // particle diagnostics
    int nanCount;
    double cflMax;
    // statistics (src/FluidSim2D.cpp:709-731)
    double fluidCells, gridEnergy, particleEnergy;
    // semi-Lagrangian skew
    double maxDisp;
    // y-slab PCG (distpeer.cuh): a wait on a peer rank timed out (the solve was stopped)
    int distError;
    double reservedD;
    int bbox[4];  // FLUID cells: min i, max i, min j, max j (this step's projection)
    int lsBox[4]; // cells the eikonal sweeps can change (phi < 0)
    int alphaIter;  // PCG: 1-based iteration whose alpha is in `alpha` (stamped by applyA; axpyPKernel runs iff it matches)
    int slOverflow; // exact semi-Lagrangian advection: a footprint left the window (sl.cu)
    int pendingP; // fused PCG: the loop ended before the backward solve could apply p += alpha s (pcgFinishKernel does)
    unsigned long long marchedSlots;  // layout slots the triangular solves march (chunks that hold fluid)
    // split extrapolation after updateVelocity (projection.cu stageUpdateVelocity): bits of the largest |velocity| on a known
    // face, and the number of BFS layers the particle stages can reach (the rest is filled beside them)
    unsigned long long vmaxBits;
    int nearLayers;
    int partDist;  // largest BFS layer of a face of a cell that holds particles but is not FLUID (0: every particle sits in a FLUID cell)
};

struct Sim {
    int nx, ny, ppcSqrt, mode;
    double dt, dx, dr, rho, gx, gy, alpha, currentTime;
    fsim_options opt;
    Frame fr;
    int device;
    cudaStream_t stream;
    // second stream: in PIC/FLIP frames the particle-to-grid transfer (with its extrapolation) runs beside the level set --
    // both are latency-bound and together still fit the SMs; skipSort: the frame has sorted the particles already
    cudaStream_t stream2;
    cudaEvent_t evFork, evJoin;
    bool skipSort;
    // the structure of updateVelocity's extrapolation (masks, distances, layer lists) depends on the labels only: runFrame
    // builds it on the second stream beside the projection (evPrep = done; extrapReady = stageUpdateVelocity may skip it)
    cudaEvent_t evPrep;
    bool extrapReady, prepPending;
    // runFrame only: the extrapolation after updateVelocity is cut at DevCtl::nearLayers -- the far layers are filled on the
    // second stream beside the particle stages (farPending until runFrame joins it; evNear / evFar order the two streams)
    bool splitFill, farPending, deferFarJoin, lastSplit;
    cudaEvent_t evNear, evFar;
    // fsim_step_host with pinned mirrors: the uploads, and the downloads of fields that are final before the frame ends,
    // run on a copy stream beside the stages (`mirror` is set only inside such a call; mirrorDone = M_* bits issued)
    cudaStream_t copyStream;
    cudaEvent_t evUpload, evMirror;
    cudaEvent_t evChunk[8];  // chunked particle advection: one per chunk, its download waits for it (particles.cu)
    const fsim_host_mirror* mirror;
    bool mirrorOverlap, uploadPending;
    unsigned mirrorDone;

    // frame-shaped double arrays; pointers address logical (0,0)
    double *u, *v, *nu, *nv, *p, *phi, *phiTmp;
    double *Adiag, *Ax, *Ay, *rhs, *fmask, *pc, *D, *Ux, *Uy, *Lx, *Ly, *r, *z, *s, *t;
    double *lsPx, *lsPy, *lsId;
    // PCG state in the strip-diagonal layout (sdwave.cuh)
    sd::Geom sdg;
    double *sAd, *sAx, *sAy, *sLx, *sLy, *sD, *sUx, *sUy, *sR, *sP, *sS, *sZ, *sT;
    unsigned long long* sdHand;
    size_t sdHandWords;
    int* sdRange;  // [2 * strips]: first / last storage chunk of each strip that holds fluid (sd::Control::range)
    // in-place sweeps (level set, MIC(0) factor) run on an SD layout of their own skew (sdsweep.cuh); the PCG's SD
    // arrays serve as scratch, the hand-off slots are separate (up to 3 planes)
    sd::Geom swg;
    unsigned long long* swHand;
    size_t swPlaneWords;
    unsigned char* lsTileNeg[2];  // level-set sweeps [construct, redistance]: per (strip, 32-column block) "holds a cell the sweep may change" ...
    int* lsTileStamp[2];          // ... and the index of the last sweep that changed a cell there (sdsweep.cuh, OpSkip)
    double *slU, *slV;  // semi-Lagrangian snapshot of the pre-advection grid
    uint8_t *cell, *unkU, *unkV;  // labels; 1 = unknown face (extrapolation masks)
    int *distU, *distV, *distTmp;  // distTmp holds two planes
    uint32_t *layerCellsU, *layerCellsV;  // unknown faces sorted by BFS layer (frame offsets)
    uint8_t *layerMaskU, *layerMaskV;     // their smaller-layer neighbour masks
    unsigned long long *layerConsU, *layerConsV;  // and the (CTA, slot) targets of the next-layer faces that read them
    int *layerStartU, *layerStartV;       // [maxLayers+2]
    int maxLayers;
    std::vector<void*> rawAllocs;

    // particles (original order is the API order and is never permuted)
    size_t np, npCap;
    double2 *pos, *vel;
    uint32_t *pcell, *sortedIdx, *cellStart, *cellCursor, *scanTmp;

    // control
    DevCtl* ctl;
    DevCtl* hctl;  // pinned mirror
    double* partials;
    unsigned int* counters;  // last-block counters (zero between launches)
    int *wfTicket, *wfFinished;
    unsigned long long* hand;
    size_t handWords;
    double* dbgState;
    int* slProgress;

    // PCG polling
    int* hPcgFlags;  // pinned [2*slots]
    int* hBox;       // pinned [4]: bounding box of the fluid cells (read back once per projection)
    void* renderBuf = nullptr; size_t renderBytes = 0;  // fsim_render_fill's device staging (render.cu)
    double *hDiag, *dDiag;  // fsim_diagnostics: sum p, sum p in fluid, fluid count, max |vel| (pinned / device)
    cudaEvent_t evT0, evT1; // fsim_step_timed
    cudaStream_t axpyStream;           // p += alpha s beside the forward solve (projection.cu launchAxpy)
    cudaEvent_t evAxpyA, evAxpyP;
    int lsWin[2];    // origin of the window the eikonal sweeps run on
    long long lastSolveCells;  // cells the last projection's solve covered
    cudaEvent_t pollEv[2];

    cudaEvent_t stageEv[10];
    float stageMs[8];
    int numStages;
    unsigned long long launches;
    int lastPcgIters, lastHitMax;
    bool statsValid;

    // y-slab decomposition of the projection over the GPUs of one node (fsim_dist_init): every rank keeps the full
    // replicated state, assembles the whole system, and solves only its slab of rows [j0, j1) with block-MIC(0)
    struct Dist {
        bool on = false;
        int rank = 0, world = 1;
        void* comm = nullptr;     // ncclComm_t
        int boxStrip0 = 0, boxStrips = 0;  // strips of the fluid cells' bounding box (this step)
        int strip0 = 0, nOwn = 0; // own strips of 32 rows: [strip0, strip0 + nOwn)
        int j0 = 0, j1 = 0;       // own rows
        sd::Geom gExt, gOwn;      // slab plus one halo strip on each side / own strips only
        // peer memory (distpeer.cuh): the local PeerBlock and every rank's block as mapped into this process
        void* peerLocal = nullptr;
        void* peerBlk[16] = {};
        int ghostPitch = 0;
        unsigned int epoch = 0;   // projections so far (stamp prefix)
        int lastIters = 0;
    } dist;

    // optional per-kernel timing (fsim_profile_*)
    bool profile;
    std::vector<cudaEvent_t> profEv;  // pairs
    std::vector<int> profClass;
    size_t profUsed;
};

static inline void profBegin(Sim* s, int klass) {
    if (!s->profile || s->profUsed + 2 > s->profEv.size()) return;
    s->profClass.push_back(klass);
    cudaEventRecord(s->profEv[s->profUsed], s->stream);
}
static inline void profEnd(Sim* s) {
    if (!s->profile || s->profUsed + 2 > s->profEv.size()) return;
    cudaEventRecord(s->profEv[s->profUsed + 1], s->stream);
    s->profUsed += 2;
}

#define LAUNCH_COUNT(s) ((s)->launches++)

// stage entry points (each returns an FSIM_* code and only enqueues work on s->stream)
int stageCreateWaterLevelSet(Sim* s);
int stageTransferVelocityToGrid(Sim* s);
int stageApplySemiLagrangianAdvection(Sim* s);
int stageApplyGravity(Sim* s);
int stageApplyProjection(Sim* s);
int distInit(Sim* s, int rank, int world, const void* uniqueId);
void distDestroy(Sim* s);
int distGetUniqueId(void* out128);
int stageUpdateVelocity(Sim* s);
int stageUpdateParticleVelocities(Sim* s);
int stageApplyAdvection(Sim* s);

// shared building blocks
int joinUpload(Sim* s);  // the first reader of the grid velocities waits for fsim_step_host's upload (copy stream)
int sortParticlesByCell(Sim* s);
int extrapolatePair(Sim* s, double* a, double* b, const uint8_t* knownA, const uint8_t* knownB);
int extrapolatePrepare(Sim* s, const uint8_t* knownA, const uint8_t* knownB);  // the part that needs the masks only
int extrapolateFill(Sim* s, double* a, double* b, const uint8_t* knownA, const uint8_t* knownB, int part = -1, double* a2 = nullptr, double* b2 = nullptr);
int prepareVelocityExtrapolation(Sim* s);
int forkExtrapolationPrepare(Sim* s);  // ... on the second stream, behind what s->stream holds so far
int fillHandSentinel(Sim* s);
int copyNewMacToMac(Sim* s);
int joinFarFill(Sim* s);  // s->stream waits for the far layers (no-op unless farPending)
int particleEnergy(Sim* s);
