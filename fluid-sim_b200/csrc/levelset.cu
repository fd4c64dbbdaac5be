// createWaterLevelSet (reference src/FluidSim2D.cpp:653-732): particle level set construction
// (LevelSet::constructFromParticles, :752-794), redistancing (LevelSet::redistance, :796-920), relabelling
// (:657-664) and the per-step statistics (:709-731).
//
// The 16 closest-particle sweeps and the 16 eikonal sweeps are sequential Gauss-Seidel passes in the
// reference; they are run here by the exact wavefront scheduler (wavefront.cuh), so phi -- and with it the
// FLUID/EMPTY labels -- is reproduced bit for bit.  The floating-point expressions below use explicit
// round-to-nearest intrinsics in the contraction pattern the reference's build (-O2 -mfma) compiles to:
//   |p - node|: a = fma(-i, dx, px), b = fma(-j, dx, py), sqrt(fma(a, a, b*b)) - dr
//   eikonal:    0.5 * ((phi0 + phi1) + sqrt(fma(2dx, dx, -(phi1-phi0)^2)))
// A round of four sweeps that changes nothing leaves a fixed point, so later rounds are skipped (exact).
#include <stdlib.h>

#include "sampling.cuh"
#include "sdpack.cuh"
#include "sdsweep.cuh"
#include "wf_launch.cuh"

namespace {

constexpr unsigned long long ID_NONE = ~0ULL;  // (size_t)-1, src/FluidSim2D.cpp:760

__device__ __forceinline__ double nodeDistance(double px, double py, int i, int j, double dx, double dr) {
    double a = __fma_rn(-(double)i, dx, px);
    double b = __fma_rn(-(double)j, dx, py);
    return __dsub_rn(__dsqrt_rn(__fma_rn(a, a, __dmul_rn(b, b))), dr);
}

// :757-773 -- phi = +inf, t = none; every particle offers its distance to the lower-left node of its cell;
// the lowest particle index wins ties (strict < in index order), which the stable cell sort preserves.
__global__ void lsBinKernel(const double2* __restrict__ pos, const uint32_t* __restrict__ cellStart,
                            const uint32_t* __restrict__ sortedIdx, int nx, int ny, int pitch, double dx, double dr,
                            double* __restrict__ phi, double* __restrict__ lpx, double* __restrict__ lpy,
                            double* __restrict__ lid) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nx || j >= ny) return;
    double best = __longlong_as_double(0x7FF0000000000000LL);  // HUGE_VAL
    double bx = 0.0, by = 0.0;
    unsigned long long bid = ID_NONE;
    uint32_t kb = cellStart[(size_t)j * nx + i], ke = cellStart[(size_t)j * nx + i + 1];
    for (uint32_t k = kb; k < ke; ++k) {
        uint32_t e = sortedIdx[k];
        double2 p = pos[e];
        double d = nodeDistance(p.x, p.y, i, j, dx, dr);
        if (d < best) { best = d; bx = p.x; by = p.y; bid = e; }
    }
    long long o = (long long)j * pitch + i;
    phi[o] = best; lpx[o] = bx; lpy[o] = by; lid[o] = __longlong_as_double((long long)bid);
}

// one visit of the closest-particle propagation (:775-792) at cell (i,j)
template <int SX, int SY>
struct OpLsConstruct {
    static constexpr int NIN = 4, NOUT = 4, W = 3;
    static constexpr bool UPROW = true, LOOK = true, INPLACE = true;
    const double* in[4];  // phi, px, py, id
    double* out[4];
    int nx, ny;
    double dx, dr;
    int* sweepCounter;

    __device__ void boundaryState(double* st) const {
        st[0] = 0.0; st[1] = 0.0; st[2] = __longlong_as_double((long long)ID_NONE);
    }
    __device__ __forceinline__ void offer(double cpx, double cpy, double cid, int i, int j, double& phi, double& px,
                                          double& py, double& id, bool& changed) const {
        if ((unsigned long long)__double_as_longlong(cid) == ID_NONE) return;
        double d = nodeDistance(cpx, cpy, i, j, dx, dr);
        if (d < phi) { phi = d; px = cpx; py = cpy; id = cid; changed = true; }
    }
    __device__ bool cell(int i, int j, const double* own, const double* right, const double* up, const double* left,
                         const double* down, double* o, double* st, double& acc) const {
        double phi = own[0], px = own[1], py = own[2], id = own[3];
        bool changed = false;
        if (i < nx && j < ny) {
            // neighbour order of the reference: (i-1,j), (i+1,j), (i,j-1), (i,j+1); the march-previous ones
            // carry this sweep's values (registers / shuffle), the others still hold the previous sweep's
            if (i - 1 >= 0) {
                if (SX > 0) offer(left[0], left[1], left[2], i, j, phi, px, py, id, changed);
                else offer(right[1], right[2], right[3], i, j, phi, px, py, id, changed);
            }
            if (i + 1 < nx) {
                if (SX > 0) offer(right[1], right[2], right[3], i, j, phi, px, py, id, changed);
                else offer(left[0], left[1], left[2], i, j, phi, px, py, id, changed);
            }
            if (j - 1 >= 0) {
                if (SY > 0) offer(down[0], down[1], down[2], i, j, phi, px, py, id, changed);
                else offer(up[1], up[2], up[3], i, j, phi, px, py, id, changed);
            }
            if (j + 1 < ny) {
                if (SY > 0) offer(up[1], up[2], up[3], i, j, phi, px, py, id, changed);
                else offer(down[0], down[1], down[2], i, j, phi, px, py, id, changed);
            }
        }
        o[0] = phi; o[1] = px; o[2] = py; o[3] = id;
        st[0] = px; st[1] = py; st[2] = id;
        return changed;
    }
    __device__ void stripDone(int, double) const {}
    __device__ void allDone(int) const { atomicAdd(sweepCounter, 1); }
};

// :803-842 -- cells on either side of a sign change across a +x / +y edge are "surface"; every other negative
// cell is reset to -inf before the eikonal sweeps.  (The phi assignments at :814-828 re-store the value that
// is already there.)  Gather form: a cell looks at its four edges with the loop bounds of the reference.
__global__ void lsSurfaceKernel(const double* __restrict__ src, double* __restrict__ dst, int nx, int ny, int pitch, DevCtl* ctl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nx || j >= ny) return;
    long long o = (long long)j * pitch + i;
    double c = src[o];
    bool surf = false;
    if (i <= nx - 2 && j <= ny - 2) {
        surf |= __dmul_rn(c, src[o + 1]) < 0;
        surf |= __dmul_rn(c, src[o + pitch]) < 0;
    }
    if (i >= 1 && j <= ny - 2) surf |= __dmul_rn(src[o - 1], c) < 0;
    if (j >= 1 && i <= nx - 2) surf |= __dmul_rn(src[o - pitch], c) < 0;
    dst[o] = (!surf && c < 0) ? __longlong_as_double(0xFFF0000000000000LL) : c;
    // bounding box of the negative cells: the eikonal sweeps only ever change those (:848, 862, 876, 890)
    const bool neg = c < 0;
    if (__any_sync(__activemask(), neg)) {
        if (neg) {
            if (i < ctl->lsBox[0]) atomicMin(&ctl->lsBox[0], i);
            if (i > ctl->lsBox[1]) atomicMax(&ctl->lsBox[1], i);
            if (j < ctl->lsBox[2]) atomicMin(&ctl->lsBox[2], j);
            if (j > ctl->lsBox[3]) atomicMax(&ctl->lsBox[3], j);
        }
    }
}
__global__ void lsBoxResetKernel(DevCtl* ctl) {
    ctl->lsBox[0] = 0x7fffffff; ctl->lsBox[1] = -1; ctl->lsBox[2] = 0x7fffffff; ctl->lsBox[3] = -1;
}

// one directional eikonal sweep (:846-901)
template <int SX, int SY>
struct OpLsRedistance {
    static constexpr int NIN = 1, NOUT = 1, W = 1;
    static constexpr bool UPROW = false, LOOK = false, INPLACE = true;
    const double* in[1];
    double* out[1];
    int nx, ny;
    double dx;
    int* sweepCounter;

    __device__ void boundaryState(double* st) const { st[0] = 0.0; }
    __device__ bool cell(int i, int j, const double* own, const double*, const double*, const double* left,
                         const double* down, double* o, double* st, double& acc) const {
        double phi = own[0];
        bool changed = false;
        const int ilo = SX > 0 ? 1 : 0, ihi = SX > 0 ? nx - 1 : nx - 2;
        const int jlo = SY > 0 ? 1 : 0, jhi = SY > 0 ? ny - 1 : ny - 2;
        if (i >= ilo && i <= ihi && j >= jlo && j <= jhi && !(phi >= 0)) {
            double a = fabs(left[0]), b = fabs(down[0]);
            double phi0 = amlMin(a, b), phi1 = amlMax(a, b);
            double d = __dadd_rn(phi0, dx);
            if (d > phi1) {
                double diff = __dsub_rn(phi1, phi0);
                double arg = __fma_rn(__dadd_rn(dx, dx), dx, -__dmul_rn(diff, diff));
                d = __dmul_rn(0.5, __dadd_rn(__dadd_rn(phi0, phi1), __dsqrt_rn(arg)));
            }
            if (d < -phi) { phi = -d; changed = true; }
        }
        o[0] = phi;
        st[0] = phi;
        return changed;
    }
    __device__ void stripDone(int, double) const {}
    __device__ void allDone(int) const { atomicAdd(sweepCounter, 1); }
};

// :905-919 -- one Jacobi smoothing pass on the interior, everything else copied
__global__ void lsSmoothKernel(const double* __restrict__ src, double* __restrict__ dst, int nx, int ny, int pitch) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nx || j >= ny) return;
    long long o = (long long)j * pitch + i;
    double c = src[o];
    if (i >= 1 && i < nx - 1 && j >= 1 && j < ny - 1) {
        double avg = __dmul_rn(0.25, __dadd_rn(__dadd_rn(__dadd_rn(src[o - 1], src[o + 1]), src[o - pitch]), src[o + pitch]));
        if (avg < c) c = avg;
    }
    dst[o] = c;
}

// :657-664 relabel, :709-722 grid statistics (sampled from the grid the previous step left behind)
__global__ void lsRelabelStatsKernel(const double* __restrict__ phi, uint8_t* __restrict__ cell, GridView g,
                                     double rho, double gx, double gy, int doStats, double* partials,
                                     unsigned int* counters, DevCtl* ctl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    double cnt = 0.0, en = 0.0;
    if (i < g.nx && j < g.ny) {
        long long o = (long long)j * g.pitch + i;
        uint8_t c = cell[o];
        if (c != FSIM_CELL_SOLID) {
            c = phi[o] < 0.0 ? FSIM_CELL_FLUID : FSIM_CELL_EMPTY;
            cell[o] = c;
        }
        if (doStats && c == FSIM_CELL_FLUID) {
            cnt = 1.0;
            double x = i * g.dx, y = j * g.dx;
            double vx = sampleU<false>(g, x, y), vy = sampleV<false>(g, x, y);
            double m = rho * g.dx * g.dx;
            en = 0.5 * m * (vx * vx + vy * vy) - m * (gx * x + gy * y);
        }
    }
    if (!doStats) return;
    __shared__ double bufA[256], bufB[256];
    __shared__ bool isLast;
    const unsigned int tid = threadIdx.y * blockDim.x + threadIdx.x;  // 32x8 block
    bufA[tid] = cnt; bufB[tid] = en;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { bufA[tid] += bufA[tid + o]; bufB[tid] += bufB[tid + o]; }
        __syncthreads();
    }
    const unsigned int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) {
        partials[bid] = bufA[0];
        partials[nblocks + bid] = bufB[0];
        __threadfence();
        isLast = atomicAdd(&counters[1], 1u) == nblocks - 1;
    }
    __syncthreads();
    if (isLast) {
        __threadfence();
        double a = 0.0, b = 0.0;
        for (unsigned int k = tid; k < nblocks; k += 256) { a += __ldcg(&partials[k]); b += __ldcg(&partials[nblocks + k]); }
        bufA[tid] = a; bufB[tid] = b;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) { bufA[tid] += bufA[tid + o]; bufB[tid] += bufB[tid + o]; }
            __syncthreads();
        }
        if (tid == 0) { ctl->fluidCells = bufA[0]; ctl->gridEnergy = bufB[0]; counters[1] = 0; }
    }
}

// ------------------------------------------------------------------------------------------------------
// The same two sweeps as Ops of the in-place SD kernel (sdsweep.cuh).  pc/nc = previous/next column and pr/nr =
// previous/next row in march order; layout column c is grid column nx-1-c on the mirrored layout.
// ------------------------------------------------------------------------------------------------------
template <bool MIRROR, int DIR>
struct OpSdConstruct {
    static constexpr int NA = 4, NW = 4, NN = 3;
    static constexpr bool KEEP_PC = false;
    double* arr[4];  // px, py, id, phi
    int nx, ny;
    double dx, dr;
    int* sweepCounter;
    const int* chg; int* gateOut; int t;  // DevCtl::lsChanged[0], &lsGate[0][t], sweep index
    // exact skipping of quiet sub-chunks (sdsweep.cuh, OpSkip): a visit is the same function of the cell and its four
    // neighbours in every sweep, so after the first sweep a cell can only change near a change of the sweep before (or of this
    // one, upstream).  Plane 0 (the particle's x, positive) carries the hand-off flag.
    static constexpr bool SKIP = true, SKIP_ALLNB = true;
    static constexpr int SKIP_FIRST = 1, SKIP_WINDOW = 1;
    const unsigned char* tileNeg; int* tileStamp; int nblk; int mirror; int noSkip;
    // distance offered by a neighbour's particle, +inf if there is none: computed for all four neighbours up front
    // (branch-free, four independent square roots in flight) -- the dependent part of a visit is four compares
    __device__ __forceinline__ double offered(bool inb, const double (&cand)[3], int i, int j) const {
        const double d = nodeDistance(cand[0], cand[1], i, j, dx, dr);
        const bool ok = inb && (unsigned long long)__double_as_longlong(cand[2]) != ID_NONE;
        return ok ? d : __longlong_as_double(0x7FF0000000000000LL);
    }
    __device__ __forceinline__ void take(double d, const double (&cand)[3], double (&own)[4], bool& changed) const {
        if (d < own[3]) { own[3] = d; own[0] = cand[0]; own[1] = cand[1]; own[2] = cand[2]; changed = true; }
    }
    __device__ bool cell(int c, int j, double (&own)[4], const double (&pc)[3], const double (&nc)[3], const double (&pr)[3],
                         const double (&nr)[3]) const {
        const bool active = !(c < 0 || c >= nx || j >= ny);
        const int i = MIRROR ? nx - 1 - c : c;
        // grid neighbours from march neighbours: x ascends with the march iff MIRROR == (DIR < 0)
        constexpr bool XUP = MIRROR == (DIR < 0);
        const double (&xm)[3] = XUP ? pc : nc;
        const double (&xp)[3] = XUP ? nc : pc;
        const double (&ym)[3] = DIR > 0 ? pr : nr;
        const double (&yp)[3] = DIR > 0 ? nr : pr;
        // Warp-uniform early out, decided exactly without a square root: a candidate cannot win if it is the cell's own
        // particle (same distance) or if its squared distance exceeds ((phi + dr)(1 + 1e-12))^2 (then fl(sqrt(x)) >
        // phi + dr, so fl(fl(sqrt(x)) - dr) >= phi and the reference's `d < phi` is false; phi only decreases during a
        // visit, so the incoming phi's bound is safe for all four offers).  After the first round most visits end here.
        {
            const double t = (own[3] + dr) * 1.000000000001;
            const double thr = t * t * 1.000000000001;  // (+inf while the cell has no particle)
            const unsigned long long myId = (unsigned long long)__double_as_longlong(own[2]);
            auto possible = [&](bool inb, const double (&cand)[3]) {
                const unsigned long long cid = (unsigned long long)__double_as_longlong(cand[2]);
                const double a = __fma_rn(-(double)i, dx, cand[0]), b = __fma_rn(-(double)j, dx, cand[1]);
                return inb && cid != ID_NONE && cid != myId && !(__fma_rn(a, a, __dmul_rn(b, b)) > thr);
            };
            const bool any = possible(active && i - 1 >= 0, xm) | possible(active && i + 1 < nx, xp) |
                             possible(active && j - 1 >= 0, ym) | possible(active && j + 1 < ny, yp);
            if (!__any_sync(0xffffffffu, any)) return false;
        }
        const double d0 = offered(active && i - 1 >= 0, xm, i, j), d1 = offered(active && i + 1 < nx, xp, i, j);
        const double d2 = offered(active && j - 1 >= 0, ym, i, j), d3 = offered(active && j + 1 < ny, yp, i, j);
        bool changed = false;
        // neighbour order of the reference: (i-1,j), (i+1,j), (i,j-1), (i,j+1)  (:777-790); `d < phi` with d = +inf is
        // false, like the reference's skipped offer
        take(d0, xm, own, changed);
        take(d1, xp, own, changed);
        take(d2, ym, own, changed);
        take(d3, yp, own, changed);
        return changed;
    }
    __device__ void allDone(int) const {
        atomicAdd(sweepCounter, 1);
        *gateOut = __ldcg(chg + t);  // a sweep without a change is a fixed point of every later sweep
    }
};

template <bool MIRROR, int DIR>
struct OpSdRedistance {
    static constexpr int NA = 1, NW = 1, NN = 1;
    static constexpr bool KEEP_PC = true;
    double* arr[1];  // phi
    int nx, ny;      // extent of the swept window (the negative cells' bounding box plus one cell)
    int i0, j0;      // its origin in the grid
    int gnx, gny;    // the grid
    double dx;
    int* sweepCounter;
    const int* chg; int* gateOut; int t;  // DevCtl::lsChanged[1], &lsGate[1][t], sweep index
    // exact skipping of quiet sub-chunks (sdsweep.cuh, OpSkip): per (strip, block of 32 window columns) "has a negative
    // cell" and "index of the last sweep that changed a cell"; only negative cells ever change (:848, 862, 876, 890) and a
    // visit reads the march-previous neighbours through fabs() only
    static constexpr bool SKIP = true, SKIP_ALLNB = false;
    static constexpr int SKIP_FIRST = 4, SKIP_WINDOW = 3;  // a direction's sweep is idempotent: only what the three sweeps since changed matters
    const unsigned char* tileNeg; int* tileStamp; int nblk; int mirror; int noSkip;
    __device__ bool cell(int c, int j, double (&own)[1], const double (&pc)[1], const double (&nc)[1], const double (&pr)[1],
                         const double (&nr)[1]) const {
        if (c < 0 || c >= nx || j >= ny) return false;
        const int i = (MIRROR ? nx - 1 - c : c) + i0, jj = j + j0;  // grid coordinates (the loop bounds are the grid's)
        constexpr bool XUP = MIRROR == (DIR < 0);
        const int ilo = XUP ? 1 : 0, ihi = XUP ? gnx - 1 : gnx - 2;
        const int jlo = DIR > 0 ? 1 : 0, jhi = DIR > 0 ? gny - 1 : gny - 2;
        double phi = own[0];
        if (i >= ilo && i <= ihi && jj >= jlo && jj <= jhi && !(phi >= 0)) {
            double a = fabs(pc[0]), b = fabs(pr[0]);  // the march-previous neighbours (:846-901)
            double phi0 = amlMin(a, b), phi1 = amlMax(a, b);
            double d = __dadd_rn(phi0, dx);
            if (d > phi1) {
                double diff = __dsub_rn(phi1, phi0);
                double arg = __fma_rn(__dadd_rn(dx, dx), dx, -__dmul_rn(diff, diff));
                d = __dmul_rn(0.5, __dadd_rn(__dadd_rn(phi0, phi1), __dsqrt_rn(arg)));
            }
            if (d < -phi) { own[0] = -d; return true; }
        }
        return false;
    }
    __device__ void allDone(int) const {
        atomicAdd(sweepCounter, 1);
        int any = 0;  // the same direction runs again in four sweeps: a no-op unless one of the three in between changes a cell
        for (int q = t > 2 ? t - 2 : 0; q <= t; ++q) any |= __ldcg(chg + q);
        *gateOut = any;
    }
};

// which (strip of 32 rows, block of 32 columns) of the eikonal window hold a negative cell
__global__ void __launch_bounds__(256) lsTileNegKernel(const double* __restrict__ phiWin, int nxw, int nyw, int pitch, int nblk,
                                                       unsigned char* __restrict__ tileNeg) {
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), k = blockIdx.y;
    bool neg = false;
    for (int r = threadIdx.x >> 5; r < 32; r += 8) {
        const int j = 32 * k + r;
        if (i < nxw && j < nyw) neg |= phiWin[(long long)j * pitch + i] < 0;
    }
    const int any = __syncthreads_or(neg ? 1 : 0);
    if (threadIdx.x == 0) tileNeg[k * nblk + blockIdx.x] = any ? 1 : 0;
}

constexpr int LS_SUBS = 8;

static int lsClusterSize() {
    static int cl = -1;
    if (cl < 0) {
        cl = 8;
        if (const char* e = getenv("FSIM_SD_CLUSTER")) { int v = atoi(e); if (v == 1 || v == 8) cl = v; }
    }
    return cl;
}

// layouts of the level-set working set: row-major frames, SD, x-mirrored SD
enum { LS_ROW = 0, LS_SD = 1, LS_SDM = 2 };

struct LsArrays {
    int n;
    double* frame[4];  // (0,0) of the packed window
    double* sdArr[4];
    int layout;        // what the host asked for last; the device may have skipped gated switches (DevCtl::lsMirror is the truth)
    sd::Geom g;
    int kind;          // 0 construct, 1 redistance
};

// gate: the switch only happens if the sweep it prepares will run (same flag); the final switch back to rows is never
// gated and takes the mirror flag from the device
static int lsToLayout(Sim* s, LsArrays& A, int want, const int* gate = nullptr) {
    if (A.layout == want) return FSIM_OK;
    const sd::Geom& g = A.g;
    dim3 blk(32, 8), grd(g.nchunks, g.nstrips, A.n);
    sd::PackJob job;
    int* devMirror = &s->ctl->lsMirror[A.kind];
    if (A.layout != LS_ROW) {
        for (int k = 0; k < A.n; ++k) { job.src[k] = A.sdArr[k]; job.dst[k] = A.frame[k]; }
        sd::sdUnpackKernel<<<grd, blk, 0, s->stream>>>(job, g, s->fr.pitch, A.layout == LS_SDM, gate, devMirror);
        LAUNCH_COUNT(s);
    }
    if (want != LS_ROW) {
        for (int k = 0; k < A.n; ++k) { job.src[k] = A.frame[k]; job.dst[k] = A.sdArr[k]; }
        sd::sdPackKernel<<<grd, blk, 0, s->stream>>>(job, g, s->fr.pitch, want == LS_SDM, 0, gate, devMirror);
        LAUNCH_COUNT(s);
    }
    A.layout = want;
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}

// t = index of the sweep among the 16 of its kind; see DevCtl::lsGate for the two gating rules
static const int* lsGateOf(Sim* s, int kind, int t) {
    const bool gated = kind == 0 ? t >= 1 : t >= 4;
    return gated ? &s->ctl->lsGate[kind][t - 1] : nullptr;
}
template <class Op, int DIR>
static int lsLaunchSweep(Sim* s, Op& op, const sd::Geom& g, int kind, int t) {
    const bool gated = kind == 0 ? t >= 1 : t >= 4;
    op.chg = s->ctl->lsChanged[kind]; op.gateOut = &s->ctl->lsGate[kind][t]; op.t = t;
    sd::SweepControl ctl{s->wfTicket, s->wfFinished, s->swHand, s->swPlaneWords,
                         gated ? &s->ctl->lsGate[kind][t - 1] : nullptr, &s->ctl->lsChanged[kind][t]};
    profBegin(s, 5 + kind);  // 5: closest-particle sweep, 6: eikonal sweep
    CUDA_TRY((sd::launchSweep<Op, 1, DIR, LS_SUBS>(op, g, ctl, s->stream, lsClusterSize())));
    profEnd(s);
    LAUNCH_COUNT(s);
    return FSIM_OK;
}

static int lsNoSkip() {
    static int noSkip = -1;
    if (noSkip < 0) { const char* e = getenv("FSIM_LS_NOSKIP"); noSkip = e && atoi(e) ? 1 : 0; }  // A/B knob
    return noSkip;
}

// sweep with x marching in direction SX and y in direction SY (include/FluidSim2D.h:178-203)
template <int SX, int SY>
static int sdConstructSweep(Sim* s, LsArrays& A, int round) {
    constexpr bool MIRROR = SX != SY;
    int rc = lsToLayout(s, A, MIRROR ? LS_SDM : LS_SD, lsGateOf(s, 0, round));
    if (rc) return rc;
    OpSdConstruct<MIRROR, SY> op;
    for (int k = 0; k < 4; ++k) op.arr[k] = A.sdArr[k];
    op.nx = s->nx; op.ny = s->ny; op.dx = s->dx; op.dr = s->dr; op.sweepCounter = &s->ctl->sweepsRun;
    op.tileNeg = s->lsTileNeg[0]; op.tileStamp = s->lsTileStamp[0]; op.nblk = (A.g.nx + 31) / 32; op.mirror = MIRROR ? 1 : 0;
    op.noSkip = lsNoSkip();
    return lsLaunchSweep<OpSdConstruct<MIRROR, SY>, SY>(s, op, A.g, 0, round);
}

template <int SX, int SY>
static int sdRedistanceSweep(Sim* s, LsArrays& A, int round) {
    constexpr bool MIRROR = SX != SY;
    int rc = lsToLayout(s, A, MIRROR ? LS_SDM : LS_SD, lsGateOf(s, 1, round));
    if (rc) return rc;
    OpSdRedistance<MIRROR, SY> op;
    op.arr[0] = A.sdArr[0];
    op.nx = A.g.nx; op.ny = A.g.ny; op.i0 = s->lsWin[0]; op.j0 = s->lsWin[1]; op.gnx = s->nx; op.gny = s->ny;
    op.dx = s->dx; op.sweepCounter = &s->ctl->sweepsRun;
    op.tileNeg = s->lsTileNeg[1]; op.tileStamp = s->lsTileStamp[1]; op.nblk = (A.g.nx + 31) / 32; op.mirror = MIRROR ? 1 : 0;
    op.noSkip = lsNoSkip();
    return lsLaunchSweep<OpSdRedistance<MIRROR, SY>, SY>(s, op, A.g, 1, round);
}

static bool lsLegacy(const Sim* s) {
    static int legacy = -1;
    if (legacy < 0) { const char* e = getenv("FSIM_LS_LEGACY"); legacy = e && atoi(e) ? 1 : 0; }
    return legacy || s->opt.debugSimpleWavefront;
}

template <int SX, int SY>
int constructSweep(Sim* s, int round) {
    OpLsConstruct<SX, SY> op;
    op.in[0] = s->phiTmp; op.in[1] = s->lsPx; op.in[2] = s->lsPy; op.in[3] = s->lsId;
    op.out[0] = s->phiTmp; op.out[1] = s->lsPx; op.out[2] = s->lsPy; op.out[3] = s->lsId;
    op.nx = s->nx; op.ny = s->ny; op.dx = s->dx; op.dr = s->dr;
    op.sweepCounter = &s->ctl->sweepsRun;
    const int* gate = round > 0 ? &s->ctl->lsChanged[0][round - 1] : nullptr;
    return launchWavefront<OpLsConstruct<SX, SY>, SX, SY>(s, op, (s->nx + 31) / 32, (s->ny + 31) / 32, gate, 1,
                                                          &s->ctl->lsChanged[0][round]);
}

template <int SX, int SY>
int redistanceSweep(Sim* s, int round) {
    OpLsRedistance<SX, SY> op;
    op.in[0] = s->phi; op.out[0] = s->phi;
    op.nx = s->nx; op.ny = s->ny; op.dx = s->dx;
    op.sweepCounter = &s->ctl->sweepsRun;
    const int* gate = round > 0 ? &s->ctl->lsChanged[1][round - 1] : nullptr;
    return launchWavefront<OpLsRedistance<SX, SY>, SX, SY>(s, op, (s->nx + 31) / 32, (s->ny + 31) / 32, gate, 1,
                                                           &s->ctl->lsChanged[1][round]);
}

}  // namespace

int stageCreateWaterLevelSet(Sim* s) {
    int rc = sortParticlesByCell(s);
    if (rc) return rc;
    const Frame& f = s->fr;
    CUDA_TRY(cudaMemsetAsync(s->ctl->lsChanged, 0, sizeof(int) * 67, s->stream));  // lsChanged + lsGate + sweepsRun + lsMirror
    dim3 blk(32, 8), grd((s->nx + 31) / 32, (s->ny + 7) / 8);
    lsBinKernel<<<grd, blk, 0, s->stream>>>(s->pos, s->cellStart, s->sortedIdx, s->nx, s->ny, f.pitch, s->dx, s->dr,
                                            s->phiTmp, s->lsPx, s->lsPy, s->lsId);
    LAUNCH_COUNT(s);
    // sweep order of LevelSet::fastSweepIterate (include/FluidSim2D.h:178-203)
    const bool legacy = lsLegacy(s);
    if (legacy) {
        for (int k = 0; k < 4; ++k) {
            if ((rc = constructSweep<+1, +1>(s, k))) return rc;
            if ((rc = constructSweep<-1, +1>(s, k))) return rc;
            if ((rc = constructSweep<+1, -1>(s, k))) return rc;
            if ((rc = constructSweep<-1, -1>(s, k))) return rc;
        }
    } else {
        // in-place sweeps on the strip-diagonal layout; the PCG's SD vectors are free at this point of the step
        LsArrays A{4, {s->lsPx, s->lsPy, s->lsId, s->phiTmp}, {s->sS, s->sT, s->sP, s->sZ}, LS_ROW, s->swg, 0};
        {
            const size_t tiles = (size_t)((A.g.nx + 31) / 32) * A.g.nstrips;
            CUDA_TRY(cudaMemsetAsync(s->lsTileNeg[0], 1, tiles, s->stream));                      // every cell may change
            CUDA_TRY(cudaMemsetAsync(s->lsTileStamp[0], 0x80, tiles * sizeof(int), s->stream));   // "never"
        }
        for (int k = 0; k < 4; ++k) {
            if ((rc = sdConstructSweep<+1, +1>(s, A, 4 * k + 0))) return rc;
            if ((rc = sdConstructSweep<-1, +1>(s, A, 4 * k + 1))) return rc;
            if ((rc = sdConstructSweep<+1, -1>(s, A, 4 * k + 2))) return rc;
            if ((rc = sdConstructSweep<-1, -1>(s, A, 4 * k + 3))) return rc;
        }
        // only phi is needed from here on
        LsArrays P{1, {s->phiTmp}, {s->sZ}, A.layout, s->swg, 0};
        if ((rc = lsToLayout(s, P, LS_ROW))) return rc;
    }
    lsBoxResetKernel<<<1, 1, 0, s->stream>>>(s->ctl);
    lsSurfaceKernel<<<grd, blk, 0, s->stream>>>(s->phiTmp, s->phi, s->nx, s->ny, f.pitch, s->ctl);
    s->launches += 2;
    if (legacy) {
        for (int k = 0; k < 4; ++k) {
            if ((rc = redistanceSweep<+1, +1>(s, k))) return rc;
            if ((rc = redistanceSweep<-1, +1>(s, k))) return rc;
            if ((rc = redistanceSweep<+1, -1>(s, k))) return rc;
            if ((rc = redistanceSweep<-1, -1>(s, k))) return rc;
        }
    } else {
        // The eikonal sweeps only change negative cells and only read their march-previous neighbours, so they run on the
        // negative cells' bounding box grown by one cell (16 bytes read back; everything outside is untouched anyway).
        CUDA_TRY(cudaMemcpyAsync(s->hBox, s->ctl->lsBox, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (s->hBox[1] >= 0) {
            const int i0 = s->hBox[0] > 0 ? s->hBox[0] - 1 : 0, i1 = s->hBox[1] < s->nx - 1 ? s->hBox[1] + 1 : s->nx - 1;
            const int j0 = s->hBox[2] > 0 ? s->hBox[2] - 1 : 0, j1 = s->hBox[3] < s->ny - 1 ? s->hBox[3] + 1 : s->ny - 1;
            s->lsWin[0] = i0; s->lsWin[1] = j0;
            LsArrays P{1, {s->phi + (long long)j0 * f.pitch + i0}, {s->sZ}, LS_ROW, sd::makeGeom(i1 - i0 + 1, j1 - j0 + 1, 1), 1};
            {
                const int nblk = (P.g.nx + 31) / 32;
                lsTileNegKernel<<<dim3(nblk, P.g.nstrips), 256, 0, s->stream>>>(P.frame[0], P.g.nx, P.g.ny, f.pitch, nblk, s->lsTileNeg[1]);
                LAUNCH_COUNT(s);
                CUDA_TRY(cudaMemsetAsync(s->lsTileStamp[1], 0x80, (size_t)nblk * P.g.nstrips * sizeof(int), s->stream));  // "never"
            }
            for (int k = 0; k < 4; ++k) {
                if ((rc = sdRedistanceSweep<+1, +1>(s, P, 4 * k + 0))) return rc;
                if ((rc = sdRedistanceSweep<-1, +1>(s, P, 4 * k + 1))) return rc;
                if ((rc = sdRedistanceSweep<+1, -1>(s, P, 4 * k + 2))) return rc;
                if ((rc = sdRedistanceSweep<-1, -1>(s, P, 4 * k + 3))) return rc;
            }
            if ((rc = lsToLayout(s, P, LS_ROW))) return rc;
        }
    }
    lsSmoothKernel<<<grd, blk, 0, s->stream>>>(s->phi, s->phiTmp, s->nx, s->ny, f.pitch);
    lsSmoothKernel<<<grd, blk, 0, s->stream>>>(s->phiTmp, s->phi, s->nx, s->ny, f.pitch);
    if ((rc = joinUpload(s))) return rc;  // first reader of u, v in a frame (the statistics)
    if ((rc = joinFarFill(s))) return rc;  // ... and the previous frame's far extrapolation layers, if it left them running
    GridView g{s->u, s->v, s->nx, s->ny, f.pitch, s->dx};
    lsRelabelStatsKernel<<<grd, blk, 0, s->stream>>>(s->phi, s->cell, g, s->rho, s->gx, s->gy, s->opt.computeStats,
                                                     s->partials, s->counters, s->ctl);
    s->launches += 3;
    CUDA_TRY(cudaGetLastError());
    if (s->opt.computeStats) {
        if ((rc = particleEnergy(s))) return rc;
    }
    s->statsValid = true;
    return FSIM_OK;
}
