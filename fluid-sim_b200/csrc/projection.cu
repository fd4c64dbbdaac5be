// applyProjection (reference src/FluidSim2D.cpp:252-467) and updateVelocity (:469-550).
//
// Matrix assembly with the ghost-pressure free-surface terms (:260-303), the negative-divergence right-hand
// side (:334-362), the MIC(0) factor (:364-388, tau = 0.999, sigma = 0.25), and the PCG loop (:423-466) with
// its stop rule |r|_inf <= tol*|rhs|_inf and iteration cap.  The preconditioner is the reference's own:
// both triangular solves and the factor run on the exact wavefront scheduler (wavefront.cuh).
//
// Algebraic restatement of the solves (same operator, fewer bytes): with D = precon^2, Ux = Ax*D, Uy = Ay*D,
//   forward:  t(i,j) = r(i,j) - Ux(i-1,j) t(i-1,j) - Uy(i,j-1) t(i,j-1)        (t = q / precon)
//   backward: z(i,j) = D(i,j) t(i,j) - Ux(i,j) z(i+1,j) - Uy(i,j) z(i,j+1)
// Coefficients vanish on non-fluid cells, so no label reads and no range tests are needed; one dependent
// FMA per cell sits on the critical path.  Every scalar of the loop (sigma, alpha, beta, norms, iteration
// count, convergence flag) lives in DevCtl and is produced by the last block of the kernel that owns the
// reduction, in a fixed order: the solve is deterministic and never waits for the host inside an iteration.
#include <stdlib.h>
#include <string.h>

#include "distpeer.cuh"
#include "sdpack.cuh"
#include "sdsweep.cuh"
#include "sdwave.cuh"
#include "wf_launch.cuh"

namespace {

__device__ __forceinline__ uint8_t cellAt(const uint8_t* cell, int pitch, int nx, int ny, int i, int j) {
    if (i < 0 || i >= nx || j < 0 || j >= ny) return FSIM_CELL_SOLID;
    return cell[(long long)j * pitch + i];
}

__global__ void assembleKernel(const uint8_t* __restrict__ cell, const double* __restrict__ phi,
                               const double* __restrict__ u, const double* __restrict__ v, int nx, int ny, int pitch,
                               double scaleA, double invDx, double* __restrict__ Adiag, double* __restrict__ Ax,
                               double* __restrict__ Ay, double* __restrict__ rhs, double* __restrict__ fmask,
                               double* __restrict__ r, double* __restrict__ p, double* partials, unsigned int* counter,
                               DevCtl* ctl) {
    __shared__ double red[32];
    __shared__ bool isLast;
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    double ab = 0.0;
    bool isFluid = false;
    if (i < nx && j < ny) {
        long long o = (long long)j * pitch + i;
        double d = 0.0, ax = 0.0, ay = 0.0, b = 0.0, fm = 0.0;
        if (cell[o] == FSIM_CELL_FLUID) {
            fm = 1.0;
            uint8_t cl = cellAt(cell, pitch, nx, ny, i - 1, j), cr = cellAt(cell, pitch, nx, ny, i + 1, j);
            uint8_t cd = cellAt(cell, pitch, nx, ny, i, j - 1), cu = cellAt(cell, pitch, nx, ny, i, j + 1);
            double phi0 = phi[o];
            if (cl == FSIM_CELL_FLUID) d += scaleA;
            else if (cl == FSIM_CELL_EMPTY) d -= scaleA * amlMax(phi[o - 1] / phi0, -1e3);
            if (cr == FSIM_CELL_FLUID) { d += scaleA; ax = -scaleA; }
            else if (cr == FSIM_CELL_EMPTY) d += scaleA * (1 - amlMax(phi[o + 1] / phi0, -1e3));
            if (cd == FSIM_CELL_FLUID) d += scaleA;
            else if (cd == FSIM_CELL_EMPTY) d -= scaleA * amlMax(phi[o - pitch] / phi0, -1e3);
            if (cu == FSIM_CELL_FLUID) { d += scaleA; ay = -scaleA; }
            else if (cu == FSIM_CELL_EMPTY) d += scaleA * (1 - amlMax(phi[o + pitch] / phi0, -1e3));
            double u0 = u[o], u1 = u[o + 1], v0 = v[o], v1 = v[o + pitch];
            b = -invDx * (u1 - u0 + v1 - v0);
            if (cl == FSIM_CELL_SOLID) b -= invDx * (u0 - 0);
            if (cr == FSIM_CELL_SOLID) b += invDx * (u1 - 0);
            if (cd == FSIM_CELL_SOLID) b -= invDx * (v0 - 0);
            if (cu == FSIM_CELL_SOLID) b += invDx * (v1 - 0);
        }
        Adiag[o] = d; Ax[o] = ax; Ay[o] = ay; rhs[o] = b; fmask[o] = fm; r[o] = b; p[o] = 0.0;
        ab = fabs(b);
        isFluid = fm != 0.0;
    }
    // bounding box of the FLUID cells: the solve only has to cover it (everything outside is exactly zero)
    if (__any_sync(0xffffffffu, isFluid)) {
        const unsigned int lo = __reduce_min_sync(0xffffffffu, isFluid ? (unsigned)i : 0x7fffffffu);
        const unsigned int hi = __reduce_max_sync(0xffffffffu, isFluid ? (unsigned)i : 0u);
        if ((threadIdx.x & 31) == 0) {  // (a warp is one row of the block: j is uniform)
            if ((int)lo < ctl->bbox[0]) atomicMin(&ctl->bbox[0], (int)lo);
            if ((int)hi > ctl->bbox[1]) atomicMax(&ctl->bbox[1], (int)hi);
            if (j < ctl->bbox[2]) atomicMin(&ctl->bbox[2], j);
            if (j > ctl->bbox[3]) atomicMax(&ctl->bbox[3], j);
        }
    }
    // |rhs|_inf (the reference recomputes it every iteration, :447; it never changes)
    ab = warpMax(ab);
    unsigned int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if ((tid & 31) == 0) red[tid >> 5] = ab;
    __syncthreads();
    unsigned int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) {
        double m = 0.0;
        for (unsigned int k = 0; k < (blockDim.x * blockDim.y) / 32; ++k) m = fmax(m, red[k]);
        partials[bid] = m;
        __threadfence();
        isLast = atomicAdd(counter, 1u) == nblocks - 1;
    }
    __syncthreads();
    if (isLast) {
        __threadfence();
        double m = 0.0;
        for (unsigned int k = tid; k < nblocks; k += blockDim.x * blockDim.y) m = fmax(m, __ldcg(&partials[k]));
        m = warpMax(m);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = m;
        __syncthreads();
        if (tid == 0) {
            for (unsigned int k = 0; k < (blockDim.x * blockDim.y) / 32; ++k) m = fmax(m, red[k]);
            ctl->rhsNorm = m;
            ctl->iter = 0;
            ctl->hitMax = 0;
            ctl->rnorm = m;
            ctl->pcgDone = (m <= 1e-12) ? 1 : 0;  // :448 -- nothing to solve, p stays 0
            *counter = 0;
        }
    }
}

// MIC(0) factor (:368-387); state handed to the march-next cells: (precon, Ax, Ay) of this cell
struct OpFactor {
    static constexpr int NIN = 4, NOUT = 1, W = 3;
    static constexpr bool UPROW = false, LOOK = false, INPLACE = false;
    const double* in[4];  // Adiag, Ax, Ay, fluid mask
    double* out[1];       // precon
    int nx, ny;
    int jOff;             // global row of local row 0 (a y-slab factors its own rows only: block-MIC(0))
    __device__ void boundaryState(double* st) const { st[0] = 0.0; st[1] = 0.0; st[2] = 0.0; }
    __device__ bool cell(int i, int j, const double* own, const double*, const double*, const double* left,
                         const double* down, double* o, double* st, double& acc) const {
        const double tau = 0.999, sigma = 0.25;
        double pc = 0.0;
        if (own[3] != 0.0 && i >= 1 && j + jOff >= 1 && i < nx && j + jOff < ny) {
            double ad = own[0];
            double pl = left[0], axl = left[1], ayl = left[2];
            double pd = down[0], axd = down[1], ayd = down[2];
            double e = ad - (axl * pl) * (axl * pl) - (ayd * pd) * (ayd * pd) -
                       tau * (axl * ayl * pl * pl + ayd * axd * pd * pd);
            if (e < sigma * ad) e = ad;
            pc = 1.0 / sqrt(e);
        }
        o[0] = pc;
        st[0] = pc; st[1] = own[1]; st[2] = own[2];
        return false;
    }
    __device__ void stripDone(int, double) const {}
    __device__ void allDone(int) const {}
};

// The same factor loop as an Op of the in-place SD sweep kernel (sdsweep.cuh): tile arrays precon (written), Ax, Ay
// (seen by the march-next neighbours), Adiag, fluid mask.
struct OpSdFactor {
    static constexpr int NA = 5, NW = 1, NN = 3;
    static constexpr bool KEEP_PC = true;
    double* arr[5];
    int nx, ny;
    int jOff;  // global row of local row 0
    __device__ bool cell(int c, int j, double (&own)[5], const double (&left)[3], const double (&)[3], const double (&down)[3],
                         const double (&)[3]) const {
        if (c < 0 || c >= nx || j + jOff >= ny) return false;
        const double tau = 0.999, sigma = 0.25;
        double pc = 0.0;
        if (own[4] != 0.0 && c >= 1 && j + jOff >= 1) {
            double ad = own[3];
            double pl = left[0], axl = left[1], ayl = left[2];
            double pd = down[0], axd = down[1], ayd = down[2];
            double e = ad - (axl * pl) * (axl * pl) - (ayd * pd) * (ayd * pd) -
                       tau * (axl * ayl * pl * pl + ayd * axd * pd * pd);
            if (e < sigma * ad) e = ad;
            pc = 1.0 / sqrt(e);
        }
        own[0] = pc;
        return true;
    }
    __device__ void allDone(int) const {}
};

// first / last storage chunk of every strip that holds a FLUID cell (sd::Control::range): cell (c, lane t) of a strip
// sits at storage step c + sigma*t.  One block per strip; rows [32k, 32k+32) of the frame `fm` (already offset to the
// solve's first row).
__global__ void __launch_bounds__(256) stripRangeKernel(const double* __restrict__ fm, int pitch, int nxEff, int nrows, int sigma,
                                                        int rpl, int* __restrict__ range, DevCtl* ctl) {
    __shared__ int sLo[8], sHi[8];
    const int k = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int rows = 32 * rpl;
    int lo = 0x7fffffff, hi = -1;
    for (int q = w; q < rows; q += 8) {
        const int j = rows * k + q, t = q / rpl;  // lane t of the solve owns this row
        if (j >= nrows) break;
        const double* row = fm + (long long)j * pitch;
        for (int c = lane; c < nxEff; c += 32)
            if (row[c] != 0.0) { lo = min(lo, c + sigma * t); hi = max(hi, c + sigma * t); }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) { sLo[w] = lo; sHi[w] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) { lo = min(lo, sLo[i]); hi = max(hi, sHi[i]); }
        if (hi < 0) { range[2 * k] = 1; range[2 * k + 1] = 0; }
        else {
            range[2 * k] = lo / sd::CH; range[2 * k + 1] = hi / sd::CH;
            atomicAdd(&ctl->marchedSlots, (unsigned long long)(hi / sd::CH - lo / sd::CH + 1) * sd::CH * 32 * rpl);
        }
    }
}

// coefficients of the two solves from the factor
__global__ void deriveKernel(const double* __restrict__ pc, const double* __restrict__ Ax, const double* __restrict__ Ay,
                             int ncols, int nrows, int pitch, double* __restrict__ D, double* __restrict__ Ux,
                             double* __restrict__ Uy, double* __restrict__ Lx, double* __restrict__ Ly, int cutBelowRow0) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= ncols || j >= nrows) return;
    long long o = (long long)j * pitch + i;
    double q = pc[o], ql = pc[o - 1], qd = (cutBelowRow0 && j == 0) ? 0.0 : pc[o - pitch];
    D[o] = q * q;
    Ux[o] = Ax[o] * (q * q);
    Uy[o] = Ay[o] * (q * q);
    Lx[o] = Ax[o - 1] * (ql * ql);
    Ly[o] = Ay[o - pitch] * (qd * qd);
}

// ------------------------------------------------------------------------------------------------------
// PCG on the strip-diagonal layout (sdwave.cuh).  All vectors (p, r, z, s, t) and coefficients live in SD
// layout for the whole solve; rhs is packed once, p unpacked once.
// ------------------------------------------------------------------------------------------------------
// forward solve t = L^-1 r in the scaling t = q / precon (sd::solveKernel).  Its post warp writes w = D t (what
// the backward solve consumes) and accumulates sum(t * w) = q.q = z.r, the reference's sigma (:428, :457), so the
// backward solve needs neither r nor a reduction.
struct OpForward {
    static constexpr int NIN = 4, KIND = 1;
    __device__ double postScalar() const { return 0.0; }
    const double* in[4];  // r, Lx, Ly, D
    double* out;          // w = D t
    double* partials;
    DevCtl* ctl;
    int phase;  // 0: first application (sigma = z.r, :428); 1: inside the loop (:457-462)
    int dist;   // y-slab mode: the strip partials are this rank's share; combined over the ranks through peer memory
    PeerView pv;
    __device__ void stripDone(int strip, double acc) const { partials[strip] = acc; }
    __device__ void allDone(int nstrips) const {
        double sum = 0.0;
        for (int k = 0; k < nstrips; ++k) sum += __ldcg(&partials[k]);
        if (dist) {
            double unused = 0.0;
            if (!peerCombine(pv, 1, pv.stampBase | (phase ? (unsigned)(ctl->iter + 1) : 0u), sum, unused)) {
                ctl->distError = 1; ctl->pcgDone = 1;  // a peer never answered: stop rather than hang
                return;
            }
        }
        if (phase == 0) {
            ctl->sigma = sum;
        } else {
            ctl->beta = sum / ctl->sigma;
            ctl->sigma = sum;
            int it = ctl->iter + 1;
            ctl->iter = it;
            if (it >= ctl->maxIters) { ctl->pcgDone = 1; ctl->hitMax = 1; }
        }
    }
};

// backward solve z = w - Ux z(i+1,j) - Uy z(i,j+1), fused with the direction update s = z + beta s (:459): the post
// warp reads the old s from the tile ring (4th array) and writes the new one; z itself is never stored.  first = 1:
// s = z (:427).
struct OpBackward {
    static constexpr int NIN = 4, KIND = 2;
    const double* in[4];  // w, Ux, Uy, s
    double* out;          // s
    const DevCtl* ctl;
    int first;
    __device__ double postScalar() const { return first ? 0.0 : ctl->beta; }
    __device__ void stripDone(int, double) const {}
    __device__ void allDone(int) const {}
};

// Optional (fsim_options.reserved[FSIM_OPT_FUSED_AXPY] = 1; same bits as axpyKernel, tests/test_gpu_parity.py): the PCG's two
// axpys ride on the solves.  Measured on B200 at 4096^2 it is SLOWER than the separate kernel (forward solve 0.145 -> 0.26 ms
// against 0.042 ms saved): the fifth array leaves the TMA ring one stage in flight and the pre warp's pass over each chunk
// lands on the solver's critical path, so the default keeps axpyKernel.
//   * forward solve: its pre warp applies r -= alpha z (:452) to every chunk as it lands in shared memory, writes the new r
//     back and takes |r|_inf; the last strip to finish applies the stop rule (:453) BEFORE beta / sigma / iter (:457-462).
//     phase 0 (first application, :426): alpha = 0, nothing stored, no stop rule.
//   * backward solve: its post warp has the old direction in the tile for s = z + beta s (:459) and applies
//     p += alpha s (:451) to it on the way (the 5th array of the ring is p).
// When the loop ends inside the forward solve (converged or cap) the backward solve of that iteration is gated off and
// p += alpha s is still owed: DevCtl::pendingP, paid by pcgFinishKernel after the loop.  Bytes per iteration:
// applyA 41 + forward 56 (R r, z, Lx, Ly, D; W w, r) + backward 56 (R w, Ux, Uy, s, p; W s, p) = 153 instead of 203.
struct OpForwardF {
    static constexpr int NIN = 5, KIND = 1;
    static constexpr bool PRE_AXPY = true;
    const double* in[5];  // r, Lx, Ly, D, z
    double* out;          // w = D t
    double* rOut;         // r (same array as in[0])
    double* partials;     // [0, 2048): strip sums of t*w; [2048, 4096): strip maxima of |r|
    DevCtl* ctl;
    int phase;
    int dist;     // y-slab mode: the strip partials are this rank's share; combined over the ranks through peer memory
    PeerView pv;
    __device__ double postScalar() const { return 0.0; }
    __device__ double preAlpha() const { return phase ? ctl->alpha : 0.0; }
    __device__ bool preStore() const { return phase != 0; }
    __device__ void stripDone(int strip, double acc) const { partials[strip] = acc; }
    __device__ void stripMax(int strip, double m) const { partials[2048 + strip] = m; }
    __device__ void allDone(int nstrips) const {
        double sum = 0.0, rn = 0.0;
        for (int k = 0; k < nstrips; ++k) { sum += __ldcg(&partials[k]); rn = fmax(rn, __ldcg(&partials[2048 + k])); }
        if (dist && !peerCombine(pv, 1, pv.stampBase | (phase ? (unsigned)(ctl->iter + 1) : 0u), sum, rn)) {
            ctl->distError = 1; ctl->pcgDone = 1;  // a peer never answered: stop rather than hang
            return;
        }
        if (phase == 0) { ctl->sigma = sum; return; }
        ctl->rnorm = rn;
        if (rn <= ctl->tol * ctl->rhsNorm) { ctl->pcgDone = 1; ctl->pendingP = 1; return; }  // :453 (iter is not incremented)
        ctl->beta = sum / ctl->sigma;
        ctl->sigma = sum;
        const int it = ctl->iter + 1;
        ctl->iter = it;
        if (it >= ctl->maxIters) { ctl->pcgDone = 1; ctl->hitMax = 1; ctl->pendingP = 1; }
    }
};
struct OpBackwardF {
    static constexpr int NIN = 5, KIND = 3;
    const double* in[5];  // w, Ux, Uy, s, p
    double* out;          // s
    double* out2;         // p
    const DevCtl* ctl;
    int first;
    __device__ double postScalar() const { return first ? 0.0 : ctl->beta; }
    __device__ double postAlpha() const { return first ? 0.0 : ctl->alpha; }
    __device__ void stripDone(int, double) const {}
    __device__ void allDone(int) const {}
};

// y-slab variant: additionally copies the first / last own row of the new s into the neighbours' ghost rows and stamps them
struct OpBackwardFD : OpBackwardF {
    static constexpr bool HALO = true;
    double *pushLo, *pushHi;             // rank-1's upper ghost row / rank+1's lower ghost row (peer memory), or null
    unsigned int *stampLo, *stampHi;     // their haloSeq words
    unsigned int stampBase;
    __device__ void allDone(int) const {
        // (every CTA fenced at system scope before it counted itself finished: the rows are complete)
        const unsigned int stamp = stampBase | (first ? 1u : (unsigned)(ctl->iter + 1));
        if (stampLo) stReleaseSysU32(stampLo, stamp);
        if (stampHi) stReleaseSysU32(stampHi, stamp);
    }
};

// p += alpha s once more when the loop ended inside a forward solve (see above); chunk ranges as in axpyKernel
__global__ void __launch_bounds__(256) pcgFinishKernel(double* __restrict__ p, const double* __restrict__ s, int nchunks, int nstrips,
                                                       int rpl, const int* __restrict__ range, const DevCtl* ctl) {
    if (!ctl->pendingP) return;
    const double alpha = ctl->alpha;
    const int nItems = nchunks * nstrips;
    for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
        const int strip = item / nchunks, cn = item - strip * nchunks;
        if (range && (cn < range[2 * strip] || cn > range[2 * strip + 1])) continue;
        const size_t base = (size_t)item * 1024 * rpl;
        for (int h = 0; h < 2 * rpl; ++h) {
            const size_t k = base + h * 512 + threadIdx.x * 2;
            double2 pv = *reinterpret_cast<double2*>(p + k);
            const double2 sv = *reinterpret_cast<const double2*>(s + k);
            pv.x = __fma_rn(alpha, sv.x, pv.x); pv.y = __fma_rn(alpha, sv.y, pv.y);
            *reinterpret_cast<double2*>(p + k) = pv;
        }
    }
}

// y-slab variant of the plain backward solve: the same halo rows and stamps as OpBackwardFD
struct OpBackwardD : OpBackward {
    static constexpr bool HALO = true;
    double *pushLo, *pushHi;
    unsigned int *stampLo, *stampHi;
    unsigned int stampBase;
    __device__ void allDone(int) const {
        const unsigned int stamp = stampBase | (first ? 1u : (unsigned)(ctl->iter + 1));
        if (stampLo) stReleaseSysU32(stampLo, stamp);
        if (stampHi) stReleaseSysU32(stampHi, stamp);
    }
};

// z = A s in SD layout (coefficients are zero outside the fluid), fused with z.s (:433-444, :450).
// One thread per slot; the stencil neighbours are at [s-1][t], [s+1][t], [s-SIGMA][t-1], [s+SIGMA][t+1]
// (L1 hits), the first and last lane cross into the neighbouring strip.
// Persistent blocks walk (strip, chunk) tiles of 32 steps x 32 lanes, so the grid reduction has a few hundred
// partials and the +-SIGMA step halo is re-read from L1, not L2.
constexpr int AA_BLOCKS = 148 * 6;
// y-slab mode (dist = 1): the rows just outside the slab are not read from the halo strips of S but from the ghost rows the
// neighbour ranks' backward solves filled over NVLink (distpeer.cuh); the blocks wait for the two stamps first, and the
// thread that finishes z.s combines it over the ranks.
struct ApplyADist {
    int dist;
    const double *ghostLo, *ghostHi;                  // local ghost rows (row below the slab / above it), or null at the ends
    const unsigned int *haloSeqLo, *haloSeqHi;        // their stamps
    PeerView pv;
};
__global__ void __launch_bounds__(256) applyASdKernel(const double* __restrict__ Adiag, const double* __restrict__ Ax,
                                                      const double* __restrict__ Ay, const double* __restrict__ S,
                                                      double* __restrict__ Z, sd::Geom g, int kLo, int kHi,
                                                      const int* __restrict__ range, double* partials,
                                                      unsigned int* counter, DevCtl* ctl, ApplyADist dd) {
    if (ctl->pcgDone) return;
    __shared__ double red[32];
    if (dd.dist) {
        if (threadIdx.x == 0) {
            const unsigned int stamp = dd.pv.stampBase | (unsigned)(ctl->iter + 1);
            bool ok = true;
            if (dd.haloSeqLo) ok = peerWait(dd.haloSeqLo, stamp);
            if (dd.haloSeqHi) ok = peerWait(dd.haloSeqHi, stamp) && ok;
            if (!ok) ctl->distError = 1;
        }
        __syncthreads();
    }
    const int t = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int sg = g.sigma, R = g.rpl, nItems = (kHi - kLo) * g.nchunks;  // strips [kLo, kHi): a y-slab skips its halo strips
    const size_t line = (size_t)32 * R;  // slots per step
    double acc = 0.0;
    for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
        const int k = kLo + item / g.nchunks, cn = item % g.nchunks;
        // chunks without fluid: A is zero there and z stays zero (range of strip kLo + n is range[2n], range[2n+1])
        if (range && (cn < range[2 * (k - kLo)] || cn > range[2 * (k - kLo) + 1])) continue;
        if (R == 2) {
            // two rows per lane: a lane's rows are adjacent in memory, so every operand of the pair is one 16-byte load
            // and all loads of a step are issued before the arithmetic (same expression order as the scalar path below)
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) {
                const int s = cn * 32 + r4 * 8 + w;
                const int c = s - sg * t;
                if (c < 0 || c >= g.nx) continue;
                const int j0 = 64 * k + 2 * t;
                if (j0 >= g.ny) continue;
                const size_t idx = (((size_t)k * g.Sp + s) * 32 + t) * 2;
                const double2 sc = *reinterpret_cast<const double2*>(S + idx);
                const double2 ad = *reinterpret_cast<const double2*>(Adiag + idx);
                const double2 ax = *reinterpret_cast<const double2*>(Ax + idx);
                const double2 ay = *reinterpret_cast<const double2*>(Ay + idx);
                double2 sl = make_double2(0.0, 0.0), axl = sl, sr = sl;
                if (c > 0) { sl = *reinterpret_cast<const double2*>(S + idx - line); axl = *reinterpret_cast<const double2*>(Ax + idx - line); }
                if (c < g.nx - 1) sr = *reinterpret_cast<const double2*>(S + idx + line);
                double sdn = 0.0, ayd = 0.0, su = 0.0;
                if (j0 > 0) {  // row j0-1: lane t-1's second row (sg steps back) or the last row of the strip below
                    const size_t di = t > 0 ? idx - (line * sg + 1) : ((((size_t)(k - 1) * g.Sp + c + 31 * sg) * 32 + 31) * 2 + 1);
                    sdn = S[di]; ayd = Ay[di];
                    if (dd.ghostLo && k == kLo && t == 0) sdn = __ldcg(dd.ghostLo + c);
                }
                if (j0 + 1 < g.ny - 1) {  // row j0+2: lane t+1's first row (sg steps on) or the first row of the strip above
                    const size_t ui = t < 31 ? idx + (line * sg + 2) : (((size_t)(k + 1) * g.Sp + c) * 32) * 2;
                    su = S[ui];
                    if (dd.ghostHi && k == kHi - 1 && t == 31) su = __ldcg(dd.ghostHi + c);
                }
                const double z0 = ad.x * sc.x + axl.x * sl.x + ax.x * sr.x + ayd * sdn + ay.x * sc.y;
                const double z1 = ad.y * sc.y + axl.y * sl.y + ax.y * sr.y + ay.x * sc.x + ay.y * su;
                *reinterpret_cast<double2*>(Z + idx) = make_double2(z0, z1);
                acc = __fma_rn(z0, sc.x, acc);
                acc = __fma_rn(z1, sc.y, acc);
            }
            continue;
        }
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
            const int s = cn * 32 + r4 * 8 + w;
            const int c = s - sg * t;
            if (c < 0 || c >= g.nx) continue;
            for (int rr = 0; rr < R; ++rr) {
                const int j = 32 * R * k + R * t + rr;
                if (j >= g.ny) break;
                const size_t idx = (((size_t)k * g.Sp + s) * 32 + t) * R + rr;
                double sc = S[idx];
                double sl = 0.0, axl = 0.0, sr = 0.0, sdn = 0.0, ayd = 0.0, su = 0.0;
                if (c > 0) { sl = S[idx - line]; axl = Ax[idx - line]; }
                if (c < g.nx - 1) sr = S[idx + line];
                if (j > 0) {  // row j-1: this lane's previous row, lane t-1's last row (sg steps back), or the strip below
                    const size_t di = rr > 0 ? idx - 1
                                             : (t > 0 ? idx - (line * sg + 1) : ((((size_t)(k - 1) * g.Sp + c + 31 * sg) * 32 + 31) * R + R - 1));
                    sdn = S[di]; ayd = Ay[di];
                    if (dd.ghostLo && k == kLo && t == 0 && rr == 0) sdn = __ldcg(dd.ghostLo + c);
                }
                if (j < g.ny - 1) {
                    const size_t ui = rr < R - 1 ? idx + 1 : (t < 31 ? idx + (line * sg + 1) : (((size_t)(k + 1) * g.Sp + c) * 32) * R);
                    su = S[ui];
                    if (dd.ghostHi && k == kHi - 1 && t == 31 && rr == R - 1) su = __ldcg(dd.ghostHi + c);
                }
                double zz = Adiag[idx] * sc + axl * sl + Ax[idx] * sr + ayd * sdn + Ay[idx] * su;
                Z[idx] = zz;
                acc = __fma_rn(zz, sc, acc);
            }
        }
    }
    acc = blockReduce<false>(acc, red);
    gridReduceFinish<false>(acc, partials, counter, red, [&](double zs) {
        if (dd.dist) {
            double unused = 0.0;
            if (!peerCombine(dd.pv, 0, dd.pv.stampBase | (unsigned)(ctl->iter + 1), zs, unused)) { ctl->distError = 1; ctl->pcgDone = 1; }
        }
        ctl->zs = zs;
        ctl->alpha = ctl->sigma / zs;  // :450
        ctl->alphaIter = ctl->iter + 1;  // this iteration's alpha is valid (axpyPKernel)
    });
}

// p += alpha s, r -= alpha z, |r|_inf and the stop rule (:451-453), chunk by chunk (1024 * rpl contiguous slots); chunks
// without fluid are exact zeros in all four vectors and are skipped
__global__ void __launch_bounds__(256) axpyKernel(double* __restrict__ p, double* __restrict__ r,
                                                  const double* __restrict__ s, const double* __restrict__ z, int nchunks,
                                                  int nstrips, int rpl, const int* __restrict__ range, double* partials,
                                                  unsigned int* counter, DevCtl* ctl, int dist, PeerView pv, int withP) {
    if (ctl->pcgDone) return;
    __shared__ double red[32];
    const double alpha = ctl->alpha;
    double m = 0.0;
    const int nItems = nchunks * nstrips;
    for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
        const int strip = item / nchunks, cn = item - strip * nchunks;
        if (range && (cn < range[2 * strip] || cn > range[2 * strip + 1])) continue;
        const size_t base = (size_t)item * 1024 * rpl;
        for (int h = 0; h < 2 * rpl; ++h) {
            const size_t k = base + h * 512 + threadIdx.x * 2;
            double2 rv = *reinterpret_cast<double2*>(r + k);
            const double2 zv = *reinterpret_cast<const double2*>(z + k);
            if (withP) {  // (else: axpyPKernel applies p += alpha s beside the forward solve)
                double2 pv2 = *reinterpret_cast<double2*>(p + k);
                const double2 sv = *reinterpret_cast<const double2*>(s + k);
                pv2.x = __fma_rn(alpha, sv.x, pv2.x); pv2.y = __fma_rn(alpha, sv.y, pv2.y);
                *reinterpret_cast<double2*>(p + k) = pv2;
            }
            rv.x = __fma_rn(-alpha, zv.x, rv.x); rv.y = __fma_rn(-alpha, zv.y, rv.y);
            *reinterpret_cast<double2*>(r + k) = rv;
            m = fmax(m, fmax(fabs(rv.x), fabs(rv.y)));
        }
    }
    m = blockReduce<true>(m, red);
    gridReduceFinish<true>(m, partials, counter, red, [&](double rn) {
        if (dist) {
            double unused = 0.0;
            if (!peerCombine(pv, 2, pv.stampBase | (unsigned)(ctl->iter + 1), unused, rn)) { ctl->distError = 1; ctl->pcgDone = 1; }
        }
        ctl->rnorm = rn;
        if (rn <= ctl->tol * ctl->rhsNorm) ctl->pcgDone = 1;  // :453 (iter is not incremented)
    });
}

// p += alpha s (:451) of iteration `launchIter` as a kernel of its own, on a side stream beside the forward solve: p is not
// read again before the projection ends, so only r -= alpha z has to sit between applyA and the solve.  It runs iff applyA
// of that iteration ran (DevCtl::alphaIter, stamped by the thread that finished z.s) -- also when the stop rule or the cap
// ends the loop in this very iteration, as in the reference, where p is updated before the test.
__global__ void __launch_bounds__(256) axpyPKernel(double* __restrict__ p, const double* __restrict__ s, int nchunks, int nstrips,
                                                   int rpl, const int* __restrict__ range, const DevCtl* ctl, int launchIter) {
    if (ctl->alphaIter != launchIter) return;
    const double alpha = ctl->alpha;
    const int nItems = nchunks * nstrips;
    for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
        const int strip = item / nchunks, cn = item - strip * nchunks;
        if (range && (cn < range[2 * strip] || cn > range[2 * strip + 1])) continue;
        const size_t base = (size_t)item * 1024 * rpl;
        for (int h = 0; h < 2 * rpl; ++h) {
            const size_t k = base + h * 512 + threadIdx.x * 2;
            double2 pv = *reinterpret_cast<double2*>(p + k);
            const double2 sv = *reinterpret_cast<const double2*>(s + k);
            pv.x = __fma_rn(alpha, sv.x, pv.x); pv.y = __fma_rn(alpha, sv.y, pv.y);
            *reinterpret_cast<double2*>(p + k) = pv;
        }
    }
}

__global__ void pcgParamsKernel(DevCtl* ctl, double tol, int maxIters) {
    ctl->tol = tol;
    ctl->maxIters = maxIters;
}

// updateVelocity (:476-542): pressure gradient with ghost-pressure forms, solid faces zeroed, faces without a
// fluid side flagged unknown; the last column of u and the last row of v are outside the reference's loops and
// keep their values as "known".
__global__ void updateVelocityKernel(const uint8_t* __restrict__ cell, const double* __restrict__ phi,
                                     const double* __restrict__ p, const double* __restrict__ u,
                                     const double* __restrict__ v, int nx, int ny, int pitch, double scale,
                                     double* __restrict__ nu, double* __restrict__ nv, uint8_t* __restrict__ unkU,
                                     uint8_t* __restrict__ unkV, int* anyKnown, unsigned long long* vmaxBits) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    unsigned long long vb = 0;  // bit pattern of the largest |value| this thread leaves on a known face (NaN sorts above everything)
    const bool inside = i <= nx && j <= ny;
    long long o = (long long)j * pitch + i;
    if (inside && j < ny) {  // u face (i, j)
        double val = u[o];
        uint8_t unk = 0;
        if (i < nx) {
            uint8_t c = cell[o], cl = i > 0 ? cell[o - 1] : (uint8_t)FSIM_CELL_SOLID;
            if ((i > 0 && cl == FSIM_CELL_FLUID) || c == FSIM_CELL_FLUID) {
                if ((i == 0 || cl == FSIM_CELL_SOLID) || c == FSIM_CELL_SOLID) val = 0;
                else if (i > 0 && cl == FSIM_CELL_EMPTY) val -= scale * (1 - amlMax(phi[o - 1] / phi[o], -1e3)) * p[o];
                else if (c == FSIM_CELL_EMPTY) val -= scale * (amlMax(phi[o] / phi[o - 1], -1e3) - 1) * p[o - 1];
                else if (i > 0) val -= scale * (p[o] - p[o - 1]);
                else val = 0;
            } else unk = 1;
        }
        nu[o] = val;
        unkU[o] = unk;
        if (!unk && anyKnown[0] == 0) anyKnown[0] = 1;
        if (!unk) vb = (unsigned long long)__double_as_longlong(fabs(val));
    }
    if (inside && i < nx) {  // v face (i, j)
        double val = v[o];
        uint8_t unk = 0;
        if (j < ny) {
            uint8_t c = cell[o], cd = j > 0 ? cell[o - pitch] : (uint8_t)FSIM_CELL_SOLID;
            if ((j > 0 && cd == FSIM_CELL_FLUID) || c == FSIM_CELL_FLUID) {
                if ((j == 0 || cd == FSIM_CELL_SOLID) || c == FSIM_CELL_SOLID) val = 0;
                else if (j > 0 && cd == FSIM_CELL_EMPTY) val -= scale * (1 - amlMax(phi[o - pitch] / phi[o], -1e3)) * p[o];
                else if (c == FSIM_CELL_EMPTY) val -= scale * (amlMax(phi[o] / phi[o - pitch], -1e3) - 1) * p[o - pitch];
                else if (j > 0) val -= scale * (p[o] - p[o - pitch]);
                else val = 0;
            } else unk = 1;
        }
        nv[o] = val;
        unkV[o] = unk;
        if (!unk && anyKnown[1] == 0) anyKnown[1] = 1;
        if (!unk) { const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(val)); vb = b > vb ? b : vb; }
    }
    if (vmaxBits) {  // (uniform) one atomic per block at most, none when the block cannot raise the maximum
        __shared__ unsigned long long wmax[8];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, vb, d); vb = t > vb ? t : vb; }
        const int tid = threadIdx.y * blockDim.x + threadIdx.x;
        if ((tid & 31) == 0) wmax[tid >> 5] = vb;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; ++w) vb = wmax[w] > vb ? wmax[w] : vb;
            if (vb > *reinterpret_cast<volatile unsigned long long*>(vmaxBits)) atomicMax(vmaxBits, vb);
        }
    }
}

// How many BFS layers of the extrapolation the particle stages of this frame can read (stageUpdateVelocity's split).
// A particle's cell is FLUID -- all four faces known, layer 0 -- or, rarely, EMPTY / SOLID (the level set measures from the
// cell's corner, and nothing keeps particles out of interior solids): particleCellDistKernel takes the largest layer D of
// a face of such a cell.  Both the grid-to-particle transfer (bilinear) and the RK3 advection (Catmull-Rom, stencil
// x-1 .. x+2, MAC half-cell shift) sample the grid within R = ceil(c) + 3 cells (L-inf) of the particle's cell, where c
// bounds the displacement in cells: every value on the grid is a known value or a mean of known values, so |v| <= vmax;
// a Catmull-Rom sample is at most 1.25^2 = 1.5625 times that; the three RK3 stages move a particle by at most dt times
// the largest stage velocity.  A face within R cells (L-inf) of a face of layer <= D has BFS layer <= D + 2R.
__global__ void particleCellDistKernel(const uint8_t* __restrict__ cell, const uint32_t* __restrict__ cellStart,
                                       const int* __restrict__ distU, const int* __restrict__ distV, int nx, int ny, int pitch,
                                       int* partDist) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nx || j >= ny) return;
    const long long o = (long long)j * pitch + i;
    if (cell[o] == FSIM_CELL_FLUID) return;
    const size_t c = (size_t)j * nx + i;
    if (cellStart[c + 1] == cellStart[c]) return;
    const int d = max(max(distU[o], distU[o + 1]), max(distV[o], distV[o + pitch]));
    if (d > *reinterpret_cast<volatile int*>(partDist)) atomicMax(partDist, d);
}

__global__ void nearLayersKernel(DevCtl* ctl, double dtOverDx) {
    const double vmax = __longlong_as_double((long long)ctl->vmaxBits);
    const double c = 1.5625 * vmax * dtOverDx;
    const int R = (c < 1048576.0) ? (int)ceil(c) + 3 : (1 << 22);  // (NaN takes the else branch: no split)
    const int D = ctl->partDist < (1 << 24) ? ctl->partDist : (1 << 24);
    ctl->nearLayers = 2 * R + 2 + D;
}

// The unknown masks of updateVelocityKernel alone (they depend on the labels only): lets the structure of the
// extrapolation that follows updateVelocity be built beside the projection (prepareVelocityExtrapolation).
__global__ void velocityMaskKernel(const uint8_t* __restrict__ cell, int nx, int ny, int pitch, uint8_t* __restrict__ unkU,
                                   uint8_t* __restrict__ unkV, int* anyKnown) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i > nx || j > ny) return;
    long long o = (long long)j * pitch + i;
    if (j < ny) {
        uint8_t unk = 0;
        if (i < nx) {
            uint8_t c = cell[o], cl = i > 0 ? cell[o - 1] : (uint8_t)FSIM_CELL_SOLID;
            unk = ((i > 0 && cl == FSIM_CELL_FLUID) || c == FSIM_CELL_FLUID) ? 0 : 1;
        }
        unkU[o] = unk;
        if (!unk && anyKnown[0] == 0) anyKnown[0] = 1;
    }
    if (i < nx) {
        uint8_t unk = 0;
        if (j < ny) {
            uint8_t c = cell[o], cd = j > 0 ? cell[o - pitch] : (uint8_t)FSIM_CELL_SOLID;
            unk = ((j > 0 && cd == FSIM_CELL_FLUID) || c == FSIM_CELL_FLUID) ? 0 : 1;
        }
        unkV[o] = unk;
        if (!unk && anyKnown[1] == 0) anyKnown[1] = 1;
    }
}

}  // namespace

int pcgSetParams(Sim* s, double tol, int maxIters) {
    pcgParamsKernel<<<1, 1, 0, s->stream>>>(s->ctl, tol, maxIters);
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}

constexpr int SD_SUBS = 16;  // steps per solver sub-chunk (hand-off granularity), tuned with tools/wavebench.cu

static int sdClusterSize() {
    static int cl = -1;
    if (cl < 0) {
        cl = 8;
        if (const char* e = getenv("FSIM_SD_CLUSTER")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) cl = v; }  // tuning knob
    }
    return cl;
}

template <class Op, int DIR>
static int launchSdSolve(Sim* s, const Op& op, const sd::Geom& g) {
    static int dbgPre = -1;
    if (dbgPre < 0) { const char* e = getenv("FSIM_DBG_PRE"); dbgPre = e ? atoi(e) : 0; }
    sd::Control ctl{s->wfTicket, s->wfFinished, s->sdHand, &s->ctl->pcgDone, nullptr, s->opt.reserved[2] == 1 ? nullptr : s->sdRange, dbgPre};
    const int cl = sdClusterSize();
    if (g.rpl == 2) {
        switch (g.sigma) {
            case 1: CUDA_TRY((sd::launchSolveR<Op, 2, 1, DIR, SD_SUBS>(op, g, ctl, s->stream, cl))); break;
            case 2: CUDA_TRY((sd::launchSolveR<Op, 2, 2, DIR, SD_SUBS>(op, g, ctl, s->stream, cl))); break;
            case 3: CUDA_TRY((sd::launchSolveR<Op, 2, 3, DIR, SD_SUBS>(op, g, ctl, s->stream, cl))); break;
            default: fsim_set_error("unsupported SD skew %d", g.sigma); return FSIM_E_INVALID;
        }
        LAUNCH_COUNT(s);
        return FSIM_OK;
    }
    switch (g.sigma) {
        case 2: CUDA_TRY((sd::launchSolve<Op, 2, DIR, SD_SUBS>(op, g, ctl, s->stream, cl))); break;
        case 3: CUDA_TRY((sd::launchSolve<Op, 3, DIR, SD_SUBS>(op, g, ctl, s->stream, cl))); break;
        default: fsim_set_error("unsupported SD skew %d", g.sigma); return FSIM_E_INVALID;
    }
    LAUNCH_COUNT(s);
    return FSIM_OK;
}

// z = M^-1 r (:397-421): forward solve (with sigma/beta from q.q) then backward solve
// (g, off): the whole grid, or the own strips of a y-slab (arrays offset by one halo strip)
int distPeerView(Sim* s, PeerView* pv);
static inline double* peerGhostHost(PeerBlock* b, int which, int ghostPitch) {
    return reinterpret_cast<double*>(reinterpret_cast<char*>(b) + DIST_GHOST_OFF) + (size_t)which * ghostPitch;
}
// the neighbours' ghost rows and stamps this rank's backward solve writes (peer memory)
static int dbgDist() {  // timing experiments only (bit 64: no halo stamps and no wait for them)
    static int v = -1;
    if (v < 0) { const char* e = getenv("FSIM_DBG_PRE"); v = e ? atoi(e) : 0; }
    return v;
}
template <class OpD>
static void haloTargets(Sim* s, OpD& b) {
    PeerView pv;
    distPeerView(s, &pv);
    b.pushLo = b.pushHi = nullptr; b.stampLo = b.stampHi = nullptr;
    b.stampBase = pv.stampBase;
    if (pv.rank > 0) {  // my first row is the row just above rank-1's slab
        b.pushLo = peerGhostHost(pv.blk[pv.rank - 1], 1, pv.ghostPitch);
        b.stampLo = &pv.blk[pv.rank - 1]->haloSeq[1];
    }
    if (pv.rank < pv.world - 1) {
        b.pushHi = peerGhostHost(pv.blk[pv.rank + 1], 0, pv.ghostPitch);
        b.stampHi = &pv.blk[pv.rank + 1]->haloSeq[0];
    }
    if (dbgDist() & 64) b.stampLo = b.stampHi = nullptr;
}
static int forwardSolve(Sim* s, int phase, const sd::Geom& g, size_t off) {
    OpForward f;
    f.in[0] = s->sR + off; f.in[1] = s->sLx + off; f.in[2] = s->sLy + off; f.in[3] = s->sD + off; f.out = s->sT + off;
    f.partials = s->partials; f.ctl = s->ctl; f.phase = phase;
    f.dist = s->dist.on ? 1 : 0;
    memset(&f.pv, 0, sizeof(f.pv));
    if (s->dist.on) distPeerView(s, &f.pv);
    profBegin(s, 2);
    int rc = launchSdSolve<OpForward, +1>(s, f, g);
    profEnd(s);
    return rc;
}
static int backwardSolve(Sim* s, int first, const sd::Geom& g, size_t off) {
    OpBackwardD b;
    b.in[0] = s->sT + off; b.in[1] = s->sUx + off; b.in[2] = s->sUy + off; b.in[3] = s->sS + off; b.out = s->sS + off;
    b.ctl = s->ctl; b.first = first;
    profBegin(s, 3);
    int rc;
    if (s->dist.on) {
        if (g.rpl != 2) { fsim_set_error("the y-slab projection needs the two-rows-per-lane layout"); return FSIM_E_STATE; }
        haloTargets(s, b);
        rc = launchSdSolve<OpBackwardD, -1>(s, b, g);
    } else {
        rc = launchSdSolve<OpBackward, -1>(s, static_cast<const OpBackward&>(b), g);
    }
    profEnd(s);
    return rc;
}

// the fused variants (default on the two-rows-per-lane layout; the only ones the y-slab mode uses)
static bool pcgFused(const Sim* s) {
    static int envFused = -1;
    if (envFused < 0) { const char* e = getenv("FSIM_FUSED_AXPY"); envFused = e && atoi(e) ? 1 : 0; }  // A/B knob
    return s->sdg.rpl == 2 && (s->opt.reserved[FSIM_OPT_FUSED_AXPY] == 1 || envFused);
}
static int forwardSolveF(Sim* s, int phase, const sd::Geom& g, size_t off = 0) {
    OpForwardF f;
    f.in[0] = s->sR + off; f.in[1] = s->sLx + off; f.in[2] = s->sLy + off; f.in[3] = s->sD + off; f.in[4] = s->sZ + off;
    f.out = s->sT + off; f.rOut = s->sR + off;
    f.partials = s->partials; f.ctl = s->ctl; f.phase = phase;
    f.dist = s->dist.on ? 1 : 0;
    memset(&f.pv, 0, sizeof(f.pv));
    if (s->dist.on) distPeerView(s, &f.pv);
    profBegin(s, 2);
    int rc = launchSdSolve<OpForwardF, +1>(s, f, g);
    profEnd(s);
    return rc;
}
static int backwardSolveF(Sim* s, int first, const sd::Geom& g, size_t off = 0) {
    OpBackwardFD b;
    b.in[0] = s->sT + off; b.in[1] = s->sUx + off; b.in[2] = s->sUy + off; b.in[3] = s->sS + off; b.in[4] = s->sP + off;
    b.out = s->sS + off; b.out2 = s->sP + off;
    b.ctl = s->ctl; b.first = first;
    profBegin(s, 3);
    int rc;
    if (s->dist.on) {
        haloTargets(s, b);
        rc = launchSdSolve<OpBackwardFD, -1>(s, b, g);
    } else {
        rc = launchSdSolve<OpBackwardF, -1>(s, static_cast<const OpBackwardF&>(b), g);
    }
    profEnd(s);
    return rc;
}

// FSIM_AXPY_SPLIT=1 (experiment, off by default): r -= alpha z (+ |r|_inf, stop rule) on the main stream, p += alpha s on the
// axpy stream beside the forward solve; the caller joins (joinAxpyP) before the backward solve overwrites s.  Same bits.
// Measured at 4096^2: the residual update alone takes 0.027 ms instead of 0.043, but the p update's blocks land on the SMs the
// solver warps live on and the forward solve slows from 0.152 to 0.170 ms -- no net gain (124.3 against 123.7 ms per step).
static bool axpySplit() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FSIM_AXPY_SPLIT"); v = e ? atoi(e) != 0 : 0; }
    return v != 0;
}
static int launchAxpy(Sim* s, const sd::Geom& g, size_t off, int launchIter, int dist, const PeerView& pv) {
    const bool split = axpySplit();
    profBegin(s, 1);
    axpyKernel<<<AA_BLOCKS, 256, 0, s->stream>>>(s->sP + off, s->sR + off, s->sS + off, s->sZ + off, g.nchunks, g.nstrips, g.rpl, s->sdRange,
                                                 s->partials, &s->counters[4], s->ctl, dist, pv, split ? 0 : 1);
    profEnd(s);
    LAUNCH_COUNT(s);
    if (split) {
        // behind the residual update, so that it shares the GPU with the latency-bound forward solve (48 of 148 SMs), not
        // with the bandwidth-bound kernel before it
        CUDA_TRY(cudaEventRecord(s->evAxpyA, s->stream));
        CUDA_TRY(cudaStreamWaitEvent(s->axpyStream, s->evAxpyA, 0));
        axpyPKernel<<<AA_BLOCKS, 256, 0, s->axpyStream>>>(s->sP + off, s->sS + off, g.nchunks, g.nstrips, g.rpl, s->sdRange, s->ctl, launchIter);
        CUDA_TRY(cudaEventRecord(s->evAxpyP, s->axpyStream));
        LAUNCH_COUNT(s);
    }
    return FSIM_OK;
}
static int joinAxpyP(Sim* s) {
    if (axpySplit()) CUDA_TRY(cudaStreamWaitEvent(s->stream, s->evAxpyP, 0));
    return FSIM_OK;
}

static bool factorLegacy(const Sim* s) {
    static int legacy = -1;
    if (legacy < 0) { const char* e = getenv("FSIM_FACTOR_LEGACY"); legacy = e && atoi(e) ? 1 : 0; }
    return legacy || s->opt.debugSimpleWavefront;
}

// MIC(0) factor (:364-388) of the rows [j0, j0 + 32*nstrips) taken as an independent block (the whole grid, or a
// y-slab): frames -> SD (skew 1) -> in-place sweep -> precon frame
static int factorRows(Sim* s, int j0, int nstrips, int nxEff) {
    const Frame& f = s->fr;
    const int nx = nxEff, ny = s->ny, ncb = (nx + 31) / 32;  // (columns beyond nxEff hold no fluid: precon stays 0)
    const long long rowOff = (long long)j0 * f.pitch;
    if (factorLegacy(s)) {
        OpFactor fac;
        fac.in[0] = s->Adiag + rowOff; fac.in[1] = s->Ax + rowOff; fac.in[2] = s->Ay + rowOff; fac.in[3] = s->fmask + rowOff;
        fac.out[0] = s->pc + rowOff;
        fac.nx = nx; fac.ny = ny; fac.jOff = j0;
        return launchWavefront<OpFactor, +1, +1>(s, fac, ncb, nstrips, nullptr, 0, nullptr);
    }
    sd::Geom g = sd::makeGeom(nx, 32 * nstrips, 1);
    sd::Geom gp = g;
    gp.ny = ny - j0 < g.ny ? ny - j0 : g.ny;  // rows beyond the grid read as zero
    dim3 blk(32, 8);
    sd::PackJob job;
    const double* src[4] = {s->Ax, s->Ay, s->Adiag, s->fmask};
    double* dst[4] = {s->sT, s->sP, s->sZ, s->sR};
    for (int k = 0; k < 4; ++k) { job.src[k] = src[k] + rowOff; job.dst[k] = dst[k]; }
    sd::sdPackKernel<<<dim3(g.nchunks, g.nstrips, 4), blk, 0, s->stream>>>(job, gp, f.pitch, 0);
    CUDA_TRY(cudaMemsetAsync(s->sS, 0, g.elems * sizeof(double), s->stream));
    OpSdFactor op;
    op.arr[0] = s->sS; op.arr[1] = s->sT; op.arr[2] = s->sP; op.arr[3] = s->sZ; op.arr[4] = s->sR;
    op.nx = nx; op.ny = ny; op.jOff = j0;
    sd::SweepControl ctl{s->wfTicket, s->wfFinished, s->swHand, s->swPlaneWords, nullptr, nullptr};
    profBegin(s, 8);
    CUDA_TRY((sd::launchSweep<OpSdFactor, 1, +1, 8>(op, g, ctl, s->stream, sdClusterSize())));
    profEnd(s);
    sd::PackJob uj;
    uj.src[0] = s->sS; uj.dst[0] = s->pc + rowOff;
    sd::sdUnpackKernel<<<dim3(g.nchunks, g.nstrips, 1), blk, 0, s->stream>>>(uj, gp, f.pitch, 0);
    s->launches += 3;
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}

static int stageApplyProjectionDist(Sim* s);

__global__ void bboxResetKernel(DevCtl* ctl) {
    ctl->bbox[0] = 0x7fffffff; ctl->bbox[1] = -1; ctl->bbox[2] = 0x7fffffff; ctl->bbox[3] = -1;
    ctl->marchedSlots = 0;
    ctl->pendingP = 0;
    ctl->distError = 0;
    ctl->alphaIter = 0;
}


int stageApplyProjection(Sim* s) {
    if (s->dist.on) return stageApplyProjectionDist(s);
    const Frame& f = s->fr;
    const int nx = s->nx, ny = s->ny;
    double scaleA = s->dt / (s->rho * s->dx * s->dx);  // :261
    double invDx = 1.0 / s->dx;                        // :339
    dim3 blk(32, 8), grd((nx + 31) / 32, (ny + 7) / 8);
    bboxResetKernel<<<1, 1, 0, s->stream>>>(s->ctl);
    assembleKernel<<<grd, blk, 0, s->stream>>>(s->cell, s->phi, s->u, s->v, nx, ny, f.pitch, scaleA, invDx, s->Adiag,
                                               s->Ax, s->Ay, s->rhs, s->fmask, s->r, s->p, s->partials, &s->counters[2],
                                               s->ctl);
    s->launches += 2;
    // The system only couples FLUID cells; outside their bounding box every PCG vector is exactly zero.  The solve
    // (factor, SD layout, all PCG kernels) therefore runs on the strips and columns of that box only.  This is the one
    // place where the host waits for the device inside a step (16 bytes).
    CUDA_TRY(cudaMemcpyAsync(s->hBox, s->ctl->bbox, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (s->hBox[1] < 0 || s->opt.reserved[1] == 1) { s->hBox[0] = 0; s->hBox[1] = nx - 1; s->hBox[2] = 0; s->hBox[3] = ny - 1; }  // no fluid / box disabled
    const int R = s->sdg.rpl, SR = 32 * R;  // rows per strip of the PCG layout
    const int strip0 = s->hBox[2] / SR, nstrips = s->hBox[3] / SR + 1 - strip0;
    const int j0 = SR * strip0;
    const int nxb = s->hBox[1] + 1 < nx ? s->hBox[1] + 2 : nx;  // columns [0, nxb)
    const int ncb = (nxb + 31) / 32;
    const long long rowOff = (long long)j0 * f.pitch;
    const sd::Geom g = sd::makeGeom(nxb, SR * nstrips, s->sdg.sigma, R);
    sd::Geom gp = g;  // rows beyond the grid read as zero
    gp.ny = ny - j0 < g.ny ? ny - j0 : g.ny;
    int rc = factorRows(s, j0, nstrips * R, nxb);
    if (rc) return rc;
    dim3 grdP(ncb, (nstrips * SR + 7) / 8);
    deriveKernel<<<grdP, blk, 0, s->stream>>>(s->pc + rowOff, s->Ax + rowOff, s->Ay + rowOff, ncb * 32, nstrips * SR, f.pitch,
                                              s->D + rowOff, s->Ux + rowOff, s->Uy + rowOff, s->Lx + rowOff, s->Ly + rowOff, 1);
    LAUNCH_COUNT(s);
    // everything the solve touches moves to the strip-diagonal layout; p = 0 (:424)
    sd::PackJob job;
    const double* srcs[9] = {s->Adiag, s->Ax, s->Ay, s->Lx, s->Ly, s->D, s->Ux, s->Uy, s->rhs};
    double* dsts[9] = {s->sAd, s->sAx, s->sAy, s->sLx, s->sLy, s->sD, s->sUx, s->sUy, s->sR};
    for (int k = 0; k < 9; ++k) { job.src[k] = srcs[k] + rowOff; job.dst[k] = dsts[k]; }
    sd::sdPackKernel<<<dim3(g.nchunks, g.nstrips, 9 * R), blk, 0, s->stream>>>(job, gp, f.pitch, 0);
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaMemsetAsync(s->sP, 0, g.elems * sizeof(double), s->stream));
    // the triangular solves only march, per strip, the chunks that hold fluid; outside them their outputs stay zero
    stripRangeKernel<<<g.nstrips, 256, 0, s->stream>>>(s->fmask + rowOff, f.pitch, nxb, gp.ny, g.sigma, g.rpl, s->sdRange, s->ctl);
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaMemsetAsync(s->sT, 0, g.elems * sizeof(double), s->stream));
    CUDA_TRY(cudaMemsetAsync(s->sZ, 0, g.elems * sizeof(double), s->stream));
    // r = rhs; z = M^-1 r; s = z; sigma = z.r (:424-428)
    CUDA_TRY(cudaMemsetAsync(s->sS, 0, g.elems * sizeof(double), s->stream));
    const bool fused = pcgFused(s);
    if ((rc = fused ? forwardSolveF(s, 0, g) : forwardSolve(s, 0, g, 0))) return rc;
    if ((rc = fused ? backwardSolveF(s, 1, g) : backwardSolve(s, 1, g, 0))) return rc;

    // from here on the stream holds latency-bound solves: the side work runFrame left pending starts behind the set-up
    if (s->prepPending && (rc = forkExtrapolationPrepare(s))) return rc;
    const int batch = 8;
    const int maxIters = s->opt.pcgMaxIters;
    int nbatches = (maxIters + batch - 1) / batch + 1;
    for (int b = 0; b < nbatches; ++b) {
        for (int k = 0; k < batch; ++k) {
            profBegin(s, 0);
            applyASdKernel<<<AA_BLOCKS, 256, 0, s->stream>>>(s->sAd, s->sAx, s->sAy, s->sS, s->sZ, g, 0, g.nstrips, s->sdRange, s->partials,
                                                                    &s->counters[3], s->ctl, ApplyADist{});
            profEnd(s);
            LAUNCH_COUNT(s);
            if (fused) {  // the axpys ride on the solves
                if ((rc = forwardSolveF(s, 1, g))) return rc;
                if ((rc = backwardSolveF(s, 0, g))) return rc;
                continue;
            }
            if ((rc = launchAxpy(s, g, 0, b * batch + k + 1, 0, PeerView{}))) return rc;
            if ((rc = forwardSolve(s, 1, g, 0))) return rc;
            if ((rc = joinAxpyP(s))) return rc;  // the backward solve overwrites s
            if ((rc = backwardSolve(s, 0, g, 0))) return rc;
        }
        int slot = b & 1;
        CUDA_TRY(cudaMemcpyAsync(&s->hPcgFlags[slot], &s->ctl->pcgDone, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaEventRecord(s->pollEv[slot], s->stream));
        if (b >= 1) {
            // look at the previous batch's flag while this batch is already queued: no bubble on the stream
            CUDA_TRY(cudaEventSynchronize(s->pollEv[slot ^ 1]));
            if (s->hPcgFlags[slot ^ 1]) break;
        }
    }
    if (fused) {
        pcgFinishKernel<<<AA_BLOCKS, 256, 0, s->stream>>>(s->sP, s->sS, g.nchunks, g.nstrips, g.rpl, s->sdRange, s->ctl);
        LAUNCH_COUNT(s);
    }
    sd::PackJob uj;
    uj.src[0] = s->sP; uj.dst[0] = s->p + rowOff;
    sd::sdUnpackKernel<<<dim3(g.nchunks, g.nstrips, R), blk, 0, s->stream>>>(uj, gp, f.pitch, 0);
    LAUNCH_COUNT(s);
    s->lastSolveCells = (long long)g.nx * gp.ny;
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}

// ------------------------------------------------------------------------------------------------------
// y-slab PCG over the GPUs of one node.  Every rank holds the full replicated state and
// assembles the whole system (0.4 ms); it then factors, packs and solves only its own rows [j0, j1) (whole strips).
// The iteration is the single-GPU one -- applyA + z.s, forward solve (with r -= alpha z, |r|_inf, z.r), backward solve
// (with s = z + beta s, p += alpha s) -- and talks to the other ranks through peer memory only (distpeer.cuh): the
// backward solve writes its boundary rows of s into the neighbours' ghost rows, the threads that finish z.s and z.r /
// |r|_inf exchange their partials slot by slot.  No NCCL call and no extra kernel inside the loop.  Restricting the factor
// and the solves to the slab makes the preconditioner block-MIC(0) (the Ay coupling across slab boundaries is dropped in
// M, not in A); bench.py and the tests report the iteration delta.  The rows of p are exchanged at the end (NCCL, once per
// step) so that the replicated stages that follow see the whole pressure field.
// ------------------------------------------------------------------------------------------------------
int distShareRows(Sim* s, double* frame);
void distSlabOf(int ns, int world, int r, int* strip0, int* nOwn);

static int stageApplyProjectionDist(Sim* s) {
    const Frame& f = s->fr;
    Sim::Dist& d = s->dist;
    const int nx = s->nx, ny = s->ny;
    if (s->sdg.rpl != 2) { fsim_set_error("the y-slab projection needs the two-rows-per-lane PCG layout"); return FSIM_E_STATE; }
    const bool fused = pcgFused(s);
    if (s->opt.pcgMaxIters > 60000) { fsim_set_error("the y-slab projection supports at most 60000 PCG iterations"); return FSIM_E_INVALID; }
    double scaleA = s->dt / (s->rho * s->dx * s->dx);
    double invDx = 1.0 / s->dx;
    int rc;
    dim3 blk(32, 8), grd((nx + 31) / 32, (ny + 7) / 8);
    // the whole system, replicated (|rhs|_inf is global this way)
    bboxResetKernel<<<1, 1, 0, s->stream>>>(s->ctl);
    assembleKernel<<<grd, blk, 0, s->stream>>>(s->cell, s->phi, s->u, s->v, nx, ny, f.pitch, scaleA, invDx, s->Adiag,
                                               s->Ax, s->Ay, s->rhs, s->fmask, s->r, s->p, s->partials, &s->counters[2],
                                               s->ctl);
    s->launches += 2;
    // The slabs partition the strips of the fluid cells' bounding box (identical on every rank: the state is
    // replicated), so the ranks stay balanced whatever the fluid does; the columns stop at the box as well.
    CUDA_TRY(cudaMemcpyAsync(s->hBox, s->ctl->bbox, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (s->hBox[1] < 0) { s->hBox[0] = 0; s->hBox[1] = nx - 1; s->hBox[2] = 0; s->hBox[3] = ny - 1; }
    const int R = s->sdg.rpl, SR = 32 * R;  // rows per strip of the PCG layout
    const int nsAll = (ny + SR - 1) / SR;
    int S0 = s->hBox[2] / SR, S1 = s->hBox[3] / SR + 1;
    while (S1 - S0 < d.world) { if (S1 < nsAll) ++S1; else --S0; }  // at least one strip per rank
    d.boxStrip0 = S0; d.boxStrips = S1 - S0;
    distSlabOf(d.boxStrips, d.world, d.rank, &d.strip0, &d.nOwn);
    d.strip0 += S0;
    d.j0 = SR * d.strip0;
    d.j1 = d.j0 + SR * d.nOwn < ny ? d.j0 + SR * d.nOwn : ny;
    const int nxb = s->hBox[1] + 1 < nx ? s->hBox[1] + 2 : nx;
    d.gExt = sd::makeGeom(nxb, SR * (d.nOwn + 2), s->sdg.sigma, R);
    d.gOwn = sd::makeGeom(nxb, SR * d.nOwn, s->sdg.sigma, R);
    const sd::Geom& gE = d.gExt;
    const sd::Geom& gO = d.gOwn;
    const int ncb = (nxb + 31) / 32;
    const size_t own = (size_t)gE.Sp * SR;  // offset of the first own strip in the slab's SD arrays
    d.epoch = (d.epoch + 1) & 0xffff;
    if (d.epoch == 0) d.epoch = 1;
    PeerView pv;
    distPeerView(s, &pv);
    // block-MIC(0): factor of the own rows only, no coupling to the row below j0
    const long long rowOff = (long long)d.j0 * f.pitch;
    if ((rc = factorRows(s, d.j0, d.nOwn * R, nxb))) return rc;
    dim3 grdP((ncb * 32 + 31) / 32, (d.nOwn * SR + 7) / 8);
    deriveKernel<<<grdP, blk, 0, s->stream>>>(s->pc + rowOff, s->Ax + rowOff, s->Ay + rowOff, ncb * 32, d.nOwn * SR, f.pitch,
                                              s->D + rowOff, s->Ux + rowOff, s->Uy + rowOff, s->Lx + rowOff, s->Ly + rowOff, 1);
    LAUNCH_COUNT(s);
    // A (with its true coupling across the slab boundary) over the slab plus halo strips; solver coefficients and
    // rhs over the own strips.  Slab row 0 is global row j0 - SR (the frame's zero halo for rank 0).
    const long long extOff = (long long)(d.j0 - SR) * f.pitch;
    sd::PackJob job;
    const double* srcsA[3] = {s->Adiag, s->Ax, s->Ay};
    double* dstsA[3] = {s->sAd, s->sAx, s->sAy};
    for (int k = 0; k < 3; ++k) { job.src[k] = srcsA[k] + extOff; job.dst[k] = dstsA[k]; }
    // (rows beyond the grid must read as zero: gExt.ny is clipped to the grid by the row limit below)
    sd::Geom gPack = gE;
    gPack.ny = ny - (d.j0 - SR) < gE.ny ? ny - (d.j0 - SR) : gE.ny;
    // (the halo strip of the lowest slab lies below the grid -- and below the frame's own 32-row halo when SR > 32)
    sd::sdPackKernel<<<dim3(gE.nchunks, gE.nstrips, 3 * R), blk, 0, s->stream>>>(job, gPack, f.pitch, 0, d.j0 - SR < 0 ? SR - d.j0 : 0);
    const double* srcsO[6] = {s->Lx, s->Ly, s->D, s->Ux, s->Uy, s->rhs};
    double* dstsO[6] = {s->sLx, s->sLy, s->sD, s->sUx, s->sUy, s->sR};
    for (int k = 0; k < 6; ++k) { job.src[k] = srcsO[k] + rowOff; job.dst[k] = dstsO[k] + own; }
    sd::Geom gPackO = gO;
    gPackO.ny = ny - d.j0 < gO.ny ? ny - d.j0 : gO.ny;
    sd::sdPackKernel<<<dim3(gO.nchunks, gO.nstrips, 6 * R), blk, 0, s->stream>>>(job, gPackO, f.pitch, 0);
    s->launches += 2;
    CUDA_TRY(cudaMemsetAsync(s->sP, 0, gE.elems * sizeof(double), s->stream));
    CUDA_TRY(cudaMemsetAsync(s->sS, 0, gE.elems * sizeof(double), s->stream));
    CUDA_TRY(cudaMemsetAsync(s->sZ, 0, gE.elems * sizeof(double), s->stream));
    CUDA_TRY(cudaMemsetAsync(s->sT, 0, gE.elems * sizeof(double), s->stream));
    // ghost rows: columns the neighbours do not march this step must read as zero (their s is exactly zero there).  The
    // neighbours' first writes come after their first forward solve has seen ours, i.e. after this memset (stream order).
    CUDA_TRY(cudaMemsetAsync(peerGhostHost(pv.blk[pv.rank], 0, pv.ghostPitch), 0, (size_t)2 * pv.ghostPitch * sizeof(double), s->stream));
    stripRangeKernel<<<gO.nstrips, 256, 0, s->stream>>>(s->fmask + rowOff, f.pitch, nxb, gPackO.ny, gO.sigma, gO.rpl, s->sdRange, s->ctl);
    LAUNCH_COUNT(s);
    // r = rhs; z = M^-1 r; s = z; sigma = z.r (:424-428)
    if ((rc = fused ? forwardSolveF(s, 0, gO, own) : forwardSolve(s, 0, gO, own))) return rc;
    if ((rc = fused ? backwardSolveF(s, 1, gO, own) : backwardSolve(s, 1, gO, own))) return rc;

    if (s->prepPending && (rc = forkExtrapolationPrepare(s))) return rc;  // (as in stageApplyProjection)
    ApplyADist dd;
    dd.dist = 1;
    dd.ghostLo = d.rank > 0 ? peerGhostHost(pv.blk[pv.rank], 0, pv.ghostPitch) : nullptr;
    dd.ghostHi = d.rank < d.world - 1 ? peerGhostHost(pv.blk[pv.rank], 1, pv.ghostPitch) : nullptr;
    dd.haloSeqLo = d.rank > 0 ? &pv.blk[pv.rank]->haloSeq[0] : nullptr;
    dd.haloSeqHi = d.rank < d.world - 1 ? &pv.blk[pv.rank]->haloSeq[1] : nullptr;
    if (dbgDist() & 64) dd.haloSeqLo = dd.haloSeqHi = nullptr;
    dd.pv = pv;
    const int batch = 8;
    const int maxIters = s->opt.pcgMaxIters;
    int nbatches = (maxIters + batch - 1) / batch + 1;
    for (int b = 0; b < nbatches; ++b) {
        for (int k = 0; k < batch; ++k) {
            profBegin(s, 0);
            applyASdKernel<<<AA_BLOCKS, 256, 0, s->stream>>>(s->sAd, s->sAx, s->sAy, s->sS, s->sZ, gE, 1, 1 + d.nOwn, s->sdRange, s->partials,
                                                             &s->counters[3], s->ctl, dd);
            profEnd(s);
            LAUNCH_COUNT(s);
            if (fused) {
                if ((rc = forwardSolveF(s, 1, gO, own))) return rc;
                if ((rc = backwardSolveF(s, 0, gO, own))) return rc;
                continue;
            }
            if ((rc = launchAxpy(s, gO, own, b * batch + k + 1, 1, pv))) return rc;
            if ((rc = forwardSolve(s, 1, gO, own))) return rc;
            if ((rc = joinAxpyP(s))) return rc;
            if ((rc = backwardSolve(s, 0, gO, own))) return rc;
        }
        int slot = b & 1;
        CUDA_TRY(cudaMemcpyAsync(&s->hPcgFlags[slot], &s->ctl->pcgDone, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaEventRecord(s->pollEv[slot], s->stream));
        if (b >= 1) {
            // (the flag becomes 1 on every rank in the same iteration: all scalars are combined identically everywhere;
            // ranks that notice it a batch later only enqueue gated no-ops, which neither push nor wait)
            CUDA_TRY(cudaEventSynchronize(s->pollEv[slot ^ 1]));
            if (s->hPcgFlags[slot ^ 1]) break;
        }
    }
    pcgFinishKernel<<<AA_BLOCKS, 256, 0, s->stream>>>(s->sP + own, s->sS + own, gO.nchunks, gO.nstrips, gO.rpl, s->sdRange, s->ctl);  // (no-op unless fused)
    // own rows of p back to the frame, then every rank's rows to every rank
    sd::PackJob uj;
    uj.src[0] = s->sP + own; uj.dst[0] = s->p + rowOff;
    sd::sdUnpackKernel<<<dim3(gO.nchunks, gO.nstrips, R), blk, 0, s->stream>>>(uj, gPackO, f.pitch, 0);
    s->launches += 2;
    CUDA_TRY(cudaGetLastError());
    s->lastSolveCells = (long long)gO.nx * gPackO.ny;
    return distShareRows(s, s->p);
}

// masks, distance transform and layer lists of updateVelocity's extrapolation, from the labels alone (enqueued on
// s->stream; runFrame points that at the second stream while the projection runs)
int prepareVelocityExtrapolation(Sim* s) {
    CUDA_TRY(cudaMemsetAsync(&s->ctl->anyKnown[0], 0, 2 * sizeof(int), s->stream));
    dim3 blk(32, 8), grd((s->nx + 1 + 31) / 32, (s->ny + 1 + 7) / 8);
    velocityMaskKernel<<<grd, blk, 0, s->stream>>>(s->cell, s->nx, s->ny, s->fr.pitch, s->unkU, s->unkV, s->ctl->anyKnown);
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaGetLastError());
    return extrapolatePrepare(s, s->unkU, s->unkV);
}

int stageUpdateVelocity(Sim* s) {
    const Frame& f = s->fr;
    const bool prepared = s->extrapReady;  // runFrame built the extrapolation's structure already (same masks)
    s->extrapReady = false;
    if (!prepared) CUDA_TRY(cudaMemsetAsync(&s->ctl->anyKnown[0], 0, 2 * sizeof(int), s->stream));
    double scale = s->dt / (s->rho * s->dx);  // :478
    dim3 blk(32, 8), grd((s->nx + 1 + 31) / 32, (s->ny + 1 + 7) / 8);
    const bool split = s->splitFill && s->stream2 != nullptr;
    if (split) CUDA_TRY(cudaMemsetAsync(&s->ctl->vmaxBits, 0, sizeof(unsigned long long) + 2 * sizeof(int), s->stream));  // vmaxBits, nearLayers, partDist
    updateVelocityKernel<<<grd, blk, 0, s->stream>>>(s->cell, s->phi, s->p, s->u, s->v, s->nx, s->ny, f.pitch, scale,
                                                     s->nu, s->nv, s->unkU, s->unkV, s->ctl->anyKnown,
                                                     split ? &s->ctl->vmaxBits : nullptr);
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaGetLastError());
    if (!split) {
        s->lastSplit = false;
        int rc = prepared ? extrapolateFill(s, s->nu, s->nv, s->unkU, s->unkV) : extrapolatePair(s, s->nu, s->nv, s->unkU, s->unkV);
        if (rc) return rc;
        if (s->mode == FSIM_SEMILAGRANGIAN) return copyNewMacToMac(s);  // :547-549
        return FSIM_OK;
    }
    // Split fill (runFrame only).  What is left of the frame -- grid-to-particle transfer, mac.copyFrom(newMac), particle
    // advection -- reads the extrapolated velocities within a few cells of the fluid only (nearLayersKernel bounds how
    // few); the 2500 layers beyond (4096^2 dam break), 3 us of latency each, are filled by a second launch on the second
    // stream beside those stages.  The second launch writes newMac and mac, the copy in between only touches the faces
    // up to the cut (particles.cu copyMacFromNew), so no face is written by both.  Same kernel, same arithmetic: the
    // result does not depend on the cut.
    if (!prepared) { int rc = extrapolatePrepare(s, s->unkU, s->unkV); if (rc) return rc; }
    if (s->np) {  // (the cell sort of this frame is still valid: positions only change in applyAdvection)
        dim3 grdC((s->nx + 31) / 32, (s->ny + 7) / 8);
        particleCellDistKernel<<<grdC, blk, 0, s->stream>>>(s->cell, s->cellStart, s->distU + f.org, s->distV + f.org, s->nx, s->ny,
                                                            f.pitch, &s->ctl->partDist);
        LAUNCH_COUNT(s);
    }
    nearLayersKernel<<<1, 1, 0, s->stream>>>(s->ctl, s->dt / s->dx);
    LAUNCH_COUNT(s);
    int rc = extrapolateFill(s, s->nu, s->nv, s->unkU, s->unkV, 0);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(s->evNear, s->stream));
    CUDA_TRY(cudaStreamWaitEvent(s->stream2, s->evNear, 0));
    cudaStream_t mainStream = s->stream;
    s->stream = s->stream2;
    rc = extrapolateFill(s, s->nu, s->nv, s->unkU, s->unkV, 1, s->u, s->v);
    s->stream = mainStream;
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(s->evFar, s->stream2));
    s->farPending = true;
    s->lastSplit = true;
    if (s->mode == FSIM_SEMILAGRANGIAN) return copyNewMacToMac(s);  // :547-549 (the faces up to the cut)
    return FSIM_OK;
}
