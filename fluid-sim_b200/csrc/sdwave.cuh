// sdwave.cuh -- the "strip-diagonal" (SD) layout and the wavefront kernels of the MIC(0) triangular solves.
//
// The solves (reference src/FluidSim2D.cpp:397-421) are sequential in raster order: cell (i,j) needs the new
// values at (i-1,j) and (i,j-1) (forward) or (i+1,j) and (i,j+1) (backward).  Any schedule that respects those two
// dependencies reproduces the reference's preconditioner.  On a row-major grid the natural parallel schedule
// (anti-diagonals) reads memory with a stride; here the layout is changed instead:
//
//   * rows are grouped into strips of 32*R (R = rows per lane, Geom::rpl); inside strip k, lane t owns rows
//     R*t .. R*t+R-1 for the whole sweep and visits them at the same column in one step;
//   * lane t is SIGMA columns behind lane t-1, i.e. at step s it works on column c = s - SIGMA*t;
//   * element (row 32R*k + R*t + r, column c) of EVERY PCG vector and coefficient array is stored at
//         [((k*Sp + c + SIGMA*t) * 32 + t) * R + r]          (Sp = steps per strip, padded to 32)
//     so the values a warp needs at step s are one contiguous, aligned line of 256*R bytes and consecutive
//     steps are one contiguous block: it is fetched by ONE cp.async.bulk (TMA, 1-D) per array into a
//     shared-memory ring, completion signalled on an mbarrier; shared-memory accesses are conflict-free;
//   * the backward sweep walks the same storage in reverse step order, so one layout serves both solves;
//     BLAS-1 kernels are layout-agnostic (padding slots hold zeros and stay zero), and the 5-point stencil
//     finds its neighbours at fixed offsets (applyASdKernel).
//
// Kernels, oldest first (all validated bit for bit by tools/wavebench.cu):
//   waveKernel    one warp per strip does everything (R = 1); kept for the harness only.
//   solveKernel   R = 1, SIGMA >= 2: warp-specialised (solver / TMA + hand-off in / write-back + hand-off out);
//                 the neighbour row's value is read back from the tile two steps after it was stored.
//   solveKernelR  R = 2, SIGMA = 1: the production kernel; the neighbour lane's value comes by shuffle, two
//                 dependent DFMA per step run down the lane's rows.
// Strip k+1 receives the last row of strip k through distributed shared memory inside a thread-block cluster
// (st.async + mbarrier) and through self-validating global slots between clusters (a reserved NaN payload =
// "not written yet").  Strips are claimed through an atomic ticket in march order, so the producer of a strip
// is always resident or finished (no deadlock at any grid size).  Control::range restricts every strip to the
// chunks that hold fluid.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace sd {

// Ops of solveKernelR may ask the pre warp for the PCG's residual update (Op::PRE_AXPY, see solveKernelR)
template <class Op, class = void>
struct OpPreAxpy { static constexpr bool value = false; };
template <class Op>
struct OpPreAxpy<Op, std::void_t<decltype(Op::PRE_AXPY)>> { static constexpr bool value = Op::PRE_AXPY; };
// ... and the post warp for copies of the first / last row of `out` in peer memory (Op::HALO: y-slab halo rows)
template <class Op, class = void>
struct OpHalo { static constexpr bool value = false; };
template <class Op>
struct OpHalo<Op, std::void_t<decltype(Op::HALO)>> { static constexpr bool value = Op::HALO; };

constexpr unsigned long long SENT = 0x7FF8F51D0DEAD001ULL;  // reserved quiet-NaN payload: "not written yet"
constexpr int CH = 32;                                       // steps per chunk (one 8 KB TMA block per array)
constexpr int SUB = 8;                                       // steps per hand-off poll

// rpl = rows per lane (R): a strip is 32*R rows, lane t owns rows R*t .. R*t+R-1 of it and visits them at the same
// column in one step (see solveKernelR); element (row 32R*k + R*t + r, column c) lives at
//     [((k*Sp + c + sigma*t) * 32 + t) * R + r]
struct Geom {
    int nx, ny;        // logical columns / rows of the arrays
    int nstrips;       // ceil(ny / (32 * rpl))
    int Sp;            // steps per strip, multiple of CH
    int nchunks;       // Sp / CH
    int sigma;
    int rpl;
    size_t elems;      // nstrips * Sp * 32 * rpl
};

static inline Geom makeGeom(int nx, int ny, int sigma, int rpl = 1) {
    Geom g;
    g.nx = nx; g.ny = ny; g.sigma = sigma; g.rpl = rpl;
    g.nstrips = (ny + 32 * rpl - 1) / (32 * rpl);
    g.Sp = ((nx + 31 * sigma + CH - 1) / CH) * CH;
    g.nchunks = g.Sp / CH;
    g.elems = (size_t)g.nstrips * g.Sp * 32 * rpl;
    return g;
}

__host__ __device__ __forceinline__ size_t sdIndex(const Geom& g, int i, int j) {
    const int rows = 32 * g.rpl;
    int k = j / rows, q = j - k * rows, t = q / g.rpl, r = q - t * g.rpl;
    return (((size_t)k * g.Sp + (size_t)(i + g.sigma * t)) * 32 + t) * g.rpl + r;
}

struct Control {
    int* ticket;               // zero between launches
    int* finished;             // zero between launches
    unsigned long long* hand;  // [nstrips + 1][handStride(g)]; polled slots are SENT between launches
    const int* gate;           // optional: the kernel is a no-op when *gate != 0
    long long* prof;           // optional (SD_PROFILE builds): [4 * nstrips] total / TMA-wait / poll-wait cycles, start clock
    const int* range;          // optional: [2 * nstrips] first / last storage chunk of each strip that holds anything
                               // non-zero (first > last: nothing).  solveKernel only marches those chunks; everything
                               // outside is exactly zero in every input and must be zero in the output array already.
    int dbg;                   // timing experiments (FSIM_DBG_PRE), 0 in production
};

// hand-off regions: one per producing strip plus a dummy one that absorbs the last strip's writes.  A slot is
// addressed by column + 31*sigma, so the producer can store unconditionally at every step.
static inline size_t handStride(const Geom& g) { return (size_t)g.Sp + 31 * g.sigma + 33; }  // (one row per strip)
static inline size_t handWords(const Geom& g) { return handStride(g) * (size_t)(g.nstrips + 1); }

#ifdef __CUDACC__

__device__ __forceinline__ unsigned long long ldRelaxedU64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stRelaxedU64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int smemAddr(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, unsigned int parity) {
    unsigned int ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smemAddr(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D bulk copy global -> shared (TMA), completion counted in bytes on `bar`
__device__ __forceinline__ void bulkLoad(void* smemDst, const void* gmemSrc, unsigned int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smemAddr(smemDst)),
                 "l"(gmemSrc), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}

template <class Op>
struct Layout {
    static constexpr int STAGE_DOUBLES = Op::NIN * CH * 32;
    // as many stages as fit in ~200 KB: the ring has to cover the HBM latency at one warp's consumption rate
    static constexpr int NST = (200 * 1024) / (STAGE_DOUBLES * 8) > 12 ? 12 : (200 * 1024) / (STAGE_DOUBLES * 8);
    static constexpr size_t BYTES = (size_t)NST * STAGE_DOUBLES * 8 + 3 * CH * 8 + NST * 8 + 64;
};

// DIR = +1: steps ascending, value of lane t-1 flows to lane t (forward solve); DIR = -1: steps descending,
// lane t+1 -> lane t (backward solve).  Op supplies
//   NIN, NOUT; const double* in[NIN]; double* out[NOUT];
//   template <int SIGMA> void cell(const double (&v)[NIN], double left, double down, double (&o)[NOUT], double& y, double& acc)
//   void stripDone(int strip, double acc); void allDone(int nstrips)
//
// Software pipeline of the one warp: the inputs of the next SUB steps are read from shared memory into
// registers while the current SUB steps run from registers, so the loop-carried chain is one DFMA per step
// (plus the shuffle when SIGMA = 1) and shared-memory latency never sits on it.  Hand-off slots are polled a
// whole chunk (32 columns) at a time, two chunks ahead of their use, which hides the L2 round trip.
template <class Op, int SIGMA, int DIR>
__global__ void __launch_bounds__(32, 1) waveKernel(Op op, Geom g, Control ctl) {
    using L = Layout<Op>;
    constexpr int NIN = Op::NIN, NOUT = Op::NOUT, NST = L::NST, NSUB = CH / SUB;
    constexpr int LC = DIR > 0 ? 0 : 31;   // lane that consumes the neighbouring strip's values
    constexpr int LP = DIR > 0 ? 31 : 0;   // lane that produces them for the next strip
    extern __shared__ __align__(128) unsigned char smemRaw[];
    double* tile = reinterpret_cast<double*>(smemRaw);
    double* handbuf = tile + (size_t)NST * L::STAGE_DOUBLES;  // [3][CH] ring of polled hand-off values
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(handbuf + 3 * CH);

    if (ctl.gate && *ctl.gate != 0) return;
    const int lane = threadIdx.x;
    int q = 0;
    if (lane == 0) q = atomicAdd(ctl.ticket, 1);
    q = __shfl_sync(0xffffffffu, q, 0);
    const int k = DIR > 0 ? q : g.nstrips - 1 - q;
    const bool hasProducer = q > 0;
    const size_t stripBase = (size_t)k * g.Sp * 32;
    const size_t hstride = (size_t)g.Sp + 31 * SIGMA + 33;
    // slot of column c is c + 31*SIGMA; producer lane LP is at column s - SIGMA*LP, consumer lane LC at s - SIGMA*LC
    unsigned long long* handOut = ctl.hand + (size_t)(q < g.nstrips - 1 ? q : g.nstrips) * hstride + (31 - LP) * SIGMA;
    unsigned long long* handIn = ctl.hand + (size_t)(q > 0 ? q - 1 : 0) * hstride + (31 - LC) * SIGMA;
    const int nchunks = g.nchunks;

    if (lane == 0) {
        for (int st = 0; st < NST; ++st) mbarInit(&bars[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = lane; i < 3 * CH; i += 32) handbuf[i] = 0.0;
    __syncwarp();

    auto chunkOf = [&](int n) { return DIR > 0 ? n : nchunks - 1 - n; };
    auto issueLoad = [&](int n) {  // n = chunk number in march order
        const int cn = chunkOf(n), st = n % NST;
        mbarExpectTx(&bars[st], NIN * CH * 32 * 8);
#pragma unroll
        for (int a = 0; a < NIN; ++a)
            bulkLoad(tile + ((size_t)st * NIN + a) * (CH * 32), op.in[a] + stripBase + (size_t)cn * (CH * 32), CH * 32 * 8,
                     &bars[st]);
    };
    if (lane == 0)
        for (int n = 0; n < NST - 1 && n < nchunks; ++n) issueLoad(n);

    // hand-off polling, one chunk per poll: lane l owns the slot of consumer step 32*cn + l
    auto pollNeeded = [&](int n) {
        const int c = chunkOf(n) * CH + lane - SIGMA * LC;
        return hasProducer && n < nchunks && c >= 0 && c < g.nx;
    };
    auto pollIssue = [&](int n) -> unsigned long long {
        return pollNeeded(n) ? ldRelaxedU64(handIn + chunkOf(n) * CH + lane) : SENT;
    };
    auto pollResolve = [&](int n, unsigned long long hv) {  // -> handbuf[(n % 3) * CH + lane]
        if (!hasProducer || n >= nchunks) return;
        const bool need = pollNeeded(n);
        unsigned long long* slot = handIn + chunkOf(n) * CH + lane;
        while (true) {
            const bool ok = !need || hv != SENT;
            if (__all_sync(0xffffffffu, ok)) break;
            if (need && hv == SENT) hv = ldRelaxedU64(slot);
        }
        handbuf[(n % 3) * CH + lane] = need ? __longlong_as_double((long long)hv) : 0.0;
        if (need) stRelaxedU64(slot, SENT);  // leave the slot clean for the next launch
    };
    unsigned long long preA = pollIssue(0), preB = pollIssue(1);
    pollResolve(0, preA);
    preA = preB;           // preA now belongs to chunk 1
    preB = pollIssue(2);   // preB to chunk 2

#ifdef SD_PROFILE
    long long tStart = clock64(), tBar = 0, tPoll = 0;
#define SD_T0 long long _t0 = clock64();
#define SD_T1(acc_) acc_ += clock64() - _t0;
#else
#define SD_T0
#define SD_T1(acc_)
#endif
    double y = 0.0, acc = 0.0;
    double shq[SIGMA];
#pragma unroll
    for (int i = 0; i < SIGMA; ++i) shq[i] = 0.0;

    double vbuf[2][SUB][NIN], hbuf[2][SUB];
    auto loadSub = [&](const double* tp, const double* hb, int sub, double (&v)[SUB][NIN], double (&h)[SUB]) {
#pragma unroll
        for (int e = 0; e < SUB; ++e) {
            const int ls = DIR > 0 ? sub * SUB + e : CH - 1 - sub * SUB - e;
#pragma unroll
            for (int a = 0; a < NIN; ++a) v[e][a] = tp[(a * CH + ls) * 32 + lane];
            h[e] = hb[ls];
        }
    };
    mbarWait(&bars[0], 0);
    __syncwarp();
    loadSub(tile, handbuf, 0, vbuf[0], hbuf[0]);

    for (int n = 0; n < nchunks; ++n) {
        const int sb = chunkOf(n) * CH;
        __syncwarp();
        if (lane == 0 && n + NST - 1 < nchunks) issueLoad(n + NST - 1);  // refills the stage chunk n-1 released
        // hand-off values of chunk n+1 must be in shared memory before its first sub-chunk is pre-loaded
        { SD_T0 pollResolve(n + 1, preA); SD_T1(tPoll) }
        preA = preB;
        preB = pollIssue(n + 3);
        __syncwarp();
        const double* tp = tile + (size_t)(n % NST) * L::STAGE_DOUBLES;
        const double* hb = handbuf + (n % 3) * CH;
        double* outp[NOUT > 0 ? NOUT : 1];
#pragma unroll
        for (int a = 0; a < NOUT; ++a) outp[a] = op.out[a] + stripBase + (size_t)sb * 32 + lane;
        unsigned long long* hout = handOut + sb;
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
            // pre-load the next SUB steps
            if (sub + 1 < NSUB) {
                loadSub(tp, hb, sub + 1, vbuf[(sub + 1) & 1], hbuf[(sub + 1) & 1]);
            } else if (n + 1 < nchunks) {
                { SD_T0 mbarWait(&bars[(n + 1) % NST], (unsigned int)(((n + 1) / NST) & 1)); SD_T1(tBar) }
                loadSub(tile + (size_t)((n + 1) % NST) * L::STAGE_DOUBLES, handbuf + ((n + 1) % 3) * CH, 0, vbuf[0], hbuf[0]);
            }
            double (&v)[SUB][NIN] = vbuf[sub & 1];
            double (&h)[SUB] = hbuf[sub & 1];
#pragma unroll
            for (int e = 0; e < SUB; ++e) {
                const int ls = DIR > 0 ? sub * SUB + e : CH - 1 - sub * SUB - e;  // step inside the chunk
                double down = shq[0];
                if (lane == LC) down = h[e];
                double o[NOUT > 0 ? NOUT : 1];
                op.template cell<SIGMA>(v[e], y, down, o, y, acc);
#pragma unroll
                for (int i = 0; i + 1 < SIGMA; ++i) shq[i] = shq[i + 1];
                shq[SIGMA - 1] = DIR > 0 ? __shfl_up_sync(0xffffffffu, y, 1) : __shfl_down_sync(0xffffffffu, y, 1);
#pragma unroll
#ifndef SD_NO_OUT
                for (int a = 0; a < NOUT; ++a) outp[a][ls * 32] = o[a];
#else
                for (int a = 0; a < NOUT; ++a) acc += o[a];
#endif
#ifndef SD_NO_HAND
                if (lane == LP) stRelaxedU64(hout + ls, (unsigned long long)__double_as_longlong(y));
#endif
            }
        }
    }
#ifdef SD_PROFILE
    if (lane == 0 && ctl.prof) { ctl.prof[4 * q] = clock64() - tStart; ctl.prof[4 * q + 1] = tBar; ctl.prof[4 * q + 2] = tPoll; ctl.prof[4 * q + 3] = tStart; }
#endif
    acc = warpSum(acc);
    if (lane == 0) {
        op.stripDone(k, acc);
        __threadfence();
        int t = atomicAdd(ctl.finished, 1);
        if (t == g.nstrips - 1) {
            __threadfence();
            op.allDone(g.nstrips);
            *ctl.finished = 0;
            *ctl.ticket = 0;
            __threadfence();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Warp-specialised triangular solve (the PCG hot kernel).  One CTA of three warps per strip:
//   warp 0 "solver": the dependent chain only.  Per step: four conflict-free LDS.64 (rhs, cx, cy pre-loaded one
//          sub-chunk ahead; the neighbour row's value), two DFMA, one STS.64 (y overwrites rhs in the tile).
//          The (i, j-1) value is not shuffled: lane t-1 stored it into the tile SIGMA steps earlier, so it is one
//          more LDS at a one-lane offset (a warp's shared-memory operations are performed in order).  Lane LC,
//          whose neighbour row lives in the previous strip, reads the hand-off ring instead -- same instruction,
//          per-lane address and stride.
//   warp 1 "pre":    issues the TMA bulk loads (8 KB per array per chunk, mbarrier completion) and publishes
//          `tready` (chunks landed) and `ready` (sub-chunks whose hand-off values are in the ring).  Inside a
//          cluster the values are pushed into the ring by the producer CTA (st.async, SASS STAS) and the pre warp
//          only waits on the sub-chunk's mbarrier; at a cluster boundary it polls the global self-validating
//          slots (a reserved NaN payload = "not written yet") and copies them into the ring.
//   warp 2 "post":   follows the solver's `done` counter: publishes lane LP's values to the next strip first
//          (they are on its critical path), then out = D*y (forward; plus the partial sum of y*out that gives
//          z.r) or out = y (backward) with coalesced 256-byte stores, and releases the stage to the TMA ring.
// Progress counters live in shared memory; they and the tile are accessed with volatile/asm accesses in program
// order.  No CTA-scope fences: one SM's shared-memory pipeline performs a warp's accesses in order, and every
// consumer access is control-dependent on the counter it polled (a fence here costs ~100 cycles per sub-chunk on
// the critical path -- it was 40 % of the solver's time).
// CL = thread-block cluster size (1 = no cluster).  Requires SIGMA >= 2: cell = fma(-cx, left, fma(-cy, down, rhs)).
// ---------------------------------------------------------------------------------------------------------
constexpr int HR = 512;  // hand-off ring (march positions of the producer) in the consumer's shared memory

template <class Op>
struct SolveLayout {
    static constexpr int STAGE_DOUBLES = Op::NIN * CH * 32;
    static constexpr int NST = (200 * 1024) / (STAGE_DOUBLES * 8) > 12 ? 12 : (200 * 1024) / (STAGE_DOUBLES * 8);
    // tile ring | full[NST] | hbar[HR / 4] (enough for SUBS >= 4) | hring[HR] | counters
    static constexpr size_t BYTES = (size_t)NST * STAGE_DOUBLES * 8 + NST * 8 + (HR / 4) * 8 + HR * 8 + 64;
};

__device__ __forceinline__ unsigned int clusterCtaRank() {
    unsigned int r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned int mapaShared(unsigned int localAddr, unsigned int rank) {
    unsigned int r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(localAddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void clusterBarrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// remote 8-byte store into a cluster peer's shared memory; completion is counted in bytes on the peer's mbarrier
__device__ __forceinline__ void stAsyncU64(unsigned int remoteAddr, unsigned long long v, unsigned int remoteBar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remoteAddr), "l"(v),
                 "r"(remoteBar)
                 : "memory");
}
__device__ __forceinline__ int ldVolatileS32(const int* p) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smemAddr(p)) : "memory");
    return v;
}
__device__ __forceinline__ void stVolatileS32(int* p, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smemAddr(p)), "r"(v) : "memory");
}
#define SD_COMPILER_BARRIER() asm volatile("" ::: "memory")

// Op: NIN (3 or 4: rhs, cx, cy[, 4th array for the post warp]); KIND (0: out = y; 1: out = in[3]*y and the strip's sum
// of y*out; 2: out = y + postScalar()*in[3]); const double* in[NIN]; double* out; stripDone(strip, acc); allDone(nstrips)
template <class Op, int SIGMA, int DIR, int SUBS, int CL>
__global__ void __launch_bounds__(96, 1) solveKernel(Op op, Geom g, Control ctl) {
    static_assert(SIGMA >= 2, "the neighbour row's value must be a step old");
    static_assert((32 * SIGMA) % SUBS == 0 && HR % SUBS == 0, "ring reads of one sub-chunk must not wrap");
    using L = SolveLayout<Op>;
    constexpr int NIN = Op::NIN, NST = L::NST, NSUB = CH / SUBS;
    constexpr int LC = DIR > 0 ? 0 : 31, LP = DIR > 0 ? 31 : 0;
    constexpr int TILE = CH * 32;
    constexpr int RB = HR / SUBS;  // hand-off barriers (one per sub-chunk of the ring)
    extern __shared__ __align__(128) unsigned char smemRaw[];
    double* tile = reinterpret_cast<double*>(smemRaw);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(tile + (size_t)NST * L::STAGE_DOUBLES);
    unsigned long long* hbar = full + NST;
    double* hring = reinterpret_cast<double*>(hbar + HR / 4);
    int* cnt = reinterpret_cast<int*>(hring + HR);  // [0] ready, [1] done, [2] freed chunks, [3] ticket, [4] tready

    if (ctl.gate && *ctl.gate != 0) return;  // uniform over the grid
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int rank = CL > 1 ? clusterCtaRank() : 0u;
    for (int i = threadIdx.x; i < HR; i += 96) hring[i] = 0.0;  // the first strip never receives anything
    if (threadIdx.x == 0) {
        cnt[0] = 0; cnt[1] = 0; cnt[2] = 0; cnt[4] = 0;
        if (CL == 1) cnt[3] = atomicAdd(ctl.ticket, 1);
        for (int st = 0; st < NST; ++st) mbarInit(&full[st], 1);
        if (CL > 1)
            for (int i = 0; i < RB; ++i) mbarInit(&hbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (CL > 1) {
        __syncthreads();
        if (rank == 0 && threadIdx.x == 0) {
            // one ticket per cluster: consecutive strips for consecutive ranks, in march order
            const int q0 = atomicAdd(ctl.ticket, CL);
            for (int r = 0; r < CL; ++r)
                asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(mapaShared(smemAddr(&cnt[3]), r)), "r"(q0 + r) : "memory");
        }
        clusterBarrier();
    } else {
        __syncthreads();
    }
    const int q = cnt[3];
    if (q < g.nstrips) {  // (a cluster's trailing CTAs may have no strip)
    const int k = DIR > 0 ? q : g.nstrips - 1 - q;
    const bool hasProducer = q > 0;
    const bool dsIn = CL > 1 && rank > 0;                              // values are pushed into hring by rank - 1
    const bool dsOut = CL > 1 && rank < CL - 1 && q < g.nstrips - 1;   // values are pushed to rank + 1
    const size_t stripBase = (size_t)k * g.Sp * 32;
    const size_t hstride = (size_t)g.Sp + 31 * SIGMA + 33;
    unsigned long long* handOut = ctl.hand + (size_t)(q < g.nstrips - 1 ? q : g.nstrips) * hstride + (31 - LP) * SIGMA;
    unsigned long long* handIn = ctl.hand + (size_t)(q > 0 ? q - 1 : 0) * hstride + (31 - LC) * SIGMA;
    const int Sp = g.Sp;
    // Each strip only marches the chunks that hold something (ctl.range): march chunks [nLo, nLo + nchunks), and
    // everything below works with positions RELATIVE to base = nLo*CH.  mLo / mCnt: (first march chunk, chunk count)
    auto rangeOf = [&](int kk, int& lo, int& cntOut) {
        int sLo = 0, sHi = g.nchunks - 1;
        if (ctl.range) { sLo = ctl.range[2 * kk]; sHi = ctl.range[2 * kk + 1]; }
        if (sLo > sHi) { lo = 0; cntOut = 0; return; }
        lo = DIR > 0 ? sLo : g.nchunks - 1 - sHi;
        cntOut = sHi - sLo + 1;
    };
    int nLo, nchunks;
    rangeOf(k, nLo, nchunks);
    const int nsub = nchunks * NSUB, base = nLo * CH;
    // the strip marched before this one (its lane LP feeds our lane LC) and the one after (fed by our lane LP)
    int pLo = 0, pCnt = 0, cLo = 0, cCnt = 0;
    if (hasProducer) rangeOf(DIR > 0 ? k - 1 : k + 1, pLo, pCnt);
    if (q < g.nstrips - 1) rangeOf(DIR > 0 ? k + 1 : k - 1, cLo, cCnt);
    // Our relative position u needs the producer's absolute position base + u + 31*SIGMA; it exists iff covLo <= u < covHi
    // (otherwise the value is exactly zero).  The hand-off ring is indexed by our relative position + 31*SIGMA.
    const int covLo = pLo * CH - 31 * SIGMA - base, covHi = covLo + pCnt * CH;
    // march position (absolute) -> storage step
    auto stepOf = [&](int U) { return DIR > 0 ? U : Sp - 1 - U; };
    // bytes pushed into our sub-chunk m by the producer (it pushes exactly the covered positions we march)
    auto pushedBytes = [&](int m) {
        int a = m * SUBS, b = a + SUBS;
        if (a < covLo) a = covLo;
        if (b > covHi) b = covHi;
        return b > a ? (b - a) * 8 : 0;
    };

    if (warp == 1) {
        // ------------------------------------------------------------------------------------------ pre
        int issued = 0, landed = 0;
        auto issueLoads = [&]() {
            if (lane == 0) {
                const int lim = ldVolatileS32(&cnt[2]) + NST;
                while (issued < nchunks && issued < lim) {
                    const int cn = DIR > 0 ? nLo + issued : g.nchunks - 1 - (nLo + issued), st = issued % NST;
                    mbarExpectTx(&full[st], NIN * TILE * 8);
#pragma unroll
                    for (int a = 0; a < NIN; ++a)
                        bulkLoad(tile + ((size_t)st * NIN + a) * TILE, op.in[a] + stripBase + (size_t)cn * TILE, TILE * 8, &full[st]);
                    ++issued;
                }
            }
        };
        // chunks [0, upTo) landed -> tready
        auto land = [&](int upTo) {
            if (upTo > nchunks) upTo = nchunks;
            issueLoads();  // keep the ring full whether or not anything has to be waited for
            while (landed < upTo) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                while (true) {  // the stage may still be in use
                    issueLoads();
                    if (__shfl_sync(0xffffffffu, issued, 0) > landed) break;
#ifdef SD_PRE_SLEEP
                    __nanosleep(SD_PRE_SLEEP);  // do not hammer the shared-memory pipeline the solver lives on
#endif
                }
                mbarWait(&full[landed % NST], (unsigned int)((landed / NST) & 1));
                ++landed;
            }
            __syncwarp();
            if (lane == 0) stVolatileS32(&cnt[4], landed);
        };
        auto covered = [&](int u) { return u >= covLo && u < covHi && u < nchunks * CH; };
        auto armHandoff = [&](int m) {  // (lane 0) one arrival per phase, plus the bytes the producer will push
            const int bytes = pushedBytes(m);
            if (bytes) mbarExpectTx(&hbar[m % RB], (unsigned int)bytes);
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(&hbar[m % RB])) : "memory");
        };
        issueLoads();
        const bool glIn = hasProducer && !dsIn;
        // global path: lane l polls the slots of the relative positions u = l (mod 32); the lanes of group l / SUBS
        // serve the sub-chunks j = group.  Each group keeps its in-flight poll in its own register.
        const int grp = lane / SUBS;
        int myU = lane;
        bool need = glIn && covered(myU);
        unsigned long long hv[NSUB];
#pragma unroll
        for (int i = 0; i < NSUB; ++i) hv[i] = (need && grp == i) ? ldRelaxedU64(handIn + stepOf(base + myU)) : 0ULL;
        if (dsIn && lane == 0)
            for (int m = 0; m < RB && m < nsub; ++m) armHandoff(m);
        land(2);
        for (int n = 0; n < nchunks; ++n) {
            if (!hasProducer) {
                if (lane == 0) stVolatileS32(&cnt[0], (n + 1) * NSUB);
            } else if (dsIn) {
#pragma unroll
                for (int j = 0; j < NSUB; ++j) {
                    const int m = n * NSUB + j;
                    mbarWait(&hbar[m % RB], (unsigned int)((m / RB) & 1));
                    if (lane == 0 && m + RB < nsub) armHandoff(m + RB);  // the slot's next phase
                    // positions the producer does not march are exactly zero
                    if (lane < SUBS && !covered(m * SUBS + lane))
                        asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(smemAddr(&hring[(m * SUBS + lane + 31 * SIGMA) & (HR - 1)])), "d"(0.0) : "memory");
                    __syncwarp();
                    if (lane == 0) stVolatileS32(&cnt[0], m + 1);
                }
            } else {
#pragma unroll
                for (int j = 0; j < NSUB; ++j) {
                    const bool mine = grp == j;
                    while (true) {
                        const bool valid = !mine || !need || hv[j] != SENT;
                        if (__all_sync(0xffffffffu, valid)) break;
                        if (!valid) hv[j] = ldRelaxedU64(handIn + stepOf(base + myU));
                    }
                    if (mine) {
                        // (ring entries of sub-chunk m - RB are long consumed: the TMA ring keeps this warp within
                        // NST chunks of the solver)
                        double h = 0.0;
                        if (need) {
                            h = __longlong_as_double((long long)hv[j]);
                            stRelaxedU64(handIn + stepOf(base + myU), SENT);  // leave the slot clean for the next launch
                        }
                        asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(smemAddr(&hring[(myU + 31 * SIGMA) & (HR - 1)])), "d"(h) : "memory");
                        myU += 32;
                        need = covered(myU);
                        hv[j] = need ? ldRelaxedU64(handIn + stepOf(base + myU)) : 0ULL;
                    }
                    __syncwarp();
                    if (lane == 0) stVolatileS32(&cnt[0], n * NSUB + j + 1);
                }
            }
            land(n + 3);  // the solver pre-loads one sub-chunk ahead: keep two chunks landed beyond the current one
        }
    } else if (warp == 0 && nsub > 0) {
        // ------------------------------------------------------------------------------------------ solver
        const unsigned tileA = smemAddr(tile) + lane * 8;
        const unsigned ringA = smemAddr(hring);
        const unsigned readyA = smemAddr(&cnt[0]), doneA = smemAddr(&cnt[1]), treadyA = smemAddr(&cnt[4]);
        constexpr int STAGE_BYTES = L::STAGE_DOUBLES * 8, TILE_BYTES = TILE * 8;
        constexpr int STEP = DIR > 0 ? 256 : -256;
        double y = 0.0;
        double va[2][SUBS], vx[2][SUBS], vy[2][SUBS];
        auto subAddr = [&](int m) -> unsigned {
            const int n = m / NSUB, j = m - n * NSUB;
            return tileA + (unsigned)((n % NST) * STAGE_BYTES + (DIR > 0 ? j * SUBS : CH - 1 - j * SUBS) * 256);
        };
        auto waitCnt = [&](unsigned addr, int need) {
            int v;
            do { asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); } while (v < need);
        };
        auto loadSub = [&](unsigned addr, double (&a)[SUBS], double (&x)[SUBS], double (&yy)[SUBS]) {
#pragma unroll
            for (int e = 0; e < SUBS; ++e) {
                const unsigned p = addr + (unsigned)(e * STEP);
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[e]) : "r"(p) : "memory");
                asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(x[e]) : "r"(p), "n"(TILE_BYTES) : "memory");
                asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(yy[e]) : "r"(p), "n"(2 * TILE_BYTES) : "memory");
            }
        };
        // neighbour-row value of step e (march position m*SUBS + e): lanes other than LC read the tile slot lane
        // t -+ 1 wrote SIGMA steps earlier; lane LC reads ring entry (m*SUBS + e + 31*SIGMA) & (HR-1).
        const bool isLC = lane == LC;
        const int dnStride = isLC ? 8 : STEP;
        const int dnLaneOff = DIR > 0 ? -8 : 8;
#ifdef SD_PROFILE
        long long tStart = clock64(), tWaitR = 0, tWaitT = 0;
#endif
        waitCnt(treadyA, 1);
        SD_COMPILER_BARRIER();
        loadSub(subAddr(0), va[0], vx[0], vy[0]);
        double dn[SUBS + SIGMA];
#pragma unroll
        for (int i = 0; i < SIGMA; ++i) dn[SUBS + i] = 0.0;
#pragma unroll 1
        for (int m2 = 0; m2 < nsub; m2 += 2) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = m2 + h;  // nsub is even
                const unsigned cur = subAddr(m);
                unsigned nxt = cur;  // past the end the pre-load re-reads this sub-chunk (harmless)
                if (m + 1 < nsub) {
                    nxt = subAddr(m + 1);
                    if (((m + 1) % NSUB) == 0) {
                        SD_T0 waitCnt(treadyA, (m + 1) / NSUB + 1); SD_T1(tWaitT)  // next chunk landed (rarely waits)
                    }
                }
                { SD_T0 waitCnt(readyA, m + 1); SD_T1(tWaitR) }  // hand-off values of this sub-chunk are in the ring
                SD_COMPILER_BARRIER();
                // The first SIGMA steps look back: the other lanes carry the previous sub-chunk's last values (loaded
                // before `done` moved, i.e. before the stage could be recycled); lane LC reads the ring now.
#pragma unroll
                for (int i = 0; i < SIGMA; ++i) {
                    double rv;
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(rv) : "r"(ringA + (unsigned)(((m * SUBS + 31 * SIGMA + i) & (HR - 1)) * 8)) : "memory");
                    dn[i] = isLC ? rv : dn[SUBS + i];
                }
                unsigned dnA = isLC ? ringA + (unsigned)(((m * SUBS + 32 * SIGMA) & (HR - 1)) * 8) : cur + (unsigned)dnLaneOff;
#pragma unroll
                for (int e = 0; e < SUBS; ++e) {
                    y = __fma_rn(-vx[h][e], y, __fma_rn(-vy[h][e], dn[e], va[h][e]));
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(cur + (unsigned)(e * STEP)), "d"(y) : "memory");
                    // (for lane LC the last SIGMA of these read ring entries of the next sub-chunk, which may not be
                    // there yet: they are replaced after the wait above)
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(dn[e + SIGMA]) : "r"(dnA) : "memory");
                    dnA += (unsigned)dnStride;
                    // the next sub-chunk's inputs are fetched between the chain's instructions, where the
                    // single warp would otherwise idle on the DFMA / LDS latencies
                    const unsigned p = nxt + (unsigned)(e * STEP);
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(va[h ^ 1][e]) : "r"(p) : "memory");
                    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(vx[h ^ 1][e]) : "r"(p), "n"(TILE_BYTES) : "memory");
                    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(vy[h ^ 1][e]) : "r"(p), "n"(2 * TILE_BYTES) : "memory");
                }
                __syncwarp();
                SD_COMPILER_BARRIER();  // the y stores precede `done` in program order (same in-order pipeline)
                if (lane == 0) asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(doneA), "r"(m + 1) : "memory");
            }
        }
#ifdef SD_PROFILE
        if (lane == 0 && ctl.prof) { ctl.prof[4 * q] = clock64() - tStart; ctl.prof[4 * q + 1] = tWaitT; ctl.prof[4 * q + 2] = tWaitR; ctl.prof[4 * q + 3] = tStart; }
#endif
    } else if (warp == 2) {
        // ------------------------------------------------------------------------------------------ post
        double acc = 0.0;
        const double postScalar = op.postScalar();
        int peerReady = 0;  // the consumer's `done` as last read (back-pressure of the in-cluster ring)
        const unsigned int peerRing = dsOut ? mapaShared(smemAddr(hring), rank + 1) : 0u;
        const unsigned int peerBar = dsOut ? mapaShared(smemAddr(hbar), rank + 1) : 0u;
        const unsigned int peerCnt = dsOut ? mapaShared(smemAddr(&cnt[1]), rank + 1) : 0u;  // the consumer's `done`
        for (int m = 0; m < nsub; ++m) {
            const int n = m / NSUB, j = m - n * NSUB;
            while (ldVolatileS32(&cnt[1]) < m + 1) {
#ifdef SD_SPIN_SLEEP
                __nanosleep(SD_SPIN_SLEEP);
#endif
            }
            SD_COMPILER_BARRIER();
            const volatile double* tp = tile + (size_t)(n % NST) * L::STAGE_DOUBLES;
            const int cn = DIR > 0 ? nLo + n : g.nchunks - 1 - (nLo + n);  // storage chunk
            // last row first: it is on the next strip's critical path.  Lane l's value (our relative position
            // m*SUBS + l) feeds the consumer's relative position uc; only positions the consumer marches are sent.
            const int uc = base + m * SUBS + lane - 31 * SIGMA - cLo * CH;
            const bool feeds = lane < SUBS && uc >= 0 && uc < cCnt * CH;
            if (dsOut) {
                if (__any_sync(0xffffffffu, feeds)) {
                    // the ring slot of consumer position uc was last used by uc - HR: the consumer's SOLVER must be done
                    // with that sub-chunk (its pre warp may run up to a TMA ring ahead of it)
                    int ucMax = base + m * SUBS + SUBS - 1 - 31 * SIGMA - cLo * CH;
                    if (ucMax > cCnt * CH - 1) ucMax = cCnt * CH - 1;
                    const int mC = ucMax / SUBS;
                    while (peerReady < mC - RB + 1)
                        asm volatile("ld.volatile.shared::cluster.s32 %0, [%1];" : "=r"(peerReady) : "r"(peerCnt) : "memory");
                }
                if (feeds) {
                    const int ls = DIR > 0 ? j * SUBS + lane : CH - 1 - j * SUBS - lane;
                    const double yv = tp[ls * 32 + LP];
                    stAsyncU64(peerRing + (unsigned)(((uc + 31 * SIGMA) & (HR - 1)) * 8),
                               (unsigned long long)__double_as_longlong(yv), peerBar + (unsigned)(((uc / SUBS) % RB) * 8));
                }
            } else if (feeds) {
                const int ls = DIR > 0 ? j * SUBS + lane : CH - 1 - j * SUBS - lane;
                const double yv = tp[ls * 32 + LP];
                stRelaxedU64(handOut + stepOf(base + m * SUBS + lane), (unsigned long long)__double_as_longlong(yv));
            }
            double* outp = op.out + stripBase + (size_t)cn * TILE + lane;
#pragma unroll
            for (int e = 0; e < SUBS; ++e) {
                const int ls = DIR > 0 ? j * SUBS + e : CH - 1 - j * SUBS - e;
                const double yv = tp[ls * 32 + lane];
                if (Op::KIND == 1) {         // forward: out = D*y, partial sum of y*out
                    const double w = tp[3 * TILE + ls * 32 + lane] * yv;
                    acc = __fma_rn(yv, w, acc);
                    outp[ls * 32] = w;
                } else if (Op::KIND == 2) {  // backward fused with the direction update: out = y + beta*out_old
                    outp[ls * 32] = __fma_rn(postScalar, tp[3 * TILE + ls * 32 + lane], yv);
                } else {
                    outp[ls * 32] = yv;
                }
            }
            if (j == NSUB - 1) {
                __syncwarp();
                if (lane == 0) stVolatileS32(&cnt[2], n + 1);
            }
        }
        acc = warpSum(acc);
        if (lane == 0) op.stripDone(k, acc);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        int t = atomicAdd(ctl.finished, 1);
        if (t == g.nstrips - 1) {
            __threadfence();
            op.allDone(g.nstrips);
            *ctl.finished = 0;
            *ctl.ticket = 0;
            __threadfence();
        }
    }
    }  // q < nstrips
    // a peer's shared memory must stay alive until every push into it and every read of its counters is over
    if (CL > 1) clusterBarrier();
}

// ---------------------------------------------------------------------------------------------------------
// The same solve with R rows per lane (Geom::rpl = R): a strip is 32*R rows, lane t visits its R rows at the same
// column in one step -- R dependent DFMA down the rows (~20 cycles each on sm_100), the neighbour lane's last-row value
// comes by shuffle from SIGMA steps earlier (off the chain for SIGMA = 2).  R times fewer strips to chain and half the
// cycles per row and step: the pipeline lag that bounds the solve shrinks accordingly.
// Cell formula: y = fma(-cy, down, fma(-cx, left, rhs)).
// STATUS (round 1): validated bit-exactly by tools/wavebench.cu (random per-strip ranges, clusters of 1/4/8) and
// measured at 4096^2, R = 2, SIGMA = 1: 0.42 / 0.35 ms forward / backward against 0.48 / 0.47 ms for solveKernel; the
// step costs ~80 cycles for two rows where ~50 were expected, and three 64 KB TMA stages are tight.  Not wired into
// the projection yet: the PCG kernels (applyA, axpy, pack, halo rows) still assume one row per lane.
// ---------------------------------------------------------------------------------------------------------
// CHK = steps per TMA block / ring stage of this kernel (the layout's chunks of CH = 32 steps are only the unit of
// Control::range)
template <class Op, int R, int CHK>
struct SolveLayoutR {
    static constexpr int STAGE_DOUBLES = Op::NIN * CHK * 32 * R;
    static constexpr int NST = (200 * 1024) / (STAGE_DOUBLES * 8) > 12 ? 12 : (200 * 1024) / (STAGE_DOUBLES * 8);
    static_assert(NST >= 3, "the TMA ring needs three stages");
    static constexpr size_t BYTES = (size_t)NST * STAGE_DOUBLES * 8 + NST * 8 + (HR / 4) * 8 + HR * 8 + 64;
};

// Op: NIN (3 or 4: rhs, cx, cy[, 4th array for the post warp]); KIND (0: out = y; 1: out = in[3]*y and the strip's sum
// of y*out; 2: out = y + postScalar()*in[3]); const double* in[NIN]; double* out; stripDone(strip, acc); allDone(nstrips)
template <class Op, int R, int SIGMA, int DIR, int SUBS, int CL, int CHK = (R == 2 ? 16 : 8)>
__global__ void __launch_bounds__(96, 1) solveKernelR(Op op, Geom g, Control ctl) {
    static_assert(CH % CHK == 0 && CHK % SUBS == 0, "");
    static_assert(HR % SUBS == 0 && (R == 2 || R == 4), "");
    using L = SolveLayoutR<Op, R, CHK>;
    constexpr int NIN = Op::NIN, NST = L::NST, NSUB = CHK / SUBS;
    constexpr int LC = DIR > 0 ? 0 : 31, LP = DIR > 0 ? 31 : 0;
    constexpr int TILE = CHK * 32 * R;
    constexpr int RB = HR / SUBS;  // hand-off barriers (one per sub-chunk of the ring)
    extern __shared__ __align__(128) unsigned char smemRaw[];
    double* tile = reinterpret_cast<double*>(smemRaw);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(tile + (size_t)NST * L::STAGE_DOUBLES);
    unsigned long long* hbar = full + NST;
    double* hring = reinterpret_cast<double*>(hbar + HR / 4);
    int* cnt = reinterpret_cast<int*>(hring + HR);  // [0] ready, [1] done, [2] freed chunks, [3] ticket, [4] tready
    double* preRed = reinterpret_cast<double*>(cnt + 8);  // the pre warp's |r|_inf of this strip (PRE_AXPY)
    constexpr bool PRE = OpPreAxpy<Op>::value;

    if (ctl.gate && *ctl.gate != 0) return;  // uniform over the grid
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int rank = CL > 1 ? clusterCtaRank() : 0u;
    for (int i = threadIdx.x; i < HR; i += 96) hring[i] = 0.0;  // the first strip never receives anything
    if (threadIdx.x == 0) {
        cnt[0] = 0; cnt[1] = 0; cnt[2] = 0; cnt[4] = 0;
        if (CL == 1) cnt[3] = atomicAdd(ctl.ticket, 1);
        for (int st = 0; st < NST; ++st) mbarInit(&full[st], 1);
        if (CL > 1)
            for (int i = 0; i < RB; ++i) mbarInit(&hbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (CL > 1) {
        __syncthreads();
        if (rank == 0 && threadIdx.x == 0) {
            // one ticket per cluster: consecutive strips for consecutive ranks, in march order
            const int q0 = atomicAdd(ctl.ticket, CL);
            for (int r = 0; r < CL; ++r)
                asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(mapaShared(smemAddr(&cnt[3]), r)), "r"(q0 + r) : "memory");
        }
        clusterBarrier();
    } else {
        __syncthreads();
    }
    const int q = cnt[3];
    if (q < g.nstrips) {  // (a cluster's trailing CTAs may have no strip)
    const int k = DIR > 0 ? q : g.nstrips - 1 - q;
    const bool hasProducer = q > 0;
    const bool dsIn = CL > 1 && rank > 0;                              // values are pushed into hring by rank - 1
    const bool dsOut = CL > 1 && rank < CL - 1 && q < g.nstrips - 1;   // values are pushed to rank + 1
    const size_t stripBase = (size_t)k * g.Sp * 32 * R;
    const size_t hstride = (size_t)g.Sp + 31 * SIGMA + 33;
    unsigned long long* handOut = ctl.hand + (size_t)(q < g.nstrips - 1 ? q : g.nstrips) * hstride + (31 - LP) * SIGMA;
    unsigned long long* handIn = ctl.hand + (size_t)(q > 0 ? q - 1 : 0) * hstride + (31 - LC) * SIGMA;
    const int Sp = g.Sp;
    // Each strip only marches the chunks that hold something (ctl.range): march chunks [nLo, nLo + nchunks), and
    // everything below works with positions RELATIVE to base = nLo*CHK.  mLo / mCnt: (first march chunk, chunk count)
    auto rangeOf = [&](int kk, int& lo, int& cntOut) {
        constexpr int F = CH / CHK;  // kernel chunks per layout chunk
        const int nk = g.nchunks * F;
        int sLo = 0, sHi = nk - 1;
        if (ctl.range) { sLo = ctl.range[2 * kk] * F; sHi = ctl.range[2 * kk + 1] * F + F - 1; }
        if (sLo > sHi) { lo = 0; cntOut = 0; return; }
        lo = DIR > 0 ? sLo : nk - 1 - sHi;
        cntOut = sHi - sLo + 1;
    };
    int nLo, nchunks;
    rangeOf(k, nLo, nchunks);
    const int nsub = nchunks * NSUB, base = nLo * CHK;
    // the strip marched before this one (its lane LP feeds our lane LC) and the one after (fed by our lane LP)
    int pLo = 0, pCnt = 0, cLo = 0, cCnt = 0;
    if (hasProducer) rangeOf(DIR > 0 ? k - 1 : k + 1, pLo, pCnt);
    if (q < g.nstrips - 1) rangeOf(DIR > 0 ? k + 1 : k - 1, cLo, cCnt);
    // Our relative position u needs the producer's absolute position base + u + 31*SIGMA; it exists iff covLo <= u < covHi
    // (otherwise the value is exactly zero).  The hand-off ring is indexed by our relative position + 31*SIGMA.
    const int covLo = pLo * CHK - 31 * SIGMA - base, covHi = covLo + pCnt * CHK;
    // march position (absolute) -> storage step
    auto stepOf = [&](int U) { return DIR > 0 ? U : Sp - 1 - U; };
    // bytes pushed into our sub-chunk m by the producer (it pushes exactly the covered positions we march)
    auto pushedBytes = [&](int m) {
        int a = m * SUBS, b = a + SUBS;
        if (a < covLo) a = covLo;
        if (b > covHi) b = covHi;
        return b > a ? (b - a) * 8 : 0;
    };

    if (warp == 1) {
        // ------------------------------------------------------------------------------------------ pre
        int issued = 0, landed = 0;
        double preAlpha = 0.0, preMax = 0.0;
        bool preStore = false;
        if constexpr (PRE) { preAlpha = op.preAlpha(); preStore = op.preStore(); }
        const int preDbg = ctl.dbg;  // timing experiments: 1 = skip the residual update altogether, 2 = no write-back
        auto issueLoads = [&]() {
            if (lane == 0) {
                const int lim = ldVolatileS32(&cnt[2]) + NST;
                while (issued < nchunks && issued < lim) {
                    const int cn = DIR > 0 ? nLo + issued : g.nchunks * (CH / CHK) - 1 - (nLo + issued), st = issued % NST;
                    mbarExpectTx(&full[st], NIN * TILE * 8);
#pragma unroll
                    for (int a = 0; a < NIN; ++a)
                        bulkLoad(tile + ((size_t)st * NIN + a) * TILE, op.in[a] + stripBase + (size_t)cn * TILE, TILE * 8, &full[st]);
                    ++issued;
                }
            }
        };
        // chunks [0, upTo) landed -> tready
        auto land = [&](int upTo) {
            if (upTo > nchunks) upTo = nchunks;
            issueLoads();  // keep the ring full whether or not anything has to be waited for
            while (landed < upTo) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                while (true) {  // the stage may still be in use
                    issueLoads();
                    if (__shfl_sync(0xffffffffu, issued, 0) > landed) break;
#ifdef SD_PRE_SLEEP
                    __nanosleep(SD_PRE_SLEEP);  // do not hammer the shared-memory pipeline the solver lives on
#endif
                }
                mbarWait(&full[landed % NST], (unsigned int)((landed / NST) & 1));
                if constexpr (PRE) if (!(preDbg & 1)) {
                    // the PCG's residual update r -= alpha z (reference src/FluidSim2D.cpp:452) on the chunk that has just
                    // landed, before the solver sees it: array 0 of the stage is r, the last one z = A s.  The new r goes
                    // back to global memory from here (the solver overwrites its tile slot with the solution) and its
                    // largest magnitude is this strip's share of |r|_inf (:453).
                    const unsigned trA = smemAddr(tile + (size_t)(landed % NST) * L::STAGE_DOUBLES) + (unsigned)(lane * 16);
                    const unsigned tzA = trA + (unsigned)((NIN - 1) * TILE * 8);
                    const int cnk = DIR > 0 ? nLo + landed : g.nchunks * (CH / CHK) - 1 - (nLo + landed);
                    double* gr = op.rOut + stripBase + (size_t)cnk * TILE + lane * 2;
                    // all loads of a batch first, then the arithmetic and the stores: the one warp keeps 2 x PB shared-memory
                    // loads in flight instead of paying a load latency per element
                    constexpr int PB = 8, NV = TILE / 64;  // vectors of two doubles per lane and batch / per lane and chunk
                    static_assert(NV % PB == 0, "");
#pragma unroll 1
                    for (int b0 = 0; b0 < NV; b0 += PB) {
                        double rx[PB], ry[PB], zx[PB], zy[PB];
#pragma unroll
                        for (int q = 0; q < PB; ++q) {
                            const unsigned o = (unsigned)((b0 + q) * 512);
                            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rx[q]), "=d"(ry[q]) : "r"(trA + o) : "memory");
                            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(zx[q]), "=d"(zy[q]) : "r"(tzA + o) : "memory");
                        }
#pragma unroll
                        for (int q = 0; q < PB; ++q) {
                            const unsigned o = (unsigned)((b0 + q) * 512);
                            rx[q] = __fma_rn(-preAlpha, zx[q], rx[q]);
                            ry[q] = __fma_rn(-preAlpha, zy[q], ry[q]);
                            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(trA + o), "d"(rx[q]), "d"(ry[q]) : "memory");
                            if (preStore && !(preDbg & 2)) *reinterpret_cast<double2*>(gr + (size_t)(b0 + q) * 64) = make_double2(rx[q], ry[q]);
                            preMax = fmax(preMax, fmax(fabs(rx[q]), fabs(ry[q])));
                        }
                    }
                    SD_COMPILER_BARRIER();  // the tile stores precede `tready` in program order
                }
                ++landed;
            }
            __syncwarp();
            if (lane == 0) stVolatileS32(&cnt[4], landed);
        };
        auto covered = [&](int u) { return u >= covLo && u < covHi && u < nchunks * CHK; };
        auto armHandoff = [&](int m) {  // (lane 0) one arrival per phase, plus the bytes the producer will push
            const int bytes = pushedBytes(m);
            if (bytes) mbarExpectTx(&hbar[m % RB], (unsigned int)bytes);
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(&hbar[m % RB])) : "memory");
        };
        issueLoads();
        const bool glIn = hasProducer && !dsIn;
        // global path: lane l polls the slots of the relative positions u = l (mod 32); the lanes of group l / SUBS
        // serve the sub-chunks j = group.  Each group keeps its in-flight poll in its own register.
        const int grp = lane / SUBS;
        int myU = lane;
        bool need = glIn && lane < CHK && covered(myU);  // (lanes >= CHK have no position in a chunk)
        unsigned long long hv[NSUB];
#pragma unroll
        for (int i = 0; i < NSUB; ++i) hv[i] = (need && grp == i) ? ldRelaxedU64(handIn + stepOf(base + myU)) : 0ULL;
        if (dsIn && lane == 0)
            for (int m = 0; m < RB && m < nsub; ++m) armHandoff(m);
        land(2);
        for (int n = 0; n < nchunks; ++n) {
            if (!hasProducer) {
                if (lane == 0) stVolatileS32(&cnt[0], (n + 1) * NSUB);
            } else if (dsIn) {
#pragma unroll
                for (int j = 0; j < NSUB; ++j) {
                    const int m = n * NSUB + j;
                    mbarWait(&hbar[m % RB], (unsigned int)((m / RB) & 1));
                    if (lane == 0 && m + RB < nsub) armHandoff(m + RB);  // the slot's next phase
                    // positions the producer does not march are exactly zero
                    if (lane < SUBS && !covered(m * SUBS + lane))
                        asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(smemAddr(&hring[(m * SUBS + lane + 31 * SIGMA) & (HR - 1)])), "d"(0.0) : "memory");
                    __syncwarp();
                    if (lane == 0) stVolatileS32(&cnt[0], m + 1);
                }
            } else {
#pragma unroll
                for (int j = 0; j < NSUB; ++j) {
                    const bool mine = grp == j;
                    while (true) {
                        const bool valid = !mine || !need || hv[j] != SENT;
                        if (__all_sync(0xffffffffu, valid)) break;
                        if (!valid) hv[j] = ldRelaxedU64(handIn + stepOf(base + myU));
                    }
                    if (mine) {
                        // (ring entries of sub-chunk m - RB are long consumed: the TMA ring keeps this warp within
                        // NST chunks of the solver)
                        double h = 0.0;
                        if (need) {
                            h = __longlong_as_double((long long)hv[j]);
                            stRelaxedU64(handIn + stepOf(base + myU), SENT);  // leave the slot clean for the next launch
                        }
                        asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(smemAddr(&hring[(myU + 31 * SIGMA) & (HR - 1)])), "d"(h) : "memory");
                        myU += CHK;
                        need = lane < CHK && covered(myU);
                        hv[j] = need ? ldRelaxedU64(handIn + stepOf(base + myU)) : 0ULL;
                    }
                    __syncwarp();
                    if (lane == 0) stVolatileS32(&cnt[0], n * NSUB + j + 1);
                }
            }
            land(n + 3);  // the solver pre-loads one sub-chunk ahead: keep two chunks landed beyond the current one
        }
        if constexpr (PRE) {
            preMax = warpMax(preMax);
            if (lane == 0) *preRed = preMax;
        }
    } else if (warp == 0 && nsub > 0) {
        // ------------------------------------------------------------------------------------------ solver
        // Lane t owns R rows and visits them at the same column in one step, one column behind lane t-1.  Per step: the
        // neighbour lane's last-row result of the previous step arrives by shuffle (lane LC: from the hand-off ring),
        // then R dependent DFMA down the lane's rows (the (i-1,j) terms are folded in beforehand, off the chain), one
        // vector store.  Inputs are fetched PFD steps ahead into a small register ring.
        constexpr int STEPB = 32 * R * 8;  // bytes per step of one array in the tile
        constexpr int STEP = DIR > 0 ? STEPB : -STEPB;
        constexpr int STAGE_BYTES = L::STAGE_DOUBLES * 8, TILE_BYTES = TILE * 8;
        constexpr int PFD = 2, QN = 4;     // prefetch distance, register ring size (divides SUBS)
        static_assert(SUBS % QN == 0 && PFD < QN, "");
        const unsigned tileA = smemAddr(tile) + lane * 8 * R;
        const unsigned ringA = smemAddr(hring);
        const unsigned readyA = smemAddr(&cnt[0]), doneA = smemAddr(&cnt[1]), treadyA = smemAddr(&cnt[4]);
        auto subAddr = [&](int m) -> unsigned {
            const int n = m / NSUB, j = m - n * NSUB;
            return tileA + (unsigned)((n % NST) * STAGE_BYTES + (DIR > 0 ? j * SUBS : CHK - 1 - j * SUBS) * STEPB);
        };
        auto waitCnt = [&](unsigned addr, int need) {
            int v;
            do { asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); } while (v < need);
        };
        auto ldsR = [&](unsigned addr, double (&v)[R]) {
#pragma unroll
            for (int r = 0; r < R; r += 2)
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[r]), "=d"(v[r + 1]) : "r"(addr + (unsigned)(r * 8)) : "memory");
        };
        auto stsR = [&](unsigned addr, const double (&v)[R]) {
#pragma unroll
            for (int r = 0; r < R; r += 2)
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr + (unsigned)(r * 8)), "d"(v[r]), "d"(v[r + 1]) : "memory");
        };
        auto loadStep = [&](unsigned p, double (&a)[R], double (&x)[R], double (&yy)[R]) {
            ldsR(p, a);
            ldsR(p + (unsigned)TILE_BYTES, x);
            ldsR(p + (unsigned)(2 * TILE_BYTES), yy);
        };
        const bool isLC = lane == LC;
#ifdef SD_PROFILE
        long long tStart = clock64(), tWaitR = 0, tWaitT = 0;
#endif
        waitCnt(treadyA, 1);
        SD_COMPILER_BARRIER();
        double qa[QN][R], qx[QN][R], qy[QN][R];
#pragma unroll
        for (int i = 0; i < PFD; ++i) loadStep(subAddr(0) + (unsigned)(i * STEP), qa[i], qx[i], qy[i]);
        double y[R];
#pragma unroll
        for (int r = 0; r < R; ++r) y[r] = 0.0;
        double shq[SIGMA];  // the neighbour lane's last-row values of the previous SIGMA steps (oldest first)
#pragma unroll
        for (int i = 0; i < SIGMA; ++i) shq[i] = 0.0;
#pragma unroll 1
        for (int m = 0; m < nsub; ++m) {
            const unsigned cur = subAddr(m);
            unsigned nxt = cur;  // past the end the pre-load re-reads this sub-chunk (harmless)
            if (m + 1 < nsub) {
                nxt = subAddr(m + 1);
                if (((m + 1) % NSUB) == 0) {
                    SD_T0 waitCnt(treadyA, (m + 1) / NSUB + 1); SD_T1(tWaitT)  // next chunk landed (rarely waits)
                }
            }
            { SD_T0 waitCnt(readyA, m + 1); SD_T1(tWaitR) }  // hand-off values of this sub-chunk are in the ring
            SD_COMPILER_BARRIER();
            double rv[SUBS];  // lane LC's neighbour-row values (ring entries m*SUBS + e + 31)
#pragma unroll
            for (int e = 0; e < SUBS; ++e)
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(rv[e]) : "r"(ringA + (unsigned)(((m * SUBS + e + 31 * SIGMA) & (HR - 1)) * 8)) : "memory");
#pragma unroll
            for (int e = 0; e < SUBS; ++e) {
                // Program order = issue order of the one warp: the independent work (the (i-1,j) terms, the loads of step
                // e + PFD, the store) is placed in the shadows of the dependent instructions (shuffle -> DFMA -> DFMA).
                const int ee = e + PFD;
                const unsigned p = ee < SUBS ? cur + (unsigned)(ee * STEP) : nxt + (unsigned)((ee - SUBS) * STEP);
                const double (&a)[R] = qa[e % QN];
                const double (&x)[R] = qx[e % QN];
                const double (&yy)[R] = qy[e % QN];
                double in[R];
#pragma unroll
                for (int r = 0; r < R; ++r) in[r] = __fma_rn(-x[r], y[r], a[r]);
                double down = isLC ? rv[e] : shq[0];
                double sh;
                if (DIR > 0) {
                    y[0] = __fma_rn(-yy[0], down, in[0]);
                    ldsR(p, qa[ee % QN]);
#pragma unroll
                    for (int r = 1; r < R; ++r) y[r] = __fma_rn(-yy[r], y[r - 1], in[r]);
                    ldsR(p + (unsigned)TILE_BYTES, qx[ee % QN]);
                    asm volatile("{ .reg .b32 lo, hi; mov.b64 {lo,hi}, %1; shfl.sync.up.b32 lo, lo, 1, 0, 0xffffffff; "
                                 "shfl.sync.up.b32 hi, hi, 1, 0, 0xffffffff; mov.b64 %0, {lo,hi}; }" : "=d"(sh) : "d"(y[R - 1]) : "memory");
                } else {
                    y[R - 1] = __fma_rn(-yy[R - 1], down, in[R - 1]);
                    ldsR(p, qa[ee % QN]);
#pragma unroll
                    for (int r = R - 2; r >= 0; --r) y[r] = __fma_rn(-yy[r], y[r + 1], in[r]);
                    ldsR(p + (unsigned)TILE_BYTES, qx[ee % QN]);
                    asm volatile("{ .reg .b32 lo, hi; mov.b64 {lo,hi}, %1; shfl.sync.down.b32 lo, lo, 1, 31, 0xffffffff; "
                                 "shfl.sync.down.b32 hi, hi, 1, 31, 0xffffffff; mov.b64 %0, {lo,hi}; }" : "=d"(sh) : "d"(y[0]) : "memory");
                }
#pragma unroll
                for (int i = 0; i + 1 < SIGMA; ++i) shq[i] = shq[i + 1];
                shq[SIGMA - 1] = sh;
                stsR(cur + (unsigned)(e * STEP), y);
                ldsR(p + (unsigned)(2 * TILE_BYTES), qy[ee % QN]);
            }
            __syncwarp();
            SD_COMPILER_BARRIER();  // the y stores precede `done` in program order (same in-order pipeline)
            if (lane == 0) asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(doneA), "r"(m + 1) : "memory");
        }
#ifdef SD_PROFILE
        if (lane == 0 && ctl.prof) { ctl.prof[4 * q] = clock64() - tStart; ctl.prof[4 * q + 1] = tWaitT; ctl.prof[4 * q + 2] = tWaitR; ctl.prof[4 * q + 3] = tStart; }
#endif
    } else if (warp == 2) {
        // ------------------------------------------------------------------------------------------ post
        double acc = 0.0;
        const double postScalar = op.postScalar();
        double postAlpha = 0.0;
        if constexpr (Op::KIND == 3) postAlpha = op.postAlpha();
        int peerReady = 0;  // the consumer's `done` as last read (back-pressure of the in-cluster ring)
        const unsigned int peerRing = dsOut ? mapaShared(smemAddr(hring), rank + 1) : 0u;
        const unsigned int peerBar = dsOut ? mapaShared(smemAddr(hbar), rank + 1) : 0u;
        const unsigned int peerCnt = dsOut ? mapaShared(smemAddr(&cnt[1]), rank + 1) : 0u;  // the consumer's `done`
        for (int m = 0; m < nsub; ++m) {
            const int n = m / NSUB, j = m - n * NSUB;
            while (ldVolatileS32(&cnt[1]) < m + 1) {
#ifdef SD_SPIN_SLEEP
                __nanosleep(SD_SPIN_SLEEP);
#endif
            }
            SD_COMPILER_BARRIER();
            const volatile double* tp = tile + (size_t)(n % NST) * L::STAGE_DOUBLES;
            const int cn = DIR > 0 ? nLo + n : g.nchunks * (CH / CHK) - 1 - (nLo + n);  // storage chunk (of CHK steps)
            // last row first: it is on the next strip's critical path.  Lane l's value (our relative position
            // m*SUBS + l) feeds the consumer's relative position uc; only positions the consumer marches are sent.
            const int uc = base + m * SUBS + lane - 31 * SIGMA - cLo * CHK;
            const bool feeds = lane < SUBS && uc >= 0 && uc < cCnt * CHK;
            if (dsOut) {
                if (__any_sync(0xffffffffu, feeds)) {
                    // the ring slot of consumer position uc was last used by uc - HR: the consumer's SOLVER must be done
                    // with that sub-chunk (its pre warp may run up to a TMA ring ahead of it)
                    int ucMax = base + m * SUBS + SUBS - 1 - 31 * SIGMA - cLo * CHK;
                    if (ucMax > cCnt * CHK - 1) ucMax = cCnt * CHK - 1;
                    const int mC = ucMax / SUBS;
                    while (peerReady < mC - RB + 1)
                        asm volatile("ld.volatile.shared::cluster.s32 %0, [%1];" : "=r"(peerReady) : "r"(peerCnt) : "memory");
                }
                if (feeds) {
                    const int ls = DIR > 0 ? j * SUBS + lane : CHK - 1 - j * SUBS - lane;
                    const double yv = tp[(ls * 32 + LP) * R + (DIR > 0 ? R - 1 : 0)];
                    stAsyncU64(peerRing + (unsigned)(((uc + 31 * SIGMA) & (HR - 1)) * 8),
                               (unsigned long long)__double_as_longlong(yv), peerBar + (unsigned)(((uc / SUBS) % RB) * 8));
                }
            } else if (feeds) {
                const int ls = DIR > 0 ? j * SUBS + lane : CHK - 1 - j * SUBS - lane;
                const double yv = tp[(ls * 32 + LP) * R + (DIR > 0 ? R - 1 : 0)];
                stRelaxedU64(handOut + stepOf(base + m * SUBS + lane), (unsigned long long)__double_as_longlong(yv));
            }
            double* outp = op.out + stripBase + (size_t)cn * TILE + lane * R;
            static_assert(R == 2, "the post warp moves one 16-byte vector (the lane's two rows) per step and array");
            {
                // batches of QB steps: every shared-memory load of a batch is issued before the first use (volatile
                // accesses keep their program order, so the order is set here, not by the compiler)
                constexpr int QB = 4;
                static_assert(SUBS % QB == 0, "");
                const unsigned tpA = smemAddr(tile + (size_t)(n % NST) * L::STAGE_DOUBLES) + (unsigned)(lane * 16);
                // y-slab (Op::HALO): the first / last row of the new direction also goes into the neighbour rank's ghost row
                // (peer memory over NVLink, contiguous by column, distpeer.cuh).  Lanes 0..SUBS-1 recompute that row's SUBS
                // values of this sub-chunk from the tile (same FMA on the same operands as lane 0 / 31 below), so each row
                // leaves as ONE coalesced store and the post warp's own loop is untouched.  (Measured at 2 GPUs: scattered
                // 8-byte stores from one lane, or handing the values over by shuffle, cost the backward solve 25-45 us --
                // the post warp of the strip that starts the march fell behind its solver and throttled every strip after it.)
                bool haloLo = false, haloHi = false;
                if constexpr (OpHalo<Op>::value) {
                    haloLo = k == 0 && op.pushLo != nullptr && !(ctl.dbg & 16);
                    haloHi = k == g.nstrips - 1 && op.pushHi != nullptr && !(ctl.dbg & 16);
                }
#pragma unroll
                for (int e0 = 0; e0 < SUBS; e0 += QB) {
                    double y0[QB], y1[QB], a0[QB], a1[QB], b0[QB], b1[QB];
#pragma unroll
                    for (int q = 0; q < QB; ++q) {
                        const int ls = DIR > 0 ? j * SUBS + e0 + q : CHK - 1 - j * SUBS - e0 - q;
                        const unsigned o = tpA + (unsigned)(ls * 512);
                        asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(y0[q]), "=d"(y1[q]) : "r"(o) : "memory");
                        if (Op::KIND != 0)
                            asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a0[q]), "=d"(a1[q]) : "r"(o + (unsigned)(3 * TILE * 8)) : "memory");
                        if (Op::KIND == 3)
                            asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(b0[q]), "=d"(b1[q]) : "r"(o + (unsigned)(4 * TILE * 8)) : "memory");
                    }
#pragma unroll
                    for (int q = 0; q < QB; ++q) {
                        const int ls = DIR > 0 ? j * SUBS + e0 + q : CHK - 1 - j * SUBS - e0 - q;
                        double2* o1 = reinterpret_cast<double2*>(outp + ls * 32 * R);
                        if (Op::KIND == 1) {         // forward: out = D*y, partial sum of y*out
                            const double w0 = a0[q] * y0[q], w1 = a1[q] * y1[q];
                            acc = __fma_rn(y0[q], w0, acc);
                            acc = __fma_rn(y1[q], w1, acc);
                            *o1 = make_double2(w0, w1);
                        } else if (Op::KIND == 2) {  // backward fused with the direction update: out = y + beta*out_old
                            const double s0 = __fma_rn(postScalar, a0[q], y0[q]), s1 = __fma_rn(postScalar, a1[q], y1[q]);
                            *o1 = make_double2(s0, s1);

                        } else if constexpr (Op::KIND == 3) {
                            // ... and with the solution update p += alpha s (:451) of the OLD direction, which is in the tile
                            const double s0 = __fma_rn(postScalar, a0[q], y0[q]), s1 = __fma_rn(postScalar, a1[q], y1[q]);
                            *o1 = make_double2(s0, s1);
                            *reinterpret_cast<double2*>(op.out2 + stripBase + (size_t)cn * TILE + lane * R + ls * 32 * R) =
                                make_double2(__fma_rn(postAlpha, a0[q], b0[q]), __fma_rn(postAlpha, a1[q], b1[q]));

                        } else {
                            *o1 = make_double2(y0[q], y1[q]);
                        }
                    }
                }
                if constexpr (OpHalo<Op>::value) {
                    if ((haloLo || haloHi) && lane < SUBS) {
                        const int ls = DIR > 0 ? j * SUBS + lane : CHK - 1 - j * SUBS - lane;  // lane e takes step e of the sub-chunk
                        const int cLo = cn * CHK + ls, cHi = cn * CHK + ls - 31 * SIGMA;      // columns of lane 0 / lane 31 at that step
                        if (haloLo && cLo >= 0 && cLo < g.nx) {  // row 0 of lane 0
                            const double yv = tp[(ls * 32 + 0) * R + 0], so = tp[3 * TILE + (ls * 32 + 0) * R + 0];
                            op.pushLo[cLo] = __fma_rn(postScalar, so, yv);
                        }
                        if (haloHi && cHi >= 0 && cHi < g.nx) {  // row R-1 of lane 31
                            const double yv = tp[(ls * 32 + 31) * R + R - 1], so = tp[3 * TILE + (ls * 32 + 31) * R + R - 1];
                            op.pushHi[cHi] = __fma_rn(postScalar, so, yv);
                        }
                    }
                }
            }
            if (j == NSUB - 1) {
                __syncwarp();
                if (lane == 0) stVolatileS32(&cnt[2], n + 1);
            }
        }
        acc = warpSum(acc);
        if (lane == 0) op.stripDone(k, acc);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if constexpr (PRE) op.stripMax(k, *preRed);
        if (OpHalo<Op>::value && !(ctl.dbg & 32)) __threadfence_system();  // this CTA's stores into peer memory precede the stamp
        else __threadfence();
        int t = atomicAdd(ctl.finished, 1);
        if (t == g.nstrips - 1) {
            __threadfence();
            op.allDone(g.nstrips);
            *ctl.finished = 0;
            *ctl.ticket = 0;
            __threadfence();
        }
    }
    }  // q < nstrips
    // a peer's shared memory must stay alive until every push into it and every read of its counters is over
    if (CL > 1) clusterBarrier();
}

// Launch helper: clusters of `cl` CTAs (1, 2, 4 or 8) along the grid; the grid is padded to a multiple of cl.
template <class Op, int SIGMA, int DIR, int SUBS, int CL>
static inline cudaError_t launchSolveCl(const Op& op, const Geom& g, const Control& ctl, cudaStream_t stream) {
    auto kern = solveKernel<Op, SIGMA, DIR, SUBS, CL>;
    const size_t bytes = SolveLayout<Op>::BYTES;
    static bool attrSet[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attrSet[dev & 15]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        attrSet[dev & 15] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((g.nstrips + CL - 1) / CL * CL));
    cfg.blockDim = dim3(96);
    cfg.dynamicSmemBytes = bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, op, g, ctl);
}
template <class Op, int SIGMA, int DIR, int SUBS>
static inline cudaError_t launchSolve(const Op& op, const Geom& g, const Control& ctl, cudaStream_t stream, int cl) {
    switch (cl) {
        case 8: return launchSolveCl<Op, SIGMA, DIR, SUBS, 8>(op, g, ctl, stream);
        case 4: return launchSolveCl<Op, SIGMA, DIR, SUBS, 4>(op, g, ctl, stream);
        case 2: return launchSolveCl<Op, SIGMA, DIR, SUBS, 2>(op, g, ctl, stream);
        default: return launchSolveCl<Op, SIGMA, DIR, SUBS, 1>(op, g, ctl, stream);
    }
}

template <class Op, int R, int SIGMA, int DIR, int SUBS, int CL>
static inline cudaError_t launchSolveRCl(const Op& op, const Geom& g, const Control& ctl, cudaStream_t stream) {
    auto kern = solveKernelR<Op, R, SIGMA, DIR, SUBS, CL>;
    const size_t bytes = SolveLayoutR<Op, R, (R == 2 ? 16 : 8)>::BYTES;
    static bool attrSet[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attrSet[dev & 15]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        attrSet[dev & 15] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((g.nstrips + CL - 1) / CL * CL));
    cfg.blockDim = dim3(96);
    cfg.dynamicSmemBytes = bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, op, g, ctl);
}
template <class Op, int R, int SIGMA, int DIR, int SUBS>
static inline cudaError_t launchSolveR(const Op& op, const Geom& g, const Control& ctl, cudaStream_t stream, int cl) {
    switch (cl) {
        case 8: return launchSolveRCl<Op, R, SIGMA, DIR, SUBS, 8>(op, g, ctl, stream);
        case 4: return launchSolveRCl<Op, R, SIGMA, DIR, SUBS, 4>(op, g, ctl, stream);
        default: return launchSolveRCl<Op, R, SIGMA, DIR, SUBS, 1>(op, g, ctl, stream);
    }
}

#endif  // __CUDACC__

}  // namespace sd
