// wavefront.cuh -- exact anti-diagonal scheduling of the reference's Gauss-Seidel-type loops on sm_100a.
//
// Every sequential raster loop of the reference whose cell (i,j) depends on the *new* value of its
// march-previous neighbours (i-SX, j) and (i, j-SY) -- the MIC(0) factor and the two triangular solves
// (src/FluidSim2D.cpp:368-387, 397-421), the closest-particle sweeps (:775-792) and the eikonal sweeps
// (:845-902) -- is run by one kernel template:
//
//   * the padded grid is cut into strips of 32 rows; one CTA per strip (claimed in order through an
//     atomic ticket, so a strip's predecessor is always running: no deadlock whatever the residency);
//   * warp 0 ("solver") owns the strip: lane t owns row t and marches along x one column per step, one
//     column behind lane t-1; the (i, j-SY) value arrives by __shfl_up, the (i-SX, j) value stays in a
//     register.  This visits cells in an order consistent with the reference's raster order, so results
//     are those of the sequential loop, not of a Jacobi relaxation;
//   * warp 1 ("loader") streams 32x32 blocks of every input array into a 4-deep shared-memory ring with
//     16-byte cp.async (coalesced 256 B row segments), two blocks ahead of the solver;
//   * warp 2 ("courier") writes finished 32x32 output blocks back with coalesced 16-byte stores, and
//     carries the strip-to-strip dependency: the last row of strip k-1 is published column by column into
//     a global hand-off buffer whose words are self-validating (a reserved NaN payload means "not yet"),
//     so no fences sit on the producer's critical path; the courier polls it and feeds a shared ring.
//
// HBM traffic is exactly one read of each input and one write of each output per launch.
#pragma once

#include "common.cuh"

namespace wf {

constexpr unsigned long long SENT = 0x7FF8F51D0DEAD001ULL;  // reserved quiet-NaN payload: "not written yet"
constexpr int NSTAGE = 4;                                    // shared-memory ring depth (32-column blocks)
constexpr int HR = 128;                                      // hand-off ring (columns)

struct Domain {
    int ncb;      // 32-column blocks: physical columns [0, 32*ncb)
    int nstrips;  // 32-row strips:    physical rows    [0, 32*nstrips)
    int pitch;    // frame pitch (elements)
};

struct Control {
    int* ticket;               // zero between launches
    int* finished;             // zero between launches
    unsigned long long* hand;  // [W][nstrips][32*ncb], all SENT between launches
    const int* gate;           // optional: run only if (*gate != 0) == gateRunIfNonZero
    int gateRunIfNonZero;
    int* changed;              // optional: set to 1 if any cell reported a change
    double* dbgState;          // simple scheduler only: [W][32*nstrips][32*ncb]
};

__device__ __forceinline__ unsigned long long ldRelaxedU64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stRelaxedU64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void cpAsync16(void* smemDst, const void* gmemSrc) {
    unsigned int d = (unsigned int)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmemSrc) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class Op>
struct Layout {
    static constexpr int ROWS = Op::UPROW ? 33 : 32;
    static constexpr int IN_DOUBLES = NSTAGE * Op::NIN * ROWS * 32;
    static constexpr int OUT_DOUBLES = Op::INPLACE ? 0 : NSTAGE * Op::NOUT * 32 * 32;
    static constexpr int HAND_WORDS = Op::W * HR;
    static constexpr size_t BYTES = (size_t)(IN_DOUBLES + OUT_DOUBLES + HAND_WORDS) * 8 + 64;
};

// sync words in shared memory
enum { S_FILLED = 0, S_FREED = 1, S_FLUSHED = 2, S_HAND_AVAIL = 3, S_STRIP = 4, S_HAND_USED = 5 };

template <class Op, int SX, int SY>
__global__ void __launch_bounds__(96, 1) wavefrontKernel(Op op, Domain dom, Control ctl) {
    using L = Layout<Op>;
    constexpr int ROWS = L::ROWS;
    constexpr int NIN = Op::NIN, NOUT = Op::NOUT, W = Op::W;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    double* stageIn = reinterpret_cast<double*>(smemRaw);
    double* stageOut = stageIn + L::IN_DOUBLES;
    unsigned long long* handRing = reinterpret_cast<unsigned long long*>(stageOut + L::OUT_DOUBLES);
    volatile int* sync = reinterpret_cast<volatile int*>(handRing + L::HAND_WORDS);

    if (ctl.gate) {
        int g = *ctl.gate;
        if ((g != 0) != (ctl.gateRunIfNonZero != 0)) return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        sync[S_STRIP] = atomicAdd(ctl.ticket, 1);
        sync[S_FILLED] = 0; sync[S_FREED] = 0; sync[S_FLUSHED] = 0; sync[S_HAND_AVAIL] = 0; sync[S_HAND_USED] = 0;
    }
    __syncthreads();
    const int strip = sync[S_STRIP];
    const int nblocks = dom.ncb;
    const int ncols = dom.ncb * 32;
    const int nrowsPad = dom.nstrips * 32;
    const size_t handStride = (size_t)dom.nstrips * ncols;  // per state word

    auto physRow = [&](int lr) { int R = strip * 32 + lr; return SY > 0 ? R : nrowsPad - 1 - R; };

    if (warp == 1) {
        // ------------------------------------------------------------------ loader
        for (int b = 0; b < nblocks; ++b) {
            // slot b%NSTAGE is reusable once block b-NSTAGE is consumed (and, for in-place ops, written back)
            if (Op::INPLACE) { while (b - sync[S_FLUSHED] >= NSTAGE) {} }
            else { while (b - sync[S_FREED] >= NSTAGE) {} }
            const int slot = b % NSTAGE;
            const int bp = SX > 0 ? b : nblocks - 1 - b;
            constexpr int CHUNKS = NIN * ROWS * 16;
            for (int q = lane; q < CHUNKS; q += 32) {
                int k = q / (ROWS * 16);
                int rem = q - k * (ROWS * 16);
                int lr = rem >> 4, ch = rem & 15;
                const double* src = op.in[k] + (long long)physRow(lr) * dom.pitch + 32 * bp + 2 * ch;
                double* dst = stageIn + ((slot * NIN + k) * ROWS + lr) * 32 + 2 * ch;
                cpAsync16(dst, src);
            }
            cpAsyncCommit();
            if (b >= 1) {
                cpAsyncWait<1>();  // everything but the newest group has landed
                __syncwarp();
                __threadfence_block();
                if (lane == 0) sync[S_FILLED] = b;
            }
        }
        cpAsyncWait<0>();
        __syncwarp();
        __threadfence_block();
        if (lane == 0) sync[S_FILLED] = nblocks;
    } else if (warp == 2) {
        // ------------------------------------------------------------------ courier: hand-off in, blocks out
        int flushed = 0, handCol = 0;
        const bool hasProducer = strip > 0;
        const unsigned long long* handIn = ctl.hand + (size_t)(hasProducer ? strip - 1 : 0) * ncols;
        while (flushed < nblocks || (hasProducer && handCol < ncols)) {
            if (hasProducer && handCol < ncols) {
                int col = handCol + lane;
                bool want = col < ncols && (col - sync[S_HAND_USED]) < HR;
                unsigned long long v[W];
                bool ok = want;
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    v[w] = want ? ldRelaxedU64(handIn + w * handStride + col) : SENT;
                    ok = ok && (v[w] != SENT);
                }
                unsigned int m = __ballot_sync(0xffffffffu, ok);
                int n = __ffs(~m) - 1;  // length of the valid prefix (32 if all valid)
                if (m == 0xffffffffu) n = 32;
                if (lane < n) {
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        handRing[w * HR + (col % HR)] = v[w];
                        stRelaxedU64(const_cast<unsigned long long*>(handIn) + w * handStride + col, SENT);
                    }
                }
                if (n > 0) {
                    __syncwarp();
                    __threadfence_block();
                    handCol += n;
                    if (lane == 0) sync[S_HAND_AVAIL] = handCol;
                }
            }
            if (flushed < sync[S_FREED]) {
                __threadfence_block();
                const int slot = flushed % NSTAGE;
                const int bp = SX > 0 ? flushed : nblocks - 1 - flushed;
#pragma unroll
                for (int k = 0; k < NOUT; ++k) {
#pragma unroll 4
                    for (int it = 0; it < 16; ++it) {
                        int lr = 2 * it + (lane >> 4), ch = lane & 15;
                        const double* src = Op::INPLACE ? stageIn + ((slot * NIN + k) * ROWS + lr) * 32 + 2 * ch
                                                        : stageOut + ((slot * NOUT + k) * 32 + lr) * 32 + 2 * ch;
                        double* dst = op.out[k] + (long long)physRow(lr) * dom.pitch + 32 * bp + 2 * ch;
                        *reinterpret_cast<double2*>(dst) = *reinterpret_cast<const double2*>(src);
                    }
                }
                __syncwarp();
                __threadfence_block();
                ++flushed;
                if (lane == 0) sync[S_FLUSHED] = flushed;
            }
        }
    } else {
        // ------------------------------------------------------------------ solver
        double bnd[W], mine[W];
        op.boundaryState(bnd);
#pragma unroll
        for (int w = 0; w < W; ++w) mine[w] = bnd[w];
        double acc = 0.0;
        bool changed = false;
        const int j = physRow(lane);
        const int nsteps = ncols + 31;
        for (int s = 0; s < nsteps; ++s) {
            if ((s & 31) == 0) {
                const int b = s >> 5;
                if (lane == 0) sync[S_FREED] = b > 0 ? b - 1 : 0;
                if (b < nblocks) {
                    int need = b + 1 + (Op::LOOK ? 1 : 0);
                    if (need > nblocks) need = nblocks;
                    while (sync[S_FILLED] < need) {}
                    if (!Op::INPLACE) { while (sync[S_FLUSHED] < b - (NSTAGE - 1)) {} }
                    __threadfence_block();
                }
            }
            const int c = s - lane;
            const bool active = (c >= 0) && (c < ncols);
            double left[W], down[W];
#pragma unroll
            for (int w = 0; w < W; ++w) {
                left[w] = mine[w];
                down[w] = __shfl_up_sync(0xffffffffu, mine[w], 1);
            }
            if (lane == 0) {
                if (strip > 0 && active) {
                    while (sync[S_HAND_AVAIL] <= c) {}
                    __threadfence_block();
#pragma unroll
                    for (int w = 0; w < W; ++w) down[w] = __longlong_as_double((long long)handRing[w * HR + (c % HR)]);
                    sync[S_HAND_USED] = c + 1;
                } else {
#pragma unroll
                    for (int w = 0; w < W; ++w) down[w] = bnd[w];
                }
            }
            if (active) {
                const int i = SX > 0 ? c : ncols - 1 - c;
                const int b = c >> 5, slot = b % NSTAGE, ci = i & 31;
                double own[NIN], right[NIN], up[NIN];
#pragma unroll
                for (int k = 0; k < NIN; ++k) {
                    own[k] = stageIn[((slot * NIN + k) * ROWS + lane) * 32 + ci];
                    if (Op::UPROW) up[k] = stageIn[((slot * NIN + k) * ROWS + lane + 1) * 32 + ci];
                    else up[k] = 0.0;
                    if (Op::LOOK) {
                        int b2 = (c + 1) >> 5;
                        if (b2 > nblocks - 1) b2 = nblocks - 1;
                        right[k] = stageIn[(((b2 % NSTAGE) * NIN + k) * ROWS + lane) * 32 + ((i + SX) & 31)];
                    } else right[k] = 0.0;
                }
                double o[NOUT > 0 ? NOUT : 1], stNew[W];
                changed |= op.cell(i, j, own, right, up, left, down, o, stNew, acc);
#pragma unroll
                for (int k = 0; k < NOUT; ++k) {
                    if (Op::INPLACE) stageIn[((slot * NIN + k) * ROWS + lane) * 32 + ci] = o[k];
                    else stageOut[((slot * NOUT + k) * 32 + lane) * 32 + ci] = o[k];
                }
#pragma unroll
                for (int w = 0; w < W; ++w) mine[w] = stNew[w];
                if (lane == 31 && strip + 1 < dom.nstrips) {
#pragma unroll
                    for (int w = 0; w < W; ++w)
                        stRelaxedU64(ctl.hand + w * handStride + (size_t)strip * ncols + c,
                                     (unsigned long long)__double_as_longlong(stNew[w]));
                }
            } else {
#pragma unroll
                for (int w = 0; w < W; ++w) mine[w] = bnd[w];
            }
        }
        __syncwarp();
        __threadfence_block();
        if (lane == 0) sync[S_FREED] = nblocks;
        if (ctl.changed && __any_sync(0xffffffffu, changed) && lane == 0) atomicOr(ctl.changed, 1);
        acc = warpSum(acc);
        if (lane == 0) op.stripDone(strip, acc);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        int t = atomicAdd(ctl.finished, 1);
        if (t == dom.nstrips - 1) {
            op.allDone(dom.nstrips);
            *ctl.finished = 0;
            *ctl.ticket = 0;
            __threadfence();
        }
    }
}

// Debug scheduler: one CTA walks the anti-diagonals with a block barrier in between.  Same Op, same
// arithmetic, trivially correct ordering; used to bisect scheduling bugs (fsim_options.debugSimpleWavefront).
template <class Op, int SX, int SY>
__global__ void __launch_bounds__(1024, 1) wavefrontSimpleKernel(Op op, Domain dom, Control ctl) {
    constexpr int NIN = Op::NIN, NOUT = Op::NOUT, W = Op::W;
    if (ctl.gate) {
        int g = *ctl.gate;
        if ((g != 0) != (ctl.gateRunIfNonZero != 0)) return;
    }
    __shared__ double red[32];
    __shared__ int anyChanged;
    if (threadIdx.x == 0) anyChanged = 0;
    const int ncols = dom.ncb * 32, nrows = dom.nstrips * 32;
    const size_t plane = (size_t)ncols * nrows;
    double bnd[W];
    op.boundaryState(bnd);
    double acc = 0.0;
    bool changed = false;
    __syncthreads();
    for (int d = 0; d < ncols + nrows - 1; ++d) {
        for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
            int c = d - r;
            if (c < 0 || c >= ncols) continue;
            int i = SX > 0 ? c : ncols - 1 - c;
            int j = SY > 0 ? r : nrows - 1 - r;
            double own[NIN], right[NIN], up[NIN], left[W], down[W], o[NOUT > 0 ? NOUT : 1], st[W];
            for (int k = 0; k < NIN; ++k) {
                const double* a = op.in[k] + (long long)j * dom.pitch + i;
                own[k] = __ldcg(a);
                right[k] = Op::LOOK ? __ldcg(a + SX) : 0.0;
                up[k] = Op::UPROW ? __ldcg(a + (long long)SY * dom.pitch) : 0.0;
            }
            for (int w = 0; w < W; ++w) {
                left[w] = c > 0 ? __ldcg(ctl.dbgState + w * plane + (size_t)r * ncols + c - 1) : bnd[w];
                down[w] = r > 0 ? __ldcg(ctl.dbgState + w * plane + (size_t)(r - 1) * ncols + c) : bnd[w];
            }
            changed |= op.cell(i, j, own, right, up, left, down, o, st, acc);
            for (int k = 0; k < NOUT; ++k) op.out[k][(long long)j * dom.pitch + i] = o[k];
            for (int w = 0; w < W; ++w) ctl.dbgState[w * plane + (size_t)r * ncols + c] = st[w];
        }
        __threadfence_block();
        __syncthreads();
    }
    if (changed) anyChanged = 1;
    acc = blockReduce<false>(acc, red);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (ctl.changed && anyChanged) atomicOr(ctl.changed, 1);
        op.stripDone(0, acc);
        // partial slots of the other strips must read as zero for allDone's fixed-order sum
        for (int k = 1; k < dom.nstrips; ++k) op.stripDone(k, 0.0);
        __threadfence();
        op.allDone(dom.nstrips);
    }
}

}  // namespace wf
