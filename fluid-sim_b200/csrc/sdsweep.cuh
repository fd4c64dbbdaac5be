// sdsweep.cuh -- in-place Gauss-Seidel sweeps on the strip-diagonal (SD) layout of sdwave.cuh.
//
// The reference's closest-particle sweeps (src/FluidSim2D.cpp:775-792), eikonal sweeps (:845-902) and MIC(0)
// factor loop (:368-387) visit the cells in raster order and update them in place: cell (i,j) sees this sweep's
// values at its march-previous neighbours and the previous sweep's values at the other two.  Any schedule that
// keeps "previous before next" along x and along y reproduces that loop exactly.  Here one warp marches a strip
// of 32 rows, lane t on row t, SIGMA columns behind lane t-1, and the strip's tiles are updated IN PLACE in
// shared memory -- so all four neighbours of (step u, lane t) are plain tile reads:
//
//      previous column  [u-1][t]          next column  [u+1][t]           (same lane)
//      previous row     [u-SIGMA][t-DIR]  next row     [u+SIGMA][t+DIR]   (neighbour lanes)
//
// and "new or old" takes care of itself: a slot holds the new value exactly when its cell has been visited.
// Two lanes look outside the strip: lane LC's previous row is the neighbouring strip's last row (this sweep's
// values: they arrive through the hand-off ring, pushed through distributed shared memory inside a thread-block
// cluster, through self-validating global slots between clusters), lane LP's next row is the next strip's first
// row (previous sweep's values: gathered from global memory before that strip can have touched them).
//
// Sweeps that run x descending while y ascends (or the reverse) use the x-mirrored copy of the layout, so the
// kernel only knows DIR = +1 (both ascending in layout coordinates) and DIR = -1 (both descending); the Op is told
// whether layout column c is grid column c or nx-1-c.
//
// Warps: 0 = solver, 1 = pre (TMA loads, next-row gather, hand-off in), 2 = post (hand-off out first, write-back of
// finished chunks one sub-chunk late -- the solver still reads a chunk's last SIGMA steps from the next one).
#pragma once

#include "sdwave.cuh"

namespace sd {

struct SweepControl {
    int* ticket;               // zero between launches
    int* finished;             // zero between launches
    unsigned long long* hand;  // [NN planes][nstrips + 1][handStride(g)]; polled slots are SENT between launches
    size_t planeWords;         // words per plane
    const int* gate;           // optional: the kernel runs only if *gate != 0
    int* changed;              // optional: set to 1 if any cell changed
};

#ifdef __CUDACC__

// Op: NA tile arrays (all loaded), the first NW are written back, the first NN are what neighbours see (and what
// is handed from strip to strip).  double* arr[NA];
//   __device__ bool cell(int c, int j, double (&own)[NA], const double (&pc)[NN], const double (&nc)[NN],
//                        const double (&pr)[NN], const double (&nr)[NN]) const
// with pc/nc = previous/next column and pr/nr = previous/next row IN MARCH ORDER; returns "changed".
// Ops with SKIP = true (the level-set sweeps) let the kernel pass over sub-chunks none of whose inputs can have changed since
// the visit that last looked at them -- an exact no-op, see the solver warp.  Such an Op provides
//   const unsigned char* tileNeg; int* tileStamp;   [strips][nblk] per (strip of 32 rows, block of 32 unmirrored columns):
//                                                   "holds a cell the sweep may change" / index of the last sweep that changed one
//   int nblk, t, mirror, noSkip;                    blocks per strip, index of this sweep, layout column c = grid column nx-1-c,
//                                                   debug switch (1 = visit everything)
//   SKIP_FIRST, SKIP_WINDOW, SKIP_ALLNB             the first SKIP_FIRST sweeps visit every chunk that holds a changeable cell;
//                                                   later ones only chunks near a change of the last SKIP_WINDOW sweeps, where
//                                                   "near" covers the march-previous neighbours (a visit reads only those) or,
//                                                   with SKIP_ALLNB, all four neighbours
// Plane 0 of the values handed from strip to strip must be non-negative wherever its sign matters to the Op: on the hand-off
// path its sign bit carries "changed in this sweep" and the kernel passes fabs() of it to the Op.
template <class Op, class = void>
struct OpSkip { static constexpr bool value = false; };
template <class Op>
struct OpSkip<Op, std::void_t<decltype(Op::SKIP)>> { static constexpr bool value = Op::SKIP; };

template <class Op>
struct SweepLayout {
    static constexpr int NST = Op::NA <= 1 ? 8 : 4;  // ring stages (power of two, >= 3: the write-back runs a sub-chunk late)
    static constexpr int RS = NST * CH;                                   // ring steps
    static constexpr size_t TILE_DOUBLES = (size_t)Op::NA * RS * 32;
    // tile | ringNew[NN][HR] | ringOld[NN][RS] | full[NST] | hbar[HR/4] | counters
    // tile | ringNew | ringOld | full | hbar | counters (64 B) | sdirty[NST] + lpChg[64] (SKIP ops, 128 B)
    static constexpr size_t BYTES = (TILE_DOUBLES + (size_t)Op::NN * HR + (size_t)Op::NN * RS) * 8 + NST * 8 + (HR / 4) * 8 + 64 + 128;
};

template <class Op, int SIGMA, int DIR, int SUBS, int CL>
__global__ void __launch_bounds__(96, 1) sweepKernel(Op op, Geom g, SweepControl ctl) {
    static_assert(HR % SUBS == 0 && CH % SUBS == 0, "sub-chunks tile the rings");
    using L = SweepLayout<Op>;
    constexpr int NA = Op::NA, NW = Op::NW, NN = Op::NN, NST = L::NST, RS = L::RS, NSUB = CH / SUBS;
    constexpr int LC = DIR > 0 ? 0 : 31, LP = DIR > 0 ? 31 : 0;
    constexpr int TILE = CH * 32;
    constexpr int RB = HR / SUBS;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    double* tile = reinterpret_cast<double*>(smemRaw);                 // [NA][RS][32]
    double* ringNew = tile + L::TILE_DOUBLES;                           // [NN][HR]  previous row of lane LC
    double* ringOld = ringNew + (size_t)NN * HR;                        // [NN][RS]  next row of lane LP
    unsigned long long* full = reinterpret_cast<unsigned long long*>(ringOld + (size_t)NN * RS);
    unsigned long long* hbar = full + NST;
    int* cnt = reinterpret_cast<int*>(hbar + HR / 4);  // [0] ready, [1] done, [2] freed chunks, [3] ticket, [4] tready
    constexpr bool SKIP = OpSkip<Op>::value;
    volatile unsigned char* sdirty = reinterpret_cast<volatile unsigned char*>(cnt + 16);  // [NST] chunk may change (pre -> solver)
    volatile unsigned char* lpChg = sdirty + 16;                                            // [64] lane LP's changed bits per sub-chunk

    if (ctl.gate && *ctl.gate == 0) return;  // uniform over the grid
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int rank = CL > 1 ? clusterCtaRank() : 0u;
    for (int i = threadIdx.x; i < NN * HR; i += 96) ringNew[i] = 0.0;
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
            const int q0 = atomicAdd(ctl.ticket, CL);
            for (int r = 0; r < CL; ++r)
                asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(mapaShared(smemAddr(&cnt[3]), r)), "r"(q0 + r) : "memory");
        }
        clusterBarrier();
    } else {
        __syncthreads();
    }
    const int q = cnt[3];
    if (q < g.nstrips) {
    const int k = DIR > 0 ? q : g.nstrips - 1 - q;
    const bool hasProducer = q > 0;
    const bool dsIn = CL > 1 && rank > 0;
    const bool dsOut = CL > 1 && rank < CL - 1 && q < g.nstrips - 1;
    const size_t stripBase = (size_t)k * g.Sp * 32;
    const size_t hstride = (size_t)g.Sp + 31 * SIGMA + 33;
    unsigned long long* handOut = ctl.hand + (size_t)(q < g.nstrips - 1 ? q : g.nstrips) * hstride + (31 - LP) * SIGMA;
    unsigned long long* handIn = ctl.hand + (size_t)(q > 0 ? q - 1 : 0) * hstride + (31 - LC) * SIGMA;
    const int nchunks = g.nchunks, nsub = nchunks * NSUB, Sp = g.Sp;
    auto stepOf = [&](int u) { return DIR > 0 ? u : Sp - 1 - u; };                       // march position -> storage step
    auto slotOf = [&](int u) { return DIR > 0 ? (u & (RS - 1)) : ((u & (RS - 1)) ^ 31); };  // march position -> ring step

    if (warp == 1) {
        // ------------------------------------------------------------------------------------------ pre
        int issued = 0, landed = 0;
        auto issueLoads = [&]() {
            if (lane == 0) {
                const int lim = ldVolatileS32(&cnt[2]) + NST;
                while (issued < nchunks && issued < lim) {
                    const int cn = DIR > 0 ? issued : nchunks - 1 - issued, st = issued % NST;
                    mbarExpectTx(&full[st], NA * TILE * 8);
#pragma unroll
                    for (int a = 0; a < NA; ++a)
                        bulkLoad(tile + ((size_t)a * RS + (size_t)st * CH) * 32, op.arr[a] + stripBase + (size_t)cn * TILE, TILE * 8, &full[st]);
                    ++issued;
                }
            }
        };
        // next-row values (previous sweep's) of lane LP for the march positions of chunk n
        const int kn = DIR > 0 ? k + 1 : k - 1;  // strip that owns lane LP's next row
        auto gatherOld = [&](int n) {
            const int u = n * CH + lane;
            const int c = stepOf(u) - SIGMA * LP;  // column of lane LP at this position
            const bool ok = kn >= 0 && kn < g.nstrips && c >= 0 && c < g.nx;
            // row 32*kn + (31 - LP) is lane (31 - LP) of strip kn; its column c sits at step c + SIGMA*(31 - LP)
            const size_t src = ((size_t)kn * Sp + (size_t)(c + SIGMA * (31 - LP))) * 32 + (31 - LP);
#pragma unroll
            for (int a = 0; a < NN; ++a) {
                double v = 0.0;
                if (ok) v = __ldcg(op.arr[a] + src);
                asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(smemAddr(&ringOld[a * RS + (u & (RS - 1))])), "d"(v) : "memory");
            }
        };
        auto land = [&](int upTo) {
            if (upTo > nchunks) upTo = nchunks;
            issueLoads();
            while (landed < upTo) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                while (true) {  // the stage may still be in use
                    issueLoads();
                    if (__shfl_sync(0xffffffffu, issued, 0) > landed) break;
                    __nanosleep(100);
                }
                gatherOld(landed);  // (its ring slots were last read one ring lap ago, like the stage's)
                if constexpr (SKIP) {
                    // can anything in this chunk change?  Only if it holds a cell the sweep may change at all and (first
                    // round: no earlier sweep of this direction) or one of the blocks its cells and their upwind neighbours
                    // lie in -- own strip and march-previous strip, the chunk's columns plus the upwind one -- changed in
                    // one of the three sweeps since this direction last ran.  Changes made earlier in THIS sweep reach the
                    // solver through its own bookkeeping and the hand-off flags.
                    const int cn = DIR > 0 ? landed : nchunks - 1 - landed;
                    constexpr bool ALLNB = Op::SKIP_ALLNB;
                    int clo = 32 * cn - 31 * SIGMA - ((ALLNB || DIR > 0) ? 1 : 0), chi = 32 * cn + 31 + ((ALLNB || DIR < 0) ? 1 : 0);
                    if (clo < 0) clo = 0;
                    if (chi > g.nx - 1) chi = g.nx - 1;
                    bool ownNeg = false, recent = false;
                    if (clo <= chi) {
                        const int wlo = op.mirror ? g.nx - 1 - chi : clo, whi = op.mirror ? g.nx - 1 - clo : chi;
                        const int bLo = wlo >> 5, bHi = whi >> 5;
                        const int grp = lane >> 3;  // 0: own strip, 1: march-previous strip, 2: march-next strip (ALLNB)
                        const int b = bLo + (lane & 7), strip = grp == 0 ? k : (grp == 1 ? k - DIR : k + DIR);
                        const bool ok = grp < (ALLNB ? 3 : 2) && b <= bHi && strip >= 0 && strip < g.nstrips;
                        bool ng = false;
                        int stp = -100;
                        if (ok) { ng = op.tileNeg[strip * op.nblk + b] != 0; stp = __ldcg(op.tileStamp + strip * op.nblk + b); }
                        ownNeg = grp == 0 && ng;
                        recent = ng && stp >= op.t - Op::SKIP_WINDOW;
                    }
                    ownNeg = __any_sync(0xffffffffu, ownNeg);
                    recent = __any_sync(0xffffffffu, recent);
                    if (lane == 0) sdirty[landed % NST] = (op.noSkip || (ownNeg && (op.t < Op::SKIP_FIRST || recent))) ? 1 : 0;
                }
                mbarWait(&full[landed % NST], (unsigned int)((landed / NST) & 1));
                ++landed;
            }
            __syncwarp();
            if (lane == 0) stVolatileS32(&cnt[4], landed);
        };
        auto needAt = [&](int u) {
            const int c = stepOf(u) - SIGMA * LC;
            return hasProducer && u < Sp && c >= 0 && c < g.nx;
        };
        const bool glIn = hasProducer && !dsIn;
        const int grp = lane / SUBS;
        int myU = lane;
        bool need = glIn && needAt(myU);
        unsigned long long hv[NSUB][NN];
#pragma unroll
        for (int i = 0; i < NSUB; ++i)
#pragma unroll
            for (int a = 0; a < NN; ++a)
                hv[i][a] = (need && grp == i) ? ldRelaxedU64(handIn + (size_t)a * ctl.planeWords + stepOf(myU)) : 0ULL;
        int waited = 0;
        if (dsIn && lane == 0)
            for (int i = 0; i < RB; ++i) mbarExpectTx(&hbar[i], SUBS * 8 * NN);
        land(2);
        for (int n = 0; n < nchunks; ++n) {
            if (!hasProducer) {
                if (lane == 0) stVolatileS32(&cnt[0], (n + 1) * NSUB);
            } else if (dsIn) {
#pragma unroll
                for (int j = 0; j < NSUB; ++j) {
                    const int m = n * NSUB + j;
                    int target = ((m + 1) * SUBS - 1 + 31 * SIGMA) / SUBS;
                    if (target > nsub - 1) target = nsub - 1;
                    while (waited <= target) {
                        mbarWait(&hbar[waited % RB], (unsigned int)((waited / RB) & 1));
                        if (lane == 0) mbarExpectTx(&hbar[waited % RB], SUBS * 8 * NN);  // arm the slot's next phase
                        ++waited;
                    }
                    if (lane == 0) stVolatileS32(&cnt[0], m + 1);
                }
            } else {
#pragma unroll
                for (int j = 0; j < NSUB; ++j) {
                    const bool mine = grp == j;
                    while (true) {
                        bool valid = true;
#pragma unroll
                        for (int a = 0; a < NN; ++a) valid = valid && (!mine || !need || hv[j][a] != SENT);
                        if (__all_sync(0xffffffffu, valid)) break;
                        if (!valid) {
#pragma unroll
                            for (int a = 0; a < NN; ++a)
                                if (hv[j][a] == SENT) hv[j][a] = ldRelaxedU64(handIn + (size_t)a * ctl.planeWords + stepOf(myU));
                        }
                    }
                    if (mine) {
#pragma unroll
                        for (int a = 0; a < NN; ++a) {
                            double h = 0.0;
                            if (need) {
                                h = __longlong_as_double((long long)hv[j][a]);
                                stRelaxedU64(handIn + (size_t)a * ctl.planeWords + stepOf(myU), SENT);  // clean for the next launch
                            }
                            asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(smemAddr(&ringNew[a * HR + ((myU + 31 * SIGMA) & (HR - 1))])), "d"(h) : "memory");
                        }
                        myU += 32;
                        need = needAt(myU);
#pragma unroll
                        for (int a = 0; a < NN; ++a)
                            hv[j][a] = need ? ldRelaxedU64(handIn + (size_t)a * ctl.planeWords + stepOf(myU)) : 0ULL;
                    }
                    __syncwarp();
                    if (lane == 0) stVolatileS32(&cnt[0], n * NSUB + j + 1);
                }
            }
            land(n + 3);
        }
    } else if (warp == 0) {
        // ------------------------------------------------------------------------------------------ solver
        const int j = 32 * k + lane;
        const bool isLC = lane == LC, isLP = lane == LP;
        bool changed = false;
        auto waitCnt = [&](const int* p, int need) { while (ldVolatileS32(p) < need) {} };
        // plain loads: __syncwarp() orders them against the warp's own in-place stores, the counter polls (volatile +
        // compiler barrier) against the other warps' writes; neighbours an Op ignores cost nothing
        const double* vt = tile;
        const double* vNew = ringNew;
        const double* vOld = ringOld;
        const int lanePr = isLC ? lane : lane - DIR, laneNr = isLP ? lane : lane + DIR;
        double pcReg[NN];
#pragma unroll
        for (int a = 0; a < NN; ++a) pcReg[a] = 0.0;
        bool prevChanged = false, chunkChanged = false;  // (SKIP) a cell of the previous sub-chunk / of this chunk changed
#pragma unroll 1
        for (int m = 0; m < nsub; ++m) {
            const int n = m / NSUB;
            waitCnt(&cnt[4], n + 2 < nchunks ? n + 2 : nchunks);  // this chunk and the next have landed
            waitCnt(&cnt[0], m + 1);                             // hand-off values of this sub-chunk are in the ring
            SD_COMPILER_BARRIER();
            bool doEval = true;
            if constexpr (SKIP) {
                // Exact skip: a visit is a function of the cell, its previous column and its previous row.  If none of them
                // has changed since this direction's last sweep (chunk not dirty), nor earlier in this sweep (previous
                // sub-chunk quiet, no flagged hand-off value), every visit of this sub-chunk would rewrite what is there.
                double hv = 0.0;
                if (lane < SUBS) hv = vNew[(m * SUBS + lane + 31 * SIGMA) & (HR - 1)];
                const bool hf = __any_sync(0xffffffffu, lane < SUBS && (__double_as_longlong(hv) < 0));
                doEval = sdirty[n % NST] != 0 || prevChanged || hf;
            }
            bool subChanged = false;
            unsigned int lpBits = 0;
            if (doEval) {
#pragma unroll 4
                for (int e = 0; e < SUBS; ++e) {
                    const int u = m * SUBS + e;
                    const int c = stepOf(u) - SIGMA * lane;
                    const int sOwn = slotOf(u), sPc = slotOf(u - 1), sNc = slotOf(u + 1), sPr = slotOf(u - SIGMA), sNr = slotOf(u + SIGMA);
                    double own[NA], pc[NN], nc[NN], pr[NN], nr[NN];
#pragma unroll
                    for (int a = 0; a < NA; ++a) own[a] = vt[((size_t)a * RS + sOwn) * 32 + lane];
#pragma unroll
                    for (int a = 0; a < NN; ++a) {
                        // (Op::KEEP_PC: the previous column's new value is this lane's own last result -- kept in a register
                        // instead of a store/load round trip through the tile on the dependent chain)
                        pc[a] = Op::KEEP_PC ? pcReg[a] : vt[((size_t)a * RS + sPc) * 32 + lane];
                        nc[a] = vt[((size_t)a * RS + sNc) * 32 + lane];
                        const double prT = vt[((size_t)a * RS + sPr) * 32 + lanePr];
                        double prG = vNew[a * HR + ((u + 31 * SIGMA) & (HR - 1))];
                        if (SKIP && a == 0) prG = fabs(prG);  // (its sign bit is the hand-off's "changed" flag)
                        pr[a] = isLC ? prG : prT;
                        const double nrT = vt[((size_t)a * RS + sNr) * 32 + laneNr];
                        const double nrG = vOld[a * RS + (u & (RS - 1))];
                        nr[a] = isLP ? nrG : nrT;
                    }
                    if (op.cell(c, j, own, pc, nc, pr, nr)) {
                        changed = true;
                        subChanged = true;
                        lpBits |= 1u << e;
#pragma unroll
                        for (int a = 0; a < NW; ++a) tile[((size_t)a * RS + sOwn) * 32 + lane] = own[a];
                    }
#pragma unroll
                    for (int a = 0; a < NN; ++a) pcReg[a] = own[a];
                    __syncwarp();
                }
            } else if (Op::KEEP_PC) {
                // the next sub-chunk's first visit needs this lane's value at the last position passed over
#pragma unroll
                for (int a = 0; a < NN; ++a) pcReg[a] = vt[((size_t)a * RS + slotOf(m * SUBS + SUBS - 1)) * 32 + lane];
            }
            if constexpr (SKIP) {
                prevChanged = __any_sync(0xffffffffu, subChanged);
                chunkChanged |= prevChanged;
                if (isLP) lpChg[m & 63] = (unsigned char)lpBits;  // rides on the hand-off values' sign bits (post warp)
                if ((m + 1) % NSUB == 0) {
                    if (chunkChanged && lane < 8) {
                        // stamp the blocks this chunk's cells lie in (conservative: the chunk's whole column range)
                        const int cn = DIR > 0 ? n : nchunks - 1 - n;
                        int clo = 32 * cn - 31 * SIGMA, chi = 32 * cn + 31;
                        if (clo < 0) clo = 0;
                        if (chi > g.nx - 1) chi = g.nx - 1;
                        if (clo <= chi) {
                            const int wlo = op.mirror ? g.nx - 1 - chi : clo, whi = op.mirror ? g.nx - 1 - clo : chi;
                            const int b = (wlo >> 5) + lane;
                            if (b <= (whi >> 5)) op.tileStamp[k * op.nblk + b] = op.t;
                        }
                    }
                    chunkChanged = false;
                }
            }
            SD_COMPILER_BARRIER();
            if (lane == 0) stVolatileS32(&cnt[1], m + 1);
        }
        if (ctl.changed && __any_sync(0xffffffffu, changed) && lane == 0) atomicOr(ctl.changed, 1);
    } else {
        // ------------------------------------------------------------------------------------------ post
        int peerReady = 0, written = 0;
        const unsigned int peerRing = dsOut ? mapaShared(smemAddr(ringNew), rank + 1) : 0u;
        const unsigned int peerBar = dsOut ? mapaShared(smemAddr(hbar), rank + 1) : 0u;
        const unsigned int peerCnt = dsOut ? mapaShared(smemAddr(&cnt[1]), rank + 1) : 0u;  // the consumer's `done`
        const volatile double* vt = tile;
        // write chunk n back (the first NW arrays) and release its stage
        auto writeBack = [&](int n) {
            const int cn = DIR > 0 ? n : nchunks - 1 - n, st = n % NST;
#pragma unroll
            for (int a = 0; a < NW; ++a) {
                double* dst = op.arr[a] + stripBase + (size_t)cn * TILE;
                const volatile double* src = vt + ((size_t)a * RS + (size_t)st * CH) * 32;
#pragma unroll 8
                for (int i = lane; i < TILE; i += 32) dst[i] = src[i];
            }
            __syncwarp();
            if (lane == 0) stVolatileS32(&cnt[2], n + 1);
        };
        for (int m = 0; m < nsub; ++m) {
            const int n = m / NSUB;
            while (ldVolatileS32(&cnt[1]) < m + 1) {}
            SD_COMPILER_BARRIER();
            // lane LP's new values first: they are on the next strip's critical path
            if (dsOut) {
                while (peerReady < m - RB + 1)
                    asm volatile("ld.volatile.shared::cluster.s32 %0, [%1];" : "=r"(peerReady) : "r"(peerCnt) : "memory");
                if (lane < SUBS) {
                    const int u = m * SUBS + lane;
#pragma unroll
                    for (int a = 0; a < NN; ++a) {
                        double v = vt[((size_t)a * RS + slotOf(u)) * 32 + LP];
                        if (SKIP && a == 0) v = ((lpChg[m & 63] >> lane) & 1) ? -fabs(v) : fabs(v);  // sign bit = changed in this sweep
                        stAsyncU64(peerRing + (unsigned)((a * HR + (u & (HR - 1))) * 8), (unsigned long long)__double_as_longlong(v),
                                   peerBar + (unsigned)((m % RB) * 8));
                    }
                }
            } else if (lane < SUBS && q < g.nstrips - 1) {
                // global slots: only the positions the consumer polls (its needAt) -- it resets exactly those to SENT, and a
                // slot left behind with a value in it would be taken for "already written" by a later launch whose geometry
                // puts one of its own slots at the same address (the level-set window, the factor's box and the full grid
                // all use this buffer)
                // (the consumer's lane LC sits at this lane LP's column when it reads the slot: 31*SIGMA march positions later)
                const int u = m * SUBS + lane;
                const int cC = stepOf(u) - SIGMA * LP;
                if (cC >= 0 && cC < g.nx) {
#pragma unroll
                    for (int a = 0; a < NN; ++a) {
                        double v = vt[((size_t)a * RS + slotOf(u)) * 32 + LP];
                        if (SKIP && a == 0) v = ((lpChg[m & 63] >> lane) & 1) ? -fabs(v) : fabs(v);
                        stRelaxedU64(handOut + (size_t)a * ctl.planeWords + stepOf(u), (unsigned long long)__double_as_longlong(v));
                    }
                }
            }
            // chunk n-1 is complete once the solver is past the first sub-chunk of chunk n
            if (n >= 1 && written < n) { writeBack(written); ++written; }
        }
        while (written < nchunks) { writeBack(written); ++written; }
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
    if (CL > 1) clusterBarrier();
}

template <class Op, int SIGMA, int DIR, int SUBS, int CL>
static inline cudaError_t launchSweepCl(const Op& op, const Geom& g, const SweepControl& ctl, cudaStream_t stream) {
    auto kern = sweepKernel<Op, SIGMA, DIR, SUBS, CL>;
    const size_t bytes = SweepLayout<Op>::BYTES;
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
static inline cudaError_t launchSweep(const Op& op, const Geom& g, const SweepControl& ctl, cudaStream_t stream, int cl) {
    if (cl >= 8) return launchSweepCl<Op, SIGMA, DIR, SUBS, 8>(op, g, ctl, stream);
    return launchSweepCl<Op, SIGMA, DIR, SUBS, 1>(op, g, ctl, stream);
}

#endif  // __CUDACC__

}  // namespace sd
