// distpeer.cuh -- peer-memory plumbing of the y-slab PCG (one process per GPU on one NVSwitch node).
//
// Inside the PCG iteration nothing goes through NCCL: every rank owns one PeerBlock in its own HBM, exported with CUDA IPC
// at fsim_dist_init and mapped by all other ranks, and the kernels of the iteration store into their peers' blocks directly
// over NVLink:
//   * halo: the backward solve's post warp writes the boundary rows of the new search direction s straight into the
//     neighbours' ghost rows (contiguous, one double per column) while it writes s itself; the kernel's last CTA stamps
//     the neighbours' haloSeq; applyA's blocks wait for the two stamps in their prologue and read the ghost rows locally.
//   * PCG scalars: the one thread per rank that finishes a reduction (applyA: z.s; axpy: |r|_inf; forward solve: z.r) stores
//     its partial into slot [rank] of EVERY rank's block (its own included), waits until all `world` slots of its own
//     block carry this reduction's stamp and combines them in rank order -- every rank computes bit-identical alpha,
//     beta, sigma and the same stop decision, with no collective call and no extra kernel.
// Stamps are (projection epoch << 16 | iteration-derived index): they come from device state that is identical on all
// ranks, so ranks whose hosts enqueue a different number of gated (no-op) launches stay in step.  Slots are reused every
// iteration; that is race-free because the reduction kinds alternate and each one's completion on a rank is ordered
// after that rank's previous read.  Remote memory is only ever written, never read.
// A wait that exceeds DIST_SPIN_CYCLES gives up (DevCtl::distError, the solve stops): a lost peer cannot hang the GPU.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

constexpr int DIST_MAXW = 16;
constexpr long long DIST_SPIN_CYCLES = 6000000000LL;  // ~3 s at 1.9 GHz

struct PeerBlock {
    unsigned int redSeq[3][DIST_MAXW];    // stamps of the partials below: [0] z.s (applyA), [1] forward solve, [2] axpy (|r|_inf)
    double redVal[3][DIST_MAXW][2];       // [kind][source rank][sum, max]
    unsigned int haloSeq[2];              // [0] ghost row j0-1 written by rank-1, [1] ghost row j1 written by rank+1
    unsigned int pad[2];
    // double ghost[2][ghostPitch] follows at DIST_GHOST_OFF
};
constexpr size_t DIST_GHOST_OFF = 1024;
static_assert(sizeof(PeerBlock) <= DIST_GHOST_OFF, "ghost rows start after the control words");

struct PeerView {
    PeerBlock* blk[DIST_MAXW];  // every rank's block as mapped into this process ([rank] = the local one)
    int rank, world;
    int ghostPitch;             // doubles per ghost row
    unsigned int stampBase;     // projection epoch << 16
};

#ifdef __CUDACC__
__device__ __forceinline__ double* peerGhost(PeerBlock* b, int which, int ghostPitch) {
    return reinterpret_cast<double*>(reinterpret_cast<char*>(b) + DIST_GHOST_OFF) + (size_t)which * ghostPitch;
}
__device__ __forceinline__ void stReleaseSysU32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ldAcquireSysU32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stRelaxedSysF64(double* p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ldRelaxedSysF64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// spins until *p == stamp; false on timeout
__device__ __forceinline__ bool peerWait(const unsigned int* p, unsigned int stamp) {
    if (ldAcquireSysU32(p) == stamp) return true;
    const long long t0 = clock64();
    while (ldAcquireSysU32(p) != stamp)
        if (clock64() - t0 > DIST_SPIN_CYCLES) return false;
    return true;
}
// All-reduce of (sum, max) over the ranks by the ONE calling thread of each rank: deterministic (rank order).
__device__ __forceinline__ bool peerCombine(const PeerView& pv, int kind, unsigned int stamp, double& sum, double& mx) {
    for (int r = 0; r < pv.world; ++r) {
        PeerBlock* d = pv.blk[r];
        stRelaxedSysF64(&d->redVal[kind][pv.rank][0], sum);
        stRelaxedSysF64(&d->redVal[kind][pv.rank][1], mx);
        stReleaseSysU32(&d->redSeq[kind][pv.rank], stamp);
    }
    PeerBlock* me = pv.blk[pv.rank];
    double s = 0.0, m = 0.0;
    bool ok = true;
    for (int r = 0; r < pv.world; ++r) {
        ok = ok && peerWait(&me->redSeq[kind][r], stamp);
        s += ldRelaxedSysF64(&me->redVal[kind][r][0]);
        m = fmax(m, ldRelaxedSysF64(&me->redVal[kind][r][1]));
    }
    sum = s; mx = m;
    return ok;
}
#endif
