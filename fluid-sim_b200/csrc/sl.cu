// applySemiLagrangianAdvection (reference src/FluidSim2D.cpp:206-235).
//
// The reference updates mac.u / mac.v IN PLACE while later faces still interpolate from them, so its
// (serial) result depends on raster order: a footprint value is the NEW one if that face precedes the
// current face in raster order and the OLD one otherwise.  A double-buffered kernel deviates
// from it by ~2e-2 per step, 200x the parity tolerance, so the default path reproduces the raster order
// exactly:
//   * a snapshot of the component being advected provides the OLD values (no anti-dependencies),
//   * "NEW" values are read from the array being written, and ordering is enforced by a skewed wavefront:
//     one row per lane, row j+1 trails row j by K = R+3 faces, where R bounds the reach of the RK3
//     backtrace and 2 more cover the 4x4 Catmull-Rom footprint.  Inside a warp the lockstep schedule alone
//     guarantees that a NEW value is there; strips of 32 rows are chained through per-row progress counters
//     (slExactKernel).  If a backtrace ever exceeds the bound the step is redone from the snapshot with a
//     larger K.
//   * fsim_options.reserved[5] = 1 selects slExactWaitKernel, where the data validates itself instead: the
//     array being written starts out filled with a NaN payload no computation produces and a reader of a NEW
//     value polls the value itself (no counters, no fence per step).  Same bits, and measured the same speed
//     (128^2: 15.2 ms, 2048^2: 164 ms per advection) -- the fence and the counter poll are not what bounds a
//     step; kept as the simpler protocol to build the shared-memory window version on.
// fsim_options.slDoubleBuffer = 1 selects the snapshot-only variant (what Bridson's text specifies), which is
// fully parallel; it is validated against the reference patched the same way.
#include "sampling.cuh"
#include "sim.h"

namespace {

__global__ void maxAbsKernel(const double* __restrict__ u, const double* __restrict__ v, int nx, int ny, int pitch,
                             double* partials, unsigned int* counter, DevCtl* ctl) {
    __shared__ double red[32];
    double m = 0.0;
    for (int j = blockIdx.x; j <= ny; j += gridDim.x)
        for (int i = threadIdx.x; i <= nx; i += blockDim.x) {
            // only the x component matters for the skew between rows (advectFace's overflow test looks at k1x, k2x, k3x)
            if (j < ny) m = fmax(m, fabs(u[(long long)j * pitch + i]));
        }
    m = blockReduce<true>(m, red);
    gridReduceFinish<true>(m, partials, counter, red, [&](double t) { ctl->maxDisp = t; });
}

// Catmull-Rom sample of the array being advected, mixing new (in place, L2) and old (snapshot) values
template <bool EXACT>
__device__ __forceinline__ double bicubicMixed(const double* inplace, const double* snap, int pitch, int NX, int NY,
                                               double px, double py, int i0, int j0) {
    int x = (int)px, y = (int)py;
    if (x < 0 || x >= NX || y < 0 || y >= NY) return 0.0;
    double fx = px - (double)x, fy = py - (double)y;
    double fx2 = fx * fx, fx3 = fx * fx * fx, fy2 = fy * fy, fy3 = fy * fy * fy;
    double wu[4], wv[4];
    wu[0] = -0.5 * fx3 + fx2 - 0.5 * fx;
    wu[1] = 1.5 * fx3 - 2.5 * fx2 + 1;
    wu[2] = -1.5 * fx3 + 2 * fx2 + 0.5 * fx;
    wu[3] = 0.5 * fx3 - 0.5 * fx2;
    wv[0] = -0.5 * fy3 + fy2 - 0.5 * fy;
    wv[1] = 1.5 * fy3 - 2.5 * fy2 + 1;
    wv[2] = -1.5 * fy3 + 2 * fy2 + 0.5 * fy;
    wv[3] = 0.5 * fy3 - 0.5 * fy2;
    double row[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        int yy = iclampd(y - 1 + jj, 0, NY - 1);
        double a[4];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            int xx = iclampd(x - 1 + ii, 0, NX - 1);
            long long o = (long long)yy * pitch + xx;
            bool earlier = EXACT && (yy < j0 || (yy == j0 && xx < i0));
            a[ii] = earlier ? __ldcg(inplace + o) : __ldg(snap + o);
        }
        row[jj] = (wu[0] * a[0] + wu[1] * a[1]) + (wu[2] * a[2] + wu[3] * a[3]);
    }
    return (row[0] * wv[0] + row[2] * wv[2]) + (row[1] * wv[1] + row[3] * wv[3]);
}

struct SlArgs {
    const double* inplaceU; const double* inplaceV;  // arrays as they currently are
    const double* snapU; const double* snapV;        // pre-advection copies
    int nx, ny, pitch;
    double dx, dt;
};

// velocity at (x,y) as seen by face (i0,j0) of component COMP
template <int COMP, bool EXACT, bool DB>
__device__ __forceinline__ void velAt(const SlArgs& a, double x, double y, int i0, int j0, double& vx, double& vy) {
    double gx = x / a.dx, gy = y / a.dx;
    double ux = amlClamp(gx, 1e-6, (double)(a.nx - 1) - 1e-6), uy = amlClamp(gy - 0.5, 1e-6, (double)(a.ny - 1) - 1e-6);
    double wx = amlClamp(gx - 0.5, 1e-6, (double)(a.nx - 1) - 1e-6), wy = amlClamp(gy, 1e-6, (double)(a.ny - 1) - 1e-6);
    if (COMP == 0) {
        vx = bicubicMixed<EXACT>(a.inplaceU, a.snapU, a.pitch, a.nx + 1, a.ny, ux, uy, i0, j0);
        // v is untouched during the u pass (double-buffered: still the snapshot)
        vy = bicubicMixed<false>(nullptr, DB ? a.snapV : a.inplaceV, a.pitch, a.nx, a.ny + 1, wx, wy, 0, 0);
    } else {
        // u is completely new during the v pass of the in-place algorithm; old in the double-buffered one
        vx = bicubicMixed<false>(nullptr, DB ? a.snapU : a.inplaceU, a.pitch, a.nx + 1, a.ny, ux, uy, 0, 0);
        vy = bicubicMixed<EXACT>(a.inplaceV, a.snapV, a.pitch, a.nx, a.ny + 1, wx, wy, i0, j0);
    }
}

template <int COMP, bool EXACT, bool DB>
__device__ __forceinline__ double advectFace(const SlArgs& a, int i, int j, double reachCells, int* overflow) {
    double x = COMP == 0 ? a.dx * (double)i : a.dx * ((double)i + 0.5);
    double y = COMP == 0 ? a.dx * ((double)j + 0.5) : a.dx * (double)j;
    double k1x, k1y, k2x, k2y, k3x, k3y;
    velAt<COMP, EXACT, DB>(a, x, y, i, j, k1x, k1y);
    velAt<COMP, EXACT, DB>(a, x - 0.5 * a.dt * k1x, y - 0.5 * a.dt * k1y, i, j, k2x, k2y);
    velAt<COMP, EXACT, DB>(a, x - 0.75 * a.dt * k2x, y - 0.75 * a.dt * k2y, i, j, k3x, k3y);
    double nxp = x - ((2. / 9.) * a.dt * k1x + (3. / 9.) * a.dt * k2x + (4. / 9.) * a.dt * k3x);
    double nyp = y - ((2. / 9.) * a.dt * k1y + (3. / 9.) * a.dt * k2y + (4. / 9.) * a.dt * k3y);
    if (EXACT) {
        double m = fmax(fmax(fabs(k1x), fabs(k2x)), fabs(k3x)) * a.dt / a.dx;
        if (!(m <= reachCells)) atomicOr(overflow, 1);
    }
    clampPos(a.nx, a.ny, a.dx, nxp, nyp);
    double vx, vy;
    // final lookup of the advected component only (:218, :231)
    double gx = nxp / a.dx, gy = nyp / a.dx;
    if (COMP == 0) {
        double ux = amlClamp(gx, 1e-6, (double)(a.nx - 1) - 1e-6), uy = amlClamp(gy - 0.5, 1e-6, (double)(a.ny - 1) - 1e-6);
        vx = bicubicMixed<EXACT>(a.inplaceU, a.snapU, a.pitch, a.nx + 1, a.ny, ux, uy, i, j);
        return vx;
    } else {
        double wx = amlClamp(gx - 0.5, 1e-6, (double)(a.nx - 1) - 1e-6), wy = amlClamp(gy, 1e-6, (double)(a.ny - 1) - 1e-6);
        vy = bicubicMixed<EXACT>(a.inplaceV, a.snapV, a.pitch, a.nx, a.ny + 1, wx, wy, i, j);
        return vy;
    }
}

__device__ __forceinline__ int ldRelaxedI32(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stRelaxedI32(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// exact raster-order semantics: skewed wavefront, one row per lane
template <int COMP>
__global__ void __launch_bounds__(32) slExactKernel(SlArgs a, double* dst, int K, double reachCells, int* ticket,
                                                    int* finished, int* progress, int* overflow) {
    const int lane = threadIdx.x;
    int strip = 0;
    if (lane == 0) strip = atomicAdd(ticket, 1);
    strip = __shfl_sync(0xffffffffu, strip, 0);
    const int NXf = COMP == 0 ? a.nx + 1 : a.nx, NYf = COMP == 0 ? a.ny : a.ny + 1;
    const int j = strip * 32 + lane;
    const bool valid = j < NYf;
    const int nsteps = NXf + 31 * K;
    for (int s = 0; s < nsteps; ++s) {
        const int c = s - lane * K;
        const bool active = valid && c >= 0 && c < NXf;
        if (lane == 0 && strip > 0 && active) {
            int need = min(c + K, NXf);
            while (ldRelaxedI32(progress + j - 1) < need) {}
        }
        __syncwarp();
        if (active) {
            double val = advectFace<COMP, true, false>(a, c, j, reachCells, overflow);
            __stcg(dst + (long long)j * a.pitch + c, val);
        }
        __threadfence();
        __syncwarp();
        if (active) stRelaxedI32(progress + j, c + 1);
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        int nstrips = (NYf + 31) / 32;
        if (atomicAdd(finished, 1) == nstrips - 1) { *finished = 0; *ticket = 0; }
    }
}

// ---- self-validating variant ---------------------------------------------------------------------------------------
constexpr unsigned long long SL_SENT = 0xFFFFFFFFFFFFFFFFull;  // fill pattern of cudaMemset(0xFF): a NaN nothing computes

__device__ __forceinline__ unsigned long long ldRelaxedU64(const double* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct SlWait {
    int K, stripRow0;  // skew; first row of this warp's strip (rows above it belong to strips that never wait for us)
    int* flag;         // 1: a backtrace left the dependency cone (redo with a larger K), 2: a wait timed out (error)
    bool dead;
};

// bicubicMixed<true> with the NEW values polled until valid.  Same arithmetic, same order.
__device__ __forceinline__ double bicubicWait(const double* inplace, const double* snap, int pitch, int NX, int NY,
                                              double px, double py, int i0, int j0, SlWait& w) {
    int x = (int)px, y = (int)py;
    if (x < 0 || x >= NX || y < 0 || y >= NY) return 0.0;
    double fx = px - (double)x, fy = py - (double)y;
    double fx2 = fx * fx, fx3 = fx * fx * fx, fy2 = fy * fy, fy3 = fy * fy * fy;
    double wu[4], wv[4];
    wu[0] = -0.5 * fx3 + fx2 - 0.5 * fx;
    wu[1] = 1.5 * fx3 - 2.5 * fx2 + 1;
    wu[2] = -1.5 * fx3 + 2 * fx2 + 0.5 * fx;
    wu[3] = 0.5 * fx3 - 0.5 * fx2;
    wv[0] = -0.5 * fy3 + fy2 - 0.5 * fy;
    wv[1] = 1.5 * fy3 - 2.5 * fy2 + 1;
    wv[2] = -1.5 * fy3 + 2 * fy2 + 0.5 * fy;
    wv[3] = 0.5 * fy3 - 0.5 * fy2;
    double a[4][4];
    unsigned long long b[4][4];
    unsigned newMask = 0;
    bool outside = false;
    // phase 1: all 16 loads are issued back to back, branch-free (one load instruction, the address selects NEW or OLD)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        int yy = iclampd(y - 1 + jj, 0, NY - 1);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            int xx = iclampd(x - 1 + ii, 0, NX - 1);
            long long o = (long long)yy * pitch + xx;
            bool earlier = yy < j0 || (yy == j0 && xx < i0);
            // a NEW value of this warp's own rows exists only inside the dependency cone of the lockstep schedule;
            // outside it nobody would ever write it while we wait
            bool cone = yy < w.stripRow0 || xx <= i0 + (j0 - yy) * w.K - 1;
            outside |= earlier && !cone;
            earlier &= cone;
            b[jj][ii] = ldRelaxedU64(earlier ? inplace + o : snap + o);
            newMask |= (earlier ? 1u : 0u) << (jj * 4 + ii);
        }
    }
    if (outside) atomicOr(w.flag, 1);
    // phase 2: which NEW values are still the marker
    unsigned pend = 0;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            a[jj][ii] = __longlong_as_double((long long)b[jj][ii]);
            pend |= (b[jj][ii] == SL_SENT ? 1u : 0u) << (jj * 4 + ii);
        }
    pend &= newMask;
    if (pend && !w.dead) {  // values of the strip above that are still on their way
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            int yy = iclampd(y - 1 + jj, 0, NY - 1);
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                if (!(pend & (1u << (jj * 4 + ii))) || w.dead) continue;
                int xx = iclampd(x - 1 + ii, 0, NX - 1);
                const double* q = inplace + (long long)yy * pitch + xx;
                unsigned long long bits;
                unsigned spins = 0;
                while ((bits = ldRelaxedU64(q)) == SL_SENT) {
                    if ((++spins & 255u) == 0 && (spins > (1u << 21) || ldRelaxedI32(w.flag) == 2)) {
                        atomicExch(w.flag, 2);
                        w.dead = true;
                        break;
                    }
                }
                a[jj][ii] = __longlong_as_double((long long)bits);
            }
        }
    }
    double row[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) row[jj] = (wu[0] * a[jj][0] + wu[1] * a[jj][1]) + (wu[2] * a[jj][2] + wu[3] * a[jj][3]);
    return (row[0] * wv[0] + row[2] * wv[2]) + (row[1] * wv[1] + row[3] * wv[3]);
}

template <int COMP>
__device__ __forceinline__ void velAtWait(const SlArgs& a, double x, double y, int i0, int j0, SlWait& w, double& vx, double& vy) {
    double gx = x / a.dx, gy = y / a.dx;
    double ux = amlClamp(gx, 1e-6, (double)(a.nx - 1) - 1e-6), uy = amlClamp(gy - 0.5, 1e-6, (double)(a.ny - 1) - 1e-6);
    double wx = amlClamp(gx - 0.5, 1e-6, (double)(a.nx - 1) - 1e-6), wy = amlClamp(gy, 1e-6, (double)(a.ny - 1) - 1e-6);
    if (COMP == 0) {
        vx = bicubicWait(a.inplaceU, a.snapU, a.pitch, a.nx + 1, a.ny, ux, uy, i0, j0, w);
        vy = bicubicMixed<false>(nullptr, a.snapV, a.pitch, a.nx, a.ny + 1, wx, wy, 0, 0);  // v: untouched during the u pass
    } else {
        vx = bicubicMixed<false>(nullptr, a.inplaceU, a.pitch, a.nx + 1, a.ny, ux, uy, 0, 0);  // u: completely new
        vy = bicubicWait(a.inplaceV, a.snapV, a.pitch, a.nx, a.ny + 1, wx, wy, i0, j0, w);
    }
}

template <int COMP>
__device__ __forceinline__ double advectFaceWait(const SlArgs& a, int i, int j, double reachCells, SlWait& w) {
    double x = COMP == 0 ? a.dx * (double)i : a.dx * ((double)i + 0.5);
    double y = COMP == 0 ? a.dx * ((double)j + 0.5) : a.dx * (double)j;
    double k1x, k1y, k2x, k2y, k3x, k3y;
    velAtWait<COMP>(a, x, y, i, j, w, k1x, k1y);
    velAtWait<COMP>(a, x - 0.5 * a.dt * k1x, y - 0.5 * a.dt * k1y, i, j, w, k2x, k2y);
    velAtWait<COMP>(a, x - 0.75 * a.dt * k2x, y - 0.75 * a.dt * k2y, i, j, w, k3x, k3y);
    double nxp = x - ((2. / 9.) * a.dt * k1x + (3. / 9.) * a.dt * k2x + (4. / 9.) * a.dt * k3x);
    double nyp = y - ((2. / 9.) * a.dt * k1y + (3. / 9.) * a.dt * k2y + (4. / 9.) * a.dt * k3y);
    double m = fmax(fmax(fabs(k1x), fabs(k2x)), fabs(k3x)) * a.dt / a.dx;
    if (!(m <= reachCells)) atomicOr(w.flag, 1);
    clampPos(a.nx, a.ny, a.dx, nxp, nyp);
    double gx = nxp / a.dx, gy = nyp / a.dx;  // final lookup of the advected component only (:218, :231)
    if (COMP == 0) {
        double ux = amlClamp(gx, 1e-6, (double)(a.nx - 1) - 1e-6), uy = amlClamp(gy - 0.5, 1e-6, (double)(a.ny - 1) - 1e-6);
        return bicubicWait(a.inplaceU, a.snapU, a.pitch, a.nx + 1, a.ny, ux, uy, i, j, w);
    } else {
        double wx = amlClamp(gx - 0.5, 1e-6, (double)(a.nx - 1) - 1e-6), wy = amlClamp(gy, 1e-6, (double)(a.ny - 1) - 1e-6);
        return bicubicWait(a.inplaceV, a.snapV, a.pitch, a.nx, a.ny + 1, wx, wy, i, j, w);
    }
}

// exact raster-order semantics, self-validating data: dst (= the in-place array) starts out as SL_SENT everywhere
template <int COMP>
__global__ void __launch_bounds__(32) slExactWaitKernel(SlArgs a, double* dst, int K, double reachCells, int* ticket,
                                                        int* finished, int* flag) {
    const int lane = threadIdx.x;
    int strip = 0;
    if (lane == 0) strip = atomicAdd(ticket, 1);  // march order: the strip above is resident or done
    strip = __shfl_sync(0xffffffffu, strip, 0);
    const int NXf = COMP == 0 ? a.nx + 1 : a.nx, NYf = COMP == 0 ? a.ny : a.ny + 1;
    const int j = strip * 32 + lane;
    const bool valid = j < NYf;
    const int nsteps = NXf + 31 * K;
    SlWait w{K, strip * 32, flag, false};
    for (int s = 0; s < nsteps; ++s) {
        const int c = s - lane * K;
        if (valid && c >= 0 && c < NXf) {
            double val = advectFaceWait<COMP>(a, c, j, reachCells, w);
            if (__double_as_longlong(val) == (long long)SL_SENT) val = __longlong_as_double(0x7FF8000000000000ll);  // a NaN, but not the marker
            __stcg(dst + (long long)j * a.pitch + c, val);
        }
        if (__any_sync(0xffffffffu, w.dead)) break;  // (also the step's warp barrier: orders the store before the next reads)
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        int nstrips = (NYf + 31) / 32;
        if (atomicAdd(finished, 1) == nstrips - 1) { *finished = 0; *ticket = 0; }
    }
}



// ---- small grids: the whole in-place component in shared memory ------------------------------------------------------
// When the component being advected fits one SM's shared memory ((NX * NY * 8 <= SL_SMALL_BYTES: up to ~158^2, which covers
// the reference's own demo scene, 128 x 128) one CTA runs every strip, one warp each, on ONE array in shared memory: a slot
// holds the NEW value exactly when its face has been visited -- the reference's in-place array -- so the 4x4 footprints are
// plain shared-memory reads (~30 cycles) instead of L2 round trips (~700), and the snapshot is not needed for OLD values:
// the skew guarantees that a face later in raster order has not been visited yet when it is read (row j+1 trails row j by
// K = reach + 3 faces, the footprint's leftmost column is i - reach - 1).  Same arithmetic, same order, same bits as
// slExactKernel; per wavefront step ~4 dependent bicubic stages of registers + LDS.
constexpr size_t SL_SMALL_BYTES = 200 * 1024;

__device__ __forceinline__ double bicubicPlain(const volatile double* a, int pitch, int NX, int NY, double px, double py) {
    int x = (int)px, y = (int)py;
    if (x < 0 || x >= NX || y < 0 || y >= NY) return 0.0;
    double fx = px - (double)x, fy = py - (double)y;
    double fx2 = fx * fx, fx3 = fx * fx * fx, fy2 = fy * fy, fy3 = fy * fy * fy;
    double wu[4], wv[4];
    wu[0] = -0.5 * fx3 + fx2 - 0.5 * fx;
    wu[1] = 1.5 * fx3 - 2.5 * fx2 + 1;
    wu[2] = -1.5 * fx3 + 2 * fx2 + 0.5 * fx;
    wu[3] = 0.5 * fx3 - 0.5 * fx2;
    wv[0] = -0.5 * fy3 + fy2 - 0.5 * fy;
    wv[1] = 1.5 * fy3 - 2.5 * fy2 + 1;
    wv[2] = -1.5 * fy3 + 2 * fy2 + 0.5 * fy;
    wv[3] = 0.5 * fy3 - 0.5 * fy2;
    double v[4][4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int yy = iclampd(y - 1 + jj, 0, NY - 1);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) v[jj][ii] = a[yy * pitch + iclampd(x - 1 + ii, 0, NX - 1)];
    }
    double row[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) row[jj] = (wu[0] * v[jj][0] + wu[1] * v[jj][1]) + (wu[2] * v[jj][2] + wu[3] * v[jj][3]);
    return (row[0] * wv[0] + row[2] * wv[2]) + (row[1] * wv[1] + row[3] * wv[3]);
}

// velocity at (x, y): the advected component from the shared-memory array, the other one from global memory (constant
// during this pass: v is untouched while u is advected, u is completely new while v is)
template <int COMP>
__device__ __forceinline__ void velAtSmall(const SlArgs& a, const volatile double* arr, int NXf, double x, double y, double& vx, double& vy) {
    double gx = x / a.dx, gy = y / a.dx;
    double ux = amlClamp(gx, 1e-6, (double)(a.nx - 1) - 1e-6), uy = amlClamp(gy - 0.5, 1e-6, (double)(a.ny - 1) - 1e-6);
    double wx = amlClamp(gx - 0.5, 1e-6, (double)(a.nx - 1) - 1e-6), wy = amlClamp(gy, 1e-6, (double)(a.ny - 1) - 1e-6);
    if (COMP == 0) {
        vx = bicubicPlain(arr, NXf, a.nx + 1, a.ny, ux, uy);
        vy = bicubicMixed<false>(nullptr, a.inplaceV, a.pitch, a.nx, a.ny + 1, wx, wy, 0, 0);
    } else {
        vx = bicubicMixed<false>(nullptr, a.inplaceU, a.pitch, a.nx + 1, a.ny, ux, uy, 0, 0);
        vy = bicubicPlain(arr, NXf, a.nx, a.ny + 1, wx, wy);
    }
}

template <int COMP>
__global__ void __launch_bounds__(256) slExactSmallKernel(SlArgs a, double* dst, int K, double reachCells, int* overflow) {
    extern __shared__ __align__(16) unsigned char slSmem[];
    const int NXf = COMP == 0 ? a.nx + 1 : a.nx, NYf = COMP == 0 ? a.ny : a.ny + 1;
    volatile double* arr = reinterpret_cast<volatile double*>(slSmem);          // [NYf][NXf], in place
    volatile int* progress = reinterpret_cast<volatile int*>(arr + (size_t)NXf * NYf);  // [NYf] faces done per row
    for (int q = threadIdx.x; q < NXf * NYf; q += blockDim.x) {
        const int jj = q / NXf, ii = q - jj * NXf;
        arr[q] = dst[(long long)jj * a.pitch + ii];
    }
    for (int q = threadIdx.x; q < NYf; q += blockDim.x) progress[q] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, strip = threadIdx.x >> 5;
    const int j = strip * 32 + lane;
    const bool valid = j < NYf;
    const int nsteps = NXf + 31 * K;
    for (int s = 0; s < nsteps; ++s) {
        const int c = s - lane * K;
        const bool active = valid && c >= 0 && c < NXf;
        if (lane == 0 && strip > 0 && active) {
            const int need = min(c + K, NXf);
            while (progress[j - 1] < need) {}
            __threadfence_block();
        }
        __syncwarp();
        if (active) {
            double x = COMP == 0 ? a.dx * (double)c : a.dx * ((double)c + 0.5);
            double y = COMP == 0 ? a.dx * ((double)j + 0.5) : a.dx * (double)j;
            double k1x, k1y, k2x, k2y, k3x, k3y;
            velAtSmall<COMP>(a, arr, NXf, x, y, k1x, k1y);
            velAtSmall<COMP>(a, arr, NXf, x - 0.5 * a.dt * k1x, y - 0.5 * a.dt * k1y, k2x, k2y);
            velAtSmall<COMP>(a, arr, NXf, x - 0.75 * a.dt * k2x, y - 0.75 * a.dt * k2y, k3x, k3y);
            double nxp = x - ((2. / 9.) * a.dt * k1x + (3. / 9.) * a.dt * k2x + (4. / 9.) * a.dt * k3x);
            double nyp = y - ((2. / 9.) * a.dt * k1y + (3. / 9.) * a.dt * k2y + (4. / 9.) * a.dt * k3y);
            const double m = fmax(fmax(fabs(k1x), fabs(k2x)), fabs(k3x)) * a.dt / a.dx;
            if (!(m <= reachCells)) atomicOr(overflow, 1);
            clampPos(a.nx, a.ny, a.dx, nxp, nyp);
            const double gx = nxp / a.dx, gy = nyp / a.dx;  // final lookup of the advected component only (:218, :231)
            double val;
            if (COMP == 0) {
                double ux = amlClamp(gx, 1e-6, (double)(a.nx - 1) - 1e-6), uy = amlClamp(gy - 0.5, 1e-6, (double)(a.ny - 1) - 1e-6);
                val = bicubicPlain(arr, NXf, a.nx + 1, a.ny, ux, uy);
            } else {
                double wx = amlClamp(gx - 0.5, 1e-6, (double)(a.nx - 1) - 1e-6), wy = amlClamp(gy, 1e-6, (double)(a.ny - 1) - 1e-6);
                val = bicubicPlain(arr, NXf, a.nx, a.ny + 1, wx, wy);
            }
            arr[j * NXf + c] = val;
        }
        __threadfence_block();  // the value is in shared memory before the row's counter moves
        __syncwarp();
        if (active) progress[j] = c + 1;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < NXf * NYf; q += blockDim.x) {
        const int jj = q / NXf, ii = q - jj * NXf;
        dst[(long long)jj * a.pitch + ii] = arr[q];
    }
}

template <int COMP>
__global__ void slDoubleBufferKernel(SlArgs a, double* dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    const int NXf = COMP == 0 ? a.nx + 1 : a.nx, NYf = COMP == 0 ? a.ny : a.ny + 1;
    if (i >= NXf || j >= NYf) return;
    dst[(long long)j * a.pitch + i] = advectFace<COMP, false, true>(a, i, j, 0.0, nullptr);
}

}  // namespace

int stageApplySemiLagrangianAdvection(Sim* s) {
    const Frame& f = s->fr;
    if (!s->slU) { fsim_set_error("semi-Lagrangian advection needs a handle created in FS_SEMILAGRANGIAN mode"); return FSIM_E_STATE; }
    size_t bytes = f.elems * sizeof(double);
    CUDA_TRY(cudaMemcpyAsync(s->slU - f.org, s->u - f.org, bytes, cudaMemcpyDeviceToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->slV - f.org, s->v - f.org, bytes, cudaMemcpyDeviceToDevice, s->stream));
    SlArgs a{s->u, s->v, s->slU, s->slV, s->nx, s->ny, f.pitch, s->dx, s->dt};
    if (s->opt.slDoubleBuffer) {
        dim3 blk(32, 4), grd((s->nx + 1 + 31) / 32, (s->ny + 1 + 3) / 4);
        slDoubleBufferKernel<0><<<grd, blk, 0, s->stream>>>(a, s->u);
        slDoubleBufferKernel<1><<<grd, blk, 0, s->stream>>>(a, s->v);
        s->launches += 2;
        CUDA_TRY(cudaGetLastError());
        return FSIM_OK;
    }
    maxAbsKernel<<<296, 256, 0, s->stream>>>(s->u, s->v, s->nx, s->ny, f.pitch, s->partials, &s->counters[5], s->ctl);
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaMemcpyAsync(&s->hctl->maxDisp, &s->ctl->maxDisp, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    // Reach of a backtrace along x in cells.  The stage velocities are Catmull-Rom samples of u (old values, and new ones that
    // are samples themselves): at most ~1.13^2 of the largest |u| on the grid.  A face that does exceed the bound raises the
    // overflow flag and the advection is redone from the snapshot with twice the reach, so this only has to be right nearly
    // always -- every cell of reach costs sizeY wavefront steps.
    double reach = ceil(1.3 * s->hctl->maxDisp * s->dt / s->dx);
    if (reach < 1.0) reach = 1.0;
    int* overflow = &s->ctl->slOverflow;
    const bool counters = s->opt.reserved[5] != 1;  // default: per-row progress counters; 1: self-validating data
    const int NYu = s->ny, NYv = s->ny + 1;
    // small grids (the reference's demo scene): one CTA, the component in shared memory (slExactSmallKernel)
    static int noSmall = -1;
    if (noSmall < 0) { const char* e = getenv("FSIM_SL_NO_SMALL"); noSmall = e && atoi(e) ? 1 : 0; }
    const size_t smallBytes = (size_t)(s->nx + 1) * (s->ny + 1) * sizeof(double) + (size_t)(s->ny + 1) * sizeof(int) + 16;
    const bool small = counters && !noSmall && s->opt.reserved[5] != 2 && smallBytes <= SL_SMALL_BYTES && s->ny + 1 <= 256;
    if (small) {
        static bool attrSet[16] = {};
        if (!attrSet[s->device & 15]) {
            CUDA_TRY(cudaFuncSetAttribute(slExactSmallKernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL_SMALL_BYTES));
            CUDA_TRY(cudaFuncSetAttribute(slExactSmallKernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL_SMALL_BYTES));
            attrSet[s->device & 15] = true;
        }
    }
    for (int attempt = 0; attempt < 6; ++attempt) {
        int K = (int)reach + 3;  // reach of the backtrace + 2 footprint cells + 1
        CUDA_TRY(cudaMemsetAsync(overflow, 0, sizeof(int), s->stream));
        if (small) {
            slExactSmallKernel<0><<<1, 32 * ((NYu + 31) / 32), smallBytes, s->stream>>>(a, s->u, K, reach, overflow);
            slExactSmallKernel<1><<<1, 32 * ((NYv + 31) / 32), smallBytes, s->stream>>>(a, s->v, K, reach, overflow);
        } else if (counters) {
            CUDA_TRY(cudaMemsetAsync(s->slProgress, 0, sizeof(int) * (f.H + 64), s->stream));
            slExactKernel<0><<<(NYu + 31) / 32, 32, 0, s->stream>>>(a, s->u, K, reach, s->wfTicket + 2, s->wfTicket + 3,
                                                                    s->slProgress, overflow);
            CUDA_TRY(cudaMemsetAsync(s->slProgress, 0, sizeof(int) * (f.H + 64), s->stream));
            slExactKernel<1><<<(NYv + 31) / 32, 32, 0, s->stream>>>(a, s->v, K, reach, s->wfTicket + 2, s->wfTicket + 3,
                                                                    s->slProgress, overflow);
        } else {
            // the faces of the component about to be advected become "not written yet" (the frame's halo stays zero)
            CUDA_TRY(cudaMemset2DAsync(s->u, f.pitch * 8, 0xFF, (size_t)(s->nx + 1) * 8, NYu, s->stream));
            slExactWaitKernel<0><<<(NYu + 31) / 32, 32, 0, s->stream>>>(a, s->u, K, reach, s->wfTicket + 2, s->wfTicket + 3, overflow);
            CUDA_TRY(cudaMemset2DAsync(s->v, f.pitch * 8, 0xFF, (size_t)s->nx * 8, NYv, s->stream));
            slExactWaitKernel<1><<<(NYv + 31) / 32, 32, 0, s->stream>>>(a, s->v, K, reach, s->wfTicket + 2, s->wfTicket + 3, overflow);
        }
        s->launches += 2;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&s->hPcgFlags[2], overflow, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (!s->hPcgFlags[2]) return FSIM_OK;
        // a backtrace outran the dependency skew (or a wait timed out): restore, then redo with twice the reach
        CUDA_TRY(cudaMemcpyAsync(s->u - f.org, s->slU - f.org, bytes, cudaMemcpyDeviceToDevice, s->stream));
        CUDA_TRY(cudaMemcpyAsync(s->v - f.org, s->slV - f.org, bytes, cudaMemcpyDeviceToDevice, s->stream));
        if (s->hPcgFlags[2] & 2) {
            fsim_set_error("semi-Lagrangian advection: a wait for a value of the strip above timed out");
            return FSIM_E_STATE;
        }
        reach *= 2.0;
    }
    fsim_set_error("semi-Lagrangian backtrace exceeded the dependency skew (velocity blow-up?)");
    return FSIM_E_STATE;
}
