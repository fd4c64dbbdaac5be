// Particle stages: cell sort, P2G (transferVelocityToGrid, reference src/FluidSim2D.cpp:144-204),
// G2P (updateParticleVelocities, :552-568) and RK3 particle advection (applyAdvection, :570-604).
//
// P2G is a deterministic gather: particles are counting-sorted by cell (stable: ascending particle index
// inside a cell), and every face sums the particles of the <= 6 cells whose bilinear footprint can reach
// it, in a fixed order.  No floating-point atomics; the result is bitwise reproducible run to run.
#include "sampling.cuh"
#include "sim.h"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

__global__ void countKernel(const double2* __restrict__ pos, size_t np, double dx, int nx, int ny,
                            uint32_t* __restrict__ pcell, uint32_t* __restrict__ count) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= np) return;
    double2 p = pos[e];
    int cx = (int)(p.x / dx), cy = (int)(p.y / dx);  // src/FluidSim2D.cpp:767
    cx = iclampd(cx, 0, nx - 1);
    cy = iclampd(cy, 0, ny - 1);
    uint32_t c = (uint32_t)cy * nx + cx;
    pcell[e] = c;
    atomicAdd(&count[c], 1u);
}

__device__ __forceinline__ uint32_t blockExclusiveScan(uint32_t v, uint32_t* smem, uint32_t& total) {
    // inclusive warp scan
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) smem[w] = x;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < (blockDim.x >> 5) ? smem[lane] : 0;
        uint32_t xs = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, xs, o);
            if (lane >= o) xs += y;
        }
        smem[lane] = xs - s;  // exclusive warp offsets
        if (lane == 31) smem[32] = xs;
    }
    __syncthreads();
    total = smem[32];
    uint32_t r = x - v + smem[w];
    __syncthreads();
    return r;
}

__global__ void scanReduceKernel(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ blockSums) {
    __shared__ uint32_t sm[33];
    size_t base = (size_t)blockIdx.x * SCAN_CHUNK;
    uint32_t s = 0;
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        size_t idx = base + (size_t)threadIdx.x * SCAN_ITEMS + k;
        if (idx < n) s += in[idx];
    }
    uint32_t total;
    blockExclusiveScan(s, sm, total);
    if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}

__global__ void scanBlockSumsKernel(uint32_t* blockSums, int nb) {
    __shared__ uint32_t sm[33];
    uint32_t carry = 0;
    for (int base = 0; base < nb; base += blockDim.x) {
        int idx = base + threadIdx.x;
        uint32_t v = idx < nb ? blockSums[idx] : 0;
        uint32_t total;
        uint32_t ex = blockExclusiveScan(v, sm, total);
        if (idx < nb) blockSums[idx] = ex + carry;
        carry += total;
    }
}

__global__ void scanFinalKernel(const uint32_t* __restrict__ in, size_t n, const uint32_t* __restrict__ blockSums,
                                uint32_t* __restrict__ out, uint32_t grandTotal) {
    __shared__ uint32_t sm[33];
    size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = blockExclusiveScan(s, sm, total) + blockSums[blockIdx.x];
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = grandTotal;
}

__global__ void fillKernel(const uint32_t* __restrict__ pcell, size_t np, const uint32_t* __restrict__ cellStart,
                           uint32_t* __restrict__ cursor, uint32_t* __restrict__ sortedIdx) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= np) return;
    uint32_t c = pcell[e];
    uint32_t slot = atomicAdd(&cursor[c], 1u);
    sortedIdx[cellStart[c] + slot] = (uint32_t)e;
}

// restore ascending particle index inside every cell (the atomics above fill in arbitrary order)
__global__ void sortCellsKernel(const uint32_t* __restrict__ cellStart, size_t ncells, uint32_t* __restrict__ sortedIdx) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    uint32_t b = cellStart[c], e = cellStart[c + 1];
    for (uint32_t a = b + 1; a < e; ++a) {
        uint32_t key = sortedIdx[a];
        uint32_t k = a;
        while (k > b && sortedIdx[k - 1] > key) { sortedIdx[k] = sortedIdx[k - 1]; --k; }
        sortedIdx[k] = key;
    }
}

// ----------------------------------------------------------------------------------------------- P2G
// COMP = 0: u faces sampled at (x/dx, y/dx-0.5); COMP = 1: v faces at (x/dx-0.5, y/dx) (:156-161).
template <int COMP>
__global__ void p2gGatherKernel(const double2* __restrict__ pos, const double2* __restrict__ vel,
                                const uint32_t* __restrict__ cellStart, const uint32_t* __restrict__ sortedIdx,
                                int nx, int ny, int pitch, double dx, double* __restrict__ out,
                                uint8_t* __restrict__ unknown, int* __restrict__ anyKnown) {
    const int NX = COMP == 0 ? nx + 1 : nx, NY = COMP == 0 ? ny : ny + 1;
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= NX || j >= NY) return;
    // cells whose particles can reach this face
    int cx0 = i - 1, cx1 = COMP == 0 ? i : i + 1;
    int cy0 = j - 1, cy1 = COMP == 0 ? j + 1 : j;
    cx0 = max(cx0, 0); cx1 = min(cx1, nx - 1);
    cy0 = max(cy0, 0); cy1 = min(cy1, ny - 1);
    double sum = 0.0, wsum = 0.0;
    for (int cy = cy0; cy <= cy1; ++cy) {
        // the cells cx0..cx1 of one row are contiguous in the sorted order
        uint32_t kb = cellStart[(size_t)cy * nx + cx0], ke = cellStart[(size_t)cy * nx + cx1 + 1];
        for (uint32_t k = kb; k < ke; ++k) {
            uint32_t e = sortedIdx[k];
            double2 p = pos[e];
            double val = COMP == 0 ? vel[e].x : vel[e].y;
            double sx = COMP == 0 ? p.x / dx : p.x / dx - 0.5;
            double sy = COMP == 0 ? p.y / dx - 0.5 : p.y / dx;
            Bilinear b = bilinearAt(sx, sy, NX, NY);
            // the four read-modify-writes of linearDistribute, in its order (include/Array2D.h:372-375)
            if (b.x1 == i && b.y1 == j) { double w = (1 - b.fx) * (1 - b.fy); sum += w * val; wsum += w; }
            if (b.x2 == i && b.y1 == j) { double w = b.fx * (1 - b.fy); sum += w * val; wsum += w; }
            if (b.x1 == i && b.y2 == j) { double w = (1 - b.fx) * b.fy; sum += w * val; wsum += w; }
            if (b.x2 == i && b.y2 == j) { double w = b.fx * b.fy; sum += w * val; wsum += w; }
        }
    }
    double r = wsum > 0 ? sum / wsum : sum;  // :175-184
    bool unk = (r == 0.0);                   // :192-201: exact zero means "no information"
    out[(long long)j * pitch + i] = r;
    unknown[(long long)j * pitch + i] = unk ? 1 : 0;
    if (!unk && *anyKnown == 0) *anyKnown = 1;
}

// ----------------------------------------------------------------------------------------------- G2P
__device__ __forceinline__ double gatherOne(const double* a, int pitch, const Bilinear& b) {
    double value = 0.0;
    value += a[(long long)b.y1 * pitch + b.x1] * (1 - b.fx) * (1 - b.fy);
    value += a[(long long)b.y1 * pitch + b.x2] * b.fx * (1 - b.fy);
    value += a[(long long)b.y2 * pitch + b.x1] * (1 - b.fx) * b.fy;
    value += a[(long long)b.y2 * pitch + b.x2] * b.fx * b.fy;
    return value;
}
// gather of (new - old), the reference's uDiff/vDiff arrays (:554-555) evaluated per corner
__device__ __forceinline__ double gatherDiff(const double* an, const double* ao, int pitch, const Bilinear& b) {
    double value = 0.0;
    long long o11 = (long long)b.y1 * pitch + b.x1, o21 = (long long)b.y1 * pitch + b.x2;
    long long o12 = (long long)b.y2 * pitch + b.x1, o22 = (long long)b.y2 * pitch + b.x2;
    value += (an[o11] - ao[o11]) * (1 - b.fx) * (1 - b.fy);
    value += (an[o21] - ao[o21]) * b.fx * (1 - b.fy);
    value += (an[o12] - ao[o12]) * (1 - b.fx) * b.fy;
    value += (an[o22] - ao[o22]) * b.fx * b.fy;
    return value;
}

__global__ void g2pKernel(const double2* __restrict__ pos, double2* __restrict__ vel, size_t np,
                          const double* __restrict__ u, const double* __restrict__ v,
                          const double* __restrict__ nu, const double* __restrict__ nv,
                          int nx, int ny, int pitch, double dx, double alpha) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= np) return;
    double2 p = pos[e];
    Bilinear bu = bilinearAt(p.x / dx, p.y / dx - 0.5, nx + 1, ny);
    Bilinear bv = bilinearAt(p.x / dx - 0.5, p.y / dx, nx, ny + 1);
    double picx = gatherOne(nu, pitch, bu), picy = gatherOne(nv, pitch, bv);
    double2 w = vel[e];
    double flipx = w.x + gatherDiff(nu, u, pitch, bu);
    double flipy = w.y + gatherDiff(nv, v, pitch, bv);
    vel[e] = make_double2(alpha * picx + (1 - alpha) * flipx, alpha * picy + (1 - alpha) * flipy);
}

// ----------------------------------------------------------------------------------------------- advection
__global__ void advectKernel(double2* __restrict__ pos, const double2* __restrict__ vel, size_t np, GridView g,
                             double dt, DevCtl* ctl) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double c = 0.0;
    int nan = 0;
    if (e < np) {
        double2 p = pos[e];
        double2 w = vel[e];
        c = (w.x + w.y) * dt / g.dx;  // :576
        nan = (isnan(p.x) || isnan(p.y)) ? 1 : 0;  // :598
        double x, y;
        rk3<+1, false>(g, dt, p.x, p.y, x, y);
        clampPos(g.nx, g.ny, g.dx, x, y);
        pos[e] = make_double2(x, y);
    }
    c = warpMax(c > 0 ? c : 0.0);
    unsigned int nanMask = __ballot_sync(0xffffffffu, nan);
    if ((threadIdx.x & 31) == 0) {
        if (c > 0) atomicMax(reinterpret_cast<unsigned long long*>(&ctl->cflMax), (unsigned long long)__double_as_longlong(c));
        if (nanMask) atomicAdd(&ctl->nanCount, __popc(nanMask));
    }
}

__global__ void particleEnergyKernel(const double2* __restrict__ pos, const double2* __restrict__ vel, size_t np,
                                     double m, double gx, double gy, double* partials, unsigned int* counter,
                                     DevCtl* ctl) {
    __shared__ double red[32];
    double acc = 0.0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < np; e += (size_t)gridDim.x * blockDim.x) {
        double2 p = pos[e], w = vel[e];
        acc += 0.5 * m * (w.x * w.x + w.y * w.y);
        acc -= m * (gx * p.x + gy * p.y);
    }
    acc = blockReduce<false>(acc, red);
    gridReduceFinish<false>(acc, partials, counter, red, [&](double t) { ctl->particleEnergy = t; });
}

__global__ void copy2Kernel(const double* __restrict__ a, double* __restrict__ b, const double* __restrict__ c,
                            double* __restrict__ d, size_t n) {
    size_t k = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    for (; k < n; k += (size_t)gridDim.x * blockDim.x * 2) {
        *reinterpret_cast<double2*>(b + k) = *reinterpret_cast<const double2*>(a + k);
        *reinterpret_cast<double2*>(d + k) = *reinterpret_cast<const double2*>(c + k);
    }
}

// mac.copyFrom(newMac) restricted to the faces whose BFS layer (distU / distV: 0 = known) is at most *nearPtr, plus the
// everything when the array has no known face at all (nothing is filled then)
__global__ void copy2NearKernel(const double* __restrict__ nu, double* __restrict__ u, const double* __restrict__ nv,
                                double* __restrict__ v, const int* __restrict__ distU, const int* __restrict__ distV, int nx,
                                int ny, int pitch, const int* nearPtr, const int* anyKnown) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    const int K = max(*nearPtr, 1);
    const bool allU = anyKnown[0] == 0, allV = anyKnown[1] == 0;  // (no fill at all: the distances are stale)
    const long long o = (long long)j * pitch + i;
    if (i <= nx && j < ny) { if (allU || distU[o] <= K) u[o] = nu[o]; }
    if (i < nx && j <= ny) { if (allV || distV[o] <= K) v[o] = nv[o]; }
}

__global__ void addConstKernel(double* __restrict__ u, double du, double* __restrict__ v, double dv, int nx, int ny,
                               int pitch) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i <= nx && j < ny) u[(long long)j * pitch + i] += du;
    if (i < nx && j <= ny) v[(long long)j * pitch + i] += dv;
}

}  // namespace

int sortParticlesByCell(Sim* s) {
    if (s->skipSort) return FSIM_OK;  // (runFrame sorted them for both consumers)
    size_t ncells = (size_t)s->nx * s->ny;
    CUDA_TRY(cudaMemsetAsync(s->cellCursor, 0, ncells * sizeof(uint32_t), s->stream));
    if (s->np == 0) {
        CUDA_TRY(cudaMemsetAsync(s->cellStart, 0, (ncells + 1) * sizeof(uint32_t), s->stream));
        return FSIM_OK;
    }
    unsigned pb = (unsigned)((s->np + 255) / 256);
    countKernel<<<pb, 256, 0, s->stream>>>(s->pos, s->np, s->dx, s->nx, s->ny, s->pcell, s->cellCursor);
    LAUNCH_COUNT(s);
    int nb = (int)((ncells + SCAN_CHUNK - 1) / SCAN_CHUNK);
    scanReduceKernel<<<nb, SCAN_THREADS, 0, s->stream>>>(s->cellCursor, ncells, s->scanTmp);
    scanBlockSumsKernel<<<1, 1024, 0, s->stream>>>(s->scanTmp, nb);
    scanFinalKernel<<<nb, SCAN_THREADS, 0, s->stream>>>(s->cellCursor, ncells, s->scanTmp, s->cellStart, (uint32_t)s->np);
    s->launches += 3;
    CUDA_TRY(cudaMemsetAsync(s->cellCursor, 0, ncells * sizeof(uint32_t), s->stream));
    fillKernel<<<pb, 256, 0, s->stream>>>(s->pcell, s->np, s->cellStart, s->cellCursor, s->sortedIdx);
    sortCellsKernel<<<(unsigned)((ncells + 255) / 256), 256, 0, s->stream>>>(s->cellStart, ncells, s->sortedIdx);
    s->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}

int stageTransferVelocityToGrid(Sim* s) {
    // the cell sort of createWaterLevelSet is still valid (positions only change in applyAdvection), but a
    // stage-wise caller may have uploaded new particles; sorting again keeps the stage self-contained.
    int rc = sortParticlesByCell(s);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(&s->ctl->anyKnown[0], 0, 2 * sizeof(int), s->stream));
    dim3 blk(32, 8), grd((s->nx + 1 + 31) / 32, (s->ny + 1 + 7) / 8);
    p2gGatherKernel<0><<<grd, blk, 0, s->stream>>>(s->pos, s->vel, s->cellStart, s->sortedIdx, s->nx, s->ny,
                                                   s->fr.pitch, s->dx, s->u, s->unkU, &s->ctl->anyKnown[0]);
    p2gGatherKernel<1><<<grd, blk, 0, s->stream>>>(s->pos, s->vel, s->cellStart, s->sortedIdx, s->nx, s->ny,
                                                   s->fr.pitch, s->dx, s->v, s->unkV, &s->ctl->anyKnown[1]);
    s->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return extrapolatePair(s, s->u, s->v, s->unkU, s->unkV);
}

int stageApplyGravity(Sim* s) {
    dim3 blk(32, 8), grd((s->nx + 1 + 31) / 32, (s->ny + 1 + 7) / 8);
    addConstKernel<<<grd, blk, 0, s->stream>>>(s->u, s->dt * s->gx, s->v, s->dt * s->gy, s->nx, s->ny, s->fr.pitch);
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}

static int copyMacFromNew(Sim* s) {
    if (s->farPending) {
        // split extrapolation (projection.cu stageUpdateVelocity): the far layers reach mac from the fill itself
        dim3 blk(32, 8), grd((s->nx + 1 + 31) / 32, (s->ny + 1 + 7) / 8);
        copy2NearKernel<<<grd, blk, 0, s->stream>>>(s->nu, s->u, s->nv, s->v, s->distU + s->fr.org, s->distV + s->fr.org,
                                                    s->nx, s->ny, s->fr.pitch, &s->ctl->nearLayers, s->ctl->anyKnown);
        LAUNCH_COUNT(s);
        CUDA_TRY(cudaGetLastError());
        return FSIM_OK;
    }
    // mac.copyFrom(newMac): whole frames (halo is zero in both)
    copy2Kernel<<<1184, 256, 0, s->stream>>>(s->nu - s->fr.org, s->u - s->fr.org, s->nv - s->fr.org, s->v - s->fr.org,
                                             s->fr.elems);
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}

int copyNewMacToMac(Sim* s) { return copyMacFromNew(s); }

int stageUpdateParticleVelocities(Sim* s) {
    if (s->np) {
        unsigned pb = (unsigned)((s->np + 255) / 256);
        g2pKernel<<<pb, 256, 0, s->stream>>>(s->pos, s->vel, s->np, s->u, s->v, s->nu, s->nv, s->nx, s->ny, s->fr.pitch,
                                             s->dx, s->alpha);
        LAUNCH_COUNT(s);
        CUDA_TRY(cudaGetLastError());
    }
    return copyMacFromNew(s);
}

int stageApplyAdvection(Sim* s) {
    CUDA_TRY(cudaMemsetAsync(&s->ctl->nanCount, 0, sizeof(int), s->stream));
    CUDA_TRY(cudaMemsetAsync(&s->ctl->cflMax, 0, sizeof(double), s->stream));
    if (s->np == 0) return FSIM_OK;
    GridView g{s->u, s->v, s->nx, s->ny, s->fr.pitch, s->dx};
    // fsim_step_host with page-locked mirrors: the positions are the last field of the frame to become final.  Advected in
    // chunks, each chunk's download starts on the copy stream as soon as its kernel is done, so only the last chunk's
    // copy remains after the frame (capi.cu mirrorDownload skips what is marked done here).
    const bool chunked = s->mirror && s->mirrorOverlap && s->mirror->particles && !(s->mirrorDone & 32u) && s->np >= (1u << 18);
    const int nch = chunked ? 8 : 1;
    const size_t per = ((s->np + nch - 1) / nch + 127) / 128 * 128;
    for (int q = 0; q < nch; ++q) {
        const size_t e0 = (size_t)q * per;
        if (e0 >= s->np) break;
        const size_t cnt = s->np - e0 < per ? s->np - e0 : per;
        advectKernel<<<(unsigned)((cnt + 127) / 128), 128, 0, s->stream>>>(s->pos + e0, s->vel + e0, cnt, g, s->dt, s->ctl);
        LAUNCH_COUNT(s);
        if (chunked) {
            CUDA_TRY(cudaEventRecord(s->evChunk[q], s->stream));
            CUDA_TRY(cudaStreamWaitEvent(s->copyStream, s->evChunk[q], 0));
            CUDA_TRY(cudaMemcpyAsync(s->mirror->particles + 2 * e0, s->pos + e0, cnt * 16, cudaMemcpyDeviceToHost, s->copyStream));
        }
    }
    if (chunked) s->mirrorDone |= 32u;  // M_POS
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}

int particleEnergy(Sim* s) {
    int pp = s->ppcSqrt * s->ppcSqrt;
    double m = s->rho * s->dx * s->dx / pp;
    particleEnergyKernel<<<296, 256, 0, s->stream>>>(s->pos, s->vel, s->np, m, s->gx, s->gy, s->partials,
                                                     &s->counters[0], s->ctl);
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}
