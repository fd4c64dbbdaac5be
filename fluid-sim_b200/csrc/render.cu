// render.cu -- SURVEY.md section 8(f) rank 3: the staging arrays the reference's renderer rebuilds from the simulation state
// every frame (demo/FluidRenderer2D.cpp:435-486, FluidRenderer2D::updateBuffers), filled from the device-resident state so a
// caller that draws the simulation does not need the per-frame download of the whole state.
//
//   waterCellLocations / solidCellLocations   (:436-448)  vec2f {(float)(i*dx), (float)(j*dx)} of every FLUID / SOLID cell,
//                                                          raster order (FluidSim2D::iterate: j outer, i inner)
//   cellVels                                   (:449-459)  per cell of the (sizeX-1) x (sizeY-1) block: the centre and
//                                                          centre + dt * mac.velInterp(centre), two vec2f
//   pressureCellLocations / pressureCellValues (:460-470)  cells with p != 0: location and sigmoid(0.01f * (float)p)
//   particleVelLines                           (:471-479)  per particle: position and position + dt * velInterp(position),
//                                                          in FLOAT arithmetic as the reference writes it
//   phiCellValues                              (:480-485)  sigmoid<float>(100.0f * phi) of every cell
//
// The variable-length lists keep the reference's raster order: a stable compaction (per-block counts, one scan, scatter).
#include "sampling.cuh"
#include "sim.h"

namespace {

constexpr int RB = 256;  // cells per compaction block (consecutive in raster order)

__device__ __forceinline__ float sigmoidf32(float x) {  // aml::sigmoid<float> (deps/altmath/src/math_utils.h:76-79)
    return (float)(1.0 / (1.0 + exp(-(double)x)));
}

// pass 1: per block of RB raster-order cells, how many are FLUID / SOLID / have p != 0
__global__ void __launch_bounds__(RB) renderCountKernel(const uint8_t* __restrict__ cell, const double* __restrict__ p, int nx, int ny,
                                                        int pitch, unsigned int* __restrict__ counts /* [3][nblocks] */, int nblocks) {
    const long long q = (long long)blockIdx.x * RB + threadIdx.x;
    bool w = false, so = false, pr = false;
    if (q < (long long)nx * ny) {
        const int j = (int)(q / nx), i = (int)(q - (long long)j * nx);
        const long long o = (long long)j * pitch + i;
        const uint8_t c = cell[o];
        w = c == FSIM_CELL_FLUID; so = c == FSIM_CELL_SOLID; pr = p[o] != 0.0;
    }
    const int cw = __syncthreads_count(w), cs = __syncthreads_count(so), cp = __syncthreads_count(pr);
    if (threadIdx.x == 0) { counts[blockIdx.x] = cw; counts[nblocks + blockIdx.x] = cs; counts[2 * nblocks + blockIdx.x] = cp; }
}

// pass 2: exclusive scan of the three count rows (one block per row; the totals go to totals[3])
__global__ void __launch_bounds__(1024) renderScanKernel(unsigned int* counts, int nblocks, unsigned int* totals) {
    __shared__ unsigned int part[1024];
    unsigned int* row = counts + (size_t)blockIdx.x * nblocks;
    const int per = (nblocks + 1023) / 1024, b0 = threadIdx.x * per, b1 = min(nblocks, b0 + per);
    unsigned int sum = 0;
    for (int k = b0; k < b1; ++k) sum += row[k];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // inclusive scan of the per-thread sums
        unsigned int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned int run = threadIdx.x ? part[threadIdx.x - 1] : 0;
    for (int k = b0; k < b1; ++k) { const unsigned int c = row[k]; row[k] = run; run += c; }
    if (threadIdx.x == 1023) totals[blockIdx.x] = part[1023];
}

// pass 3: stable scatter in raster order
__global__ void __launch_bounds__(RB) renderScatterKernel(const uint8_t* __restrict__ cell, const double* __restrict__ p, int nx, int ny,
                                                          int pitch, double dx, const unsigned int* __restrict__ offs, int nblocks,
                                                          float2* water, float2* solid, float2* pLoc, float* pVal,
                                                          unsigned int capW, unsigned int capS, unsigned int capP) {
    __shared__ unsigned int wsum[3][RB / 32];
    const long long q = (long long)blockIdx.x * RB + threadIdx.x;
    bool f[3] = {false, false, false};
    float2 loc = make_float2(0.f, 0.f);
    float pv = 0.f;
    if (q < (long long)nx * ny) {
        const int j = (int)(q / nx), i = (int)(q - (long long)j * nx);
        const long long o = (long long)j * pitch + i;
        const uint8_t c = cell[o];
        const double pd = p[o];
        f[0] = c == FSIM_CELL_FLUID; f[1] = c == FSIM_CELL_SOLID; f[2] = pd != 0.0;
        loc = make_float2((float)((double)i * dx), (float)((double)j * dx));
        pv = sigmoidf32(0.01f * (float)pd);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int pre[3];
    for (int k = 0; k < 3; ++k) {
        const unsigned int m = __ballot_sync(0xffffffffu, f[k]);
        pre[k] = __popc(m & ((1u << lane) - 1));
        if (lane == 0) wsum[k][w] = __popc(m);
    }
    __syncthreads();
    for (int k = 0; k < 3; ++k) {
        unsigned int base = offs[(size_t)k * nblocks + blockIdx.x];
        for (int ww = 0; ww < w; ++ww) base += wsum[k][ww];
        const unsigned int at = base + pre[k];
        if (!f[k]) continue;
        if (k == 0 && water && at < capW) water[at] = loc;
        if (k == 1 && solid && at < capS) solid[at] = loc;
        if (k == 2 && at < capP) { if (pLoc) pLoc[at] = loc; if (pVal) pVal[at] = pv; }
    }
}

__global__ void renderCellVelsKernel(GridView g, double dt, float2* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.nx - 1 || j >= g.ny - 1) return;
    const double x = ((double)i + 0.5) * g.dx, y = ((double)j + 0.5) * g.dx;
    const double vx = sampleU<false>(g, x, y), vy = sampleV<false>(g, x, y);
    const size_t k = 2 * ((size_t)j * (g.nx - 1) + i);
    out[k] = make_float2((float)x, (float)y);
    out[k + 1] = make_float2((float)(x + vx * dt), (float)(y + vy * dt));
}

__global__ void renderParticleLinesKernel(const double2* __restrict__ pos, size_t np, GridView g, double dt, float2* __restrict__ out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= np) return;
    const double2 p = pos[e];
    const double vx = sampleU<false>(g, p.x, p.y), vy = sampleV<false>(g, p.x, p.y);
    const float fx = (float)p.x, fy = (float)p.y, fdt = (float)dt;
    out[2 * e] = make_float2(fx, fy);
    // (float)x + (float)dt * (float)v, contracted like the reference's -O2 -mfma build does
    out[2 * e + 1] = make_float2(__fmaf_rn(fdt, (float)vx, fx), __fmaf_rn(fdt, (float)vy, fy));
}

__global__ void renderPhiKernel(const double* __restrict__ phi, int nx, int ny, int pitch, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nx || j >= ny) return;
    out[(size_t)j * nx + i] = sigmoidf32((float)(100.0 * phi[(long long)j * pitch + i]));
}

}  // namespace

extern "C" int fsim_render_fill(fsim_handle h, fsim_render_staging* io) {
    if (!h || !io) { fsim_set_error("null argument"); return FSIM_E_INVALID; }
    Sim* s = reinterpret_cast<Sim*>(h);
    CUDA_TRY(cudaSetDevice(s->device));
    const int nx = s->nx, ny = s->ny;
    const size_t ncells = (size_t)nx * ny;
    const int nblocks = (int)((ncells + RB - 1) / RB);
    // device staging (allocated on first use): lists are sized for the worst case
    const size_t nvels = (size_t)(nx - 1) * (ny - 1);
    const size_t need = ncells * 8 * 3 + ncells * 4 * 2 + nvels * 16 + s->np * 16 + (size_t)3 * nblocks * 4 + 64;
    if (s->renderBytes < need) {
        if (s->renderBuf) {
            for (auto& r : s->rawAllocs) if (r == s->renderBuf) r = nullptr;
            cudaFree(s->renderBuf);
        }
        CUDA_TRY(cudaMalloc(&s->renderBuf, need));
        s->rawAllocs.push_back(s->renderBuf);
        s->renderBytes = need;
    }
    char* q = static_cast<char*>(s->renderBuf);
    float2* dWater = reinterpret_cast<float2*>(q); q += ncells * 8;
    float2* dSolid = reinterpret_cast<float2*>(q); q += ncells * 8;
    float2* dPLoc = reinterpret_cast<float2*>(q); q += ncells * 8;
    float* dPVal = reinterpret_cast<float*>(q); q += ncells * 4;
    float* dPhi = reinterpret_cast<float*>(q); q += ncells * 4;
    float2* dVels = reinterpret_cast<float2*>(q); q += nvels * 16;
    float2* dLines = reinterpret_cast<float2*>(q); q += s->np * 16;
    unsigned int* dCounts = reinterpret_cast<unsigned int*>(q); q += (size_t)3 * nblocks * 4;
    unsigned int* dTotals = reinterpret_cast<unsigned int*>(q);
    GridView g{s->u, s->v, nx, ny, s->fr.pitch, s->dx};
    cudaStream_t st = s->stream;
    const bool lists = io->waterCells || io->solidCells || io->pressureCells || io->pressureValues;
    unsigned int totals[3] = {0, 0, 0};
    if (lists) {
        renderCountKernel<<<nblocks, RB, 0, st>>>(s->cell, s->p, nx, ny, s->fr.pitch, dCounts, nblocks);
        renderScanKernel<<<3, 1024, 0, st>>>(dCounts, nblocks, dTotals);
        renderScatterKernel<<<nblocks, RB, 0, st>>>(s->cell, s->p, nx, ny, s->fr.pitch, s->dx, dCounts, nblocks, dWater, dSolid, dPLoc, dPVal,
                                                     (unsigned)ncells, (unsigned)ncells, (unsigned)ncells);
        s->launches += 3;
        CUDA_TRY(cudaMemcpyAsync(totals, dTotals, sizeof(totals), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        io->nWater = totals[0]; io->nSolid = totals[1]; io->nPressure = totals[2];
        if (totals[0] > io->waterCap && io->waterCells) { fsim_set_error("waterCells holds %zu entries, %u needed", io->waterCap, totals[0]); return FSIM_E_INVALID; }
        if (totals[1] > io->solidCap && io->solidCells) { fsim_set_error("solidCells holds %zu entries, %u needed", io->solidCap, totals[1]); return FSIM_E_INVALID; }
        if (totals[2] > io->pressureCap && (io->pressureCells || io->pressureValues)) { fsim_set_error("pressure lists hold %zu entries, %u needed", io->pressureCap, totals[2]); return FSIM_E_INVALID; }
        if (io->waterCells && totals[0]) CUDA_TRY(cudaMemcpyAsync(io->waterCells, dWater, (size_t)totals[0] * 8, cudaMemcpyDeviceToHost, st));
        if (io->solidCells && totals[1]) CUDA_TRY(cudaMemcpyAsync(io->solidCells, dSolid, (size_t)totals[1] * 8, cudaMemcpyDeviceToHost, st));
        if (io->pressureCells && totals[2]) CUDA_TRY(cudaMemcpyAsync(io->pressureCells, dPLoc, (size_t)totals[2] * 8, cudaMemcpyDeviceToHost, st));
        if (io->pressureValues && totals[2]) CUDA_TRY(cudaMemcpyAsync(io->pressureValues, dPVal, (size_t)totals[2] * 4, cudaMemcpyDeviceToHost, st));
    }
    if (io->cellVels && nvels) {
        dim3 blk(32, 8), grd((nx - 1 + 31) / 32, (ny - 1 + 7) / 8);
        renderCellVelsKernel<<<grd, blk, 0, st>>>(g, s->dt, dVels);
        LAUNCH_COUNT(s);
        CUDA_TRY(cudaMemcpyAsync(io->cellVels, dVels, nvels * 16, cudaMemcpyDeviceToHost, st));
    }
    if (io->particleVelLines && s->np) {
        renderParticleLinesKernel<<<(unsigned)((s->np + 255) / 256), 256, 0, st>>>(s->pos, s->np, g, s->dt, dLines);
        LAUNCH_COUNT(s);
        CUDA_TRY(cudaMemcpyAsync(io->particleVelLines, dLines, s->np * 16, cudaMemcpyDeviceToHost, st));
    }
    if (io->phiValues) {
        dim3 blk(32, 8), grd((nx + 31) / 32, (ny + 7) / 8);
        renderPhiKernel<<<grd, blk, 0, st>>>(s->phi, nx, ny, s->fr.pitch, dPhi);
        LAUNCH_COUNT(s);
        CUDA_TRY(cudaMemcpyAsync(io->phiValues, dPhi, ncells * 4, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(st));
    return FSIM_OK;
}
