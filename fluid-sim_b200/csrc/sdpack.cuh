// sdpack.cuh -- row-major frame <-> strip-diagonal (SD) layout, optionally x-mirrored (layout column c holds grid
// column nx-1-c: the layout the x-descending sweeps run on).
#pragma once

#include "sdwave.cuh"

namespace sd {

struct PackJob { const double* src[10]; double* dst[10]; };

#ifdef __CUDACC__
// row-major frame -> SD (zero outside nx x ny); grid (nchunks, nstrips, arrays), block (32, 8)
static __global__ void __launch_bounds__(256) sdPackKernel(PackJob job, Geom g, int pitch, int mirror) {
    __shared__ double tile[32][33];
    const double* __restrict__ src = job.src[blockIdx.z];
    double* __restrict__ dst = job.dst[blockIdx.z];
    const int k = blockIdx.y, s0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        int c = s0 + (int)threadIdx.x - g.sigma * r, j = 32 * k + r;
        int i = mirror ? g.nx - 1 - c : c;
        tile[r][threadIdx.x] = (c >= 0 && c < g.nx && j < g.ny) ? src[(long long)j * pitch + i] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8)
        dst[((size_t)k * g.Sp + s0 + r) * 32 + threadIdx.x] = tile[threadIdx.x][r];
}

// SD -> row-major frame (logical nx x ny only); job.src = SD arrays, job.dst = frames
static __global__ void __launch_bounds__(256) sdUnpackKernel(PackJob job, Geom g, int pitch, int mirror) {
    __shared__ double tile[32][33];
    const double* __restrict__ src = job.src[blockIdx.z];
    double* __restrict__ dst = job.dst[blockIdx.z];
    const int k = blockIdx.y, s0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += 8)
        tile[threadIdx.x][r] = src[((size_t)k * g.Sp + s0 + r) * 32 + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        int c = s0 + (int)threadIdx.x - g.sigma * r, j = 32 * k + r;
        int i = mirror ? g.nx - 1 - c : c;
        if (c >= 0 && c < g.nx && j < g.ny) dst[(long long)j * pitch + i] = tile[r][threadIdx.x];
    }
}
#endif

}  // namespace sd
