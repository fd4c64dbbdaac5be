// sdpack.cuh -- row-major frame <-> strip-diagonal (SD) layout (sdwave.cuh: Geom, rows per lane included), optionally
// x-mirrored (layout column c holds grid column nx-1-c: the layout the x-descending sweeps run on).
#pragma once

#include "sdwave.cuh"

namespace sd {

struct PackJob { const double* src[10]; double* dst[10]; };

#ifdef __CUDACC__
// row-major frame -> SD (zero outside nx x ny and below local row rowLo); grid (nchunks, nstrips, arrays * rpl), block (32, 8)
// gate (optional): no-op when *gate == 0; mirrorOut (optional): records `mirror` on the device when the kernel does run
static __global__ void __launch_bounds__(256) sdPackKernel(PackJob job, Geom g, int pitch, int mirror, int rowLo = 0,
                                                           const int* gate = nullptr, int* mirrorOut = nullptr) {
    __shared__ double tile[32][33];
    if (gate && *gate == 0) return;
    if (mirrorOut && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && threadIdx.y == 0) *mirrorOut = mirror;
    const int R = g.rpl, r = blockIdx.z % R;
    const double* __restrict__ src = job.src[blockIdx.z / R];
    double* __restrict__ dst = job.dst[blockIdx.z / R];
    const int k = blockIdx.y, s0 = blockIdx.x * 32;
    for (int t = threadIdx.y; t < 32; t += 8) {  // lane t owns rows R*t .. R*t+R-1 of the strip
        int c = s0 + (int)threadIdx.x - g.sigma * t, j = 32 * R * k + R * t + r;
        int i = mirror ? g.nx - 1 - c : c;
        tile[t][threadIdx.x] = (c >= 0 && c < g.nx && j < g.ny && j >= rowLo) ? src[(long long)j * pitch + i] : 0.0;
    }
    __syncthreads();
    for (int st = threadIdx.y; st < 32; st += 8)
        dst[(((size_t)k * g.Sp + s0 + st) * 32 + threadIdx.x) * R + r] = tile[threadIdx.x][st];
}

// SD -> row-major frame (logical nx x ny only); job.src = SD arrays, job.dst = frames
// gate (optional): no-op when *gate == 0; mirrorIn (optional): the layout's mirror flag as the device recorded it
static __global__ void __launch_bounds__(256) sdUnpackKernel(PackJob job, Geom g, int pitch, int mirror, const int* gate = nullptr,
                                                             const int* mirrorIn = nullptr) {
    __shared__ double tile[32][33];
    if (gate && *gate == 0) return;
    if (mirrorIn) mirror = *mirrorIn;
    const int R = g.rpl, r = blockIdx.z % R;
    const double* __restrict__ src = job.src[blockIdx.z / R];
    double* __restrict__ dst = job.dst[blockIdx.z / R];
    const int k = blockIdx.y, s0 = blockIdx.x * 32;
    for (int st = threadIdx.y; st < 32; st += 8)
        tile[threadIdx.x][st] = src[(((size_t)k * g.Sp + s0 + st) * 32 + threadIdx.x) * R + r];
    __syncthreads();
    for (int t = threadIdx.y; t < 32; t += 8) {
        int c = s0 + (int)threadIdx.x - g.sigma * t, j = 32 * R * k + R * t + r;
        int i = mirror ? g.nx - 1 - c : c;
        if (c >= 0 && c < g.nx && j < g.ny) dst[(long long)j * pitch + i] = tile[t][threadIdx.x];
    }
}
#endif

}  // namespace sd
