// Shared device/host helpers for the sm_100a fluid-sim hot path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define FSIM_CELL_EMPTY 0
#define FSIM_CELL_FLUID 1
#define FSIM_CELL_SOLID 2

// ---------------------------------------------------------------------------------------------
// Frame: the one HBM layout shared by every grid-shaped array (cell-centred, u-faces, v-faces,
// labels, masks).  Logical element (i,j) lives at base[j*pitch + i]; `base` already points at (0,0).
// There is a halo of FR_HALO elements on every side (zero-filled, never written with non-zeros), the
// width/height are rounded up to 32 so the wavefront kernels can work on whole 32x32 blocks without
// tails, and (0,0) is 128-byte aligned with pitch a multiple of 16 doubles so 16-byte cp.async and
// vector loads are always aligned.
// ---------------------------------------------------------------------------------------------
#define FR_HALO 32

struct Frame {
    int nx, ny;      // cell counts (sizeX, sizeY)
    int W, H;        // padded logical extent: roundup(nx+1,32), roundup(ny+1,32)
    int pitch;       // elements per row
    int rows;        // total rows incl. halo
    size_t elems;    // pitch*rows
    size_t org;      // offset of (0,0) in elements
};

static inline Frame makeFrame(int nx, int ny) {
    Frame f;
    f.nx = nx; f.ny = ny;
    f.W = ((nx + 1 + 31) / 32) * 32;
    f.H = ((ny + 1 + 31) / 32) * 32;
    f.pitch = f.W + 2 * FR_HALO;
    f.rows = f.H + 2 * FR_HALO;
    f.elems = (size_t)f.pitch * f.rows;
    f.org = (size_t)FR_HALO * f.pitch + FR_HALO;
    return f;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            fsim_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return FSIM_E_CUDA;                                                             \
        }                                                                                   \
    } while (0)

void fsim_set_error(const char* fmt, ...);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double warpSum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warpMax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide reduction (blockDim.x multiple of 32, <= 1024). Result valid in thread 0.
template <bool IS_MAX>
__device__ __forceinline__ double blockReduce(double v, double* smem /* >= 32 doubles */) {
    v = IS_MAX ? warpMax(v) : warpSum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? smem[threadIdx.x] : 0.0;
    if (w == 0) v = IS_MAX ? warpMax(v) : warpSum(v);
    return v;
}

// "last block done" pattern: every block stores a partial, the last one to arrive reduces all partials in
// a fixed order (deterministic) and runs `fin(total)`.  counter must be zero on entry and is reset.
template <bool IS_MAX, class Fin>
__device__ __forceinline__ void gridReduceFinish(double blockValue, double* partials, unsigned int* counter,
                                                 double* smem, Fin fin) {
    __shared__ bool isLast;
    unsigned int nblocks = gridDim.x * gridDim.y;
    unsigned int bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        partials[bid] = blockValue;
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        isLast = (t == nblocks - 1);
    }
    __syncthreads();
    if (isLast) {
        double acc = 0.0;
        for (unsigned int k = threadIdx.x; k < nblocks; k += blockDim.x) {
            double pv = __ldcg(&partials[k]);
            acc = IS_MAX ? fmax(acc, pv) : acc + pv;
        }
        acc = blockReduce<IS_MAX>(acc, smem);
        if (threadIdx.x == 0) {
            fin(acc);
            *counter = 0;
        }
    }
}

__device__ __forceinline__ int iclampd(int v, int lo, int hi) { return max(lo, min(v, hi)); }
// aml::clamp on doubles: max(lo, min(v, hi)) with (a<b)?a:b / (a>b)?a:b (deps/altmath/src/math_utils.h:16-48)
__device__ __forceinline__ double amlMin(double a, double b) { return (a < b) ? a : b; }
__device__ __forceinline__ double amlMax(double a, double b) { return (a > b) ? a : b; }
__device__ __forceinline__ double amlClamp(double v, double lo, double hi) { return amlMax(lo, amlMin(v, hi)); }

#endif  // __CUDACC__
