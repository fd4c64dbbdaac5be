// micro-probe of the solver warp's inner loop (one warp, one SM): which ingredient costs what.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk(){ long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
constexpr int STEPS = 2048;
template <int V, int SIGMA>
__global__ void k(double* out, long long* cyc, double a, double b) {
    extern __shared__ double sm[];  // [3][32 steps][32 lanes] ring
    const int lane = threadIdx.x;
    for (int i = lane; i < 3 * 32 * 32; i += 32) sm[i] = (i % 7) * 1e-3;
    __syncwarp();
    double y = a, shq[SIGMA];
    for (int i = 0; i < SIGMA; ++i) shq[i] = 0;
    long long t0 = clk();
    if (V <= 3) {
#pragma unroll 8
        for (int s = 0; s < STEPS; ++s) {
            const int ls = s & 31;
            double r = a, cx = b, cy = b;
            if (V >= 3) { r = sm[ls * 32 + lane]; cx = sm[1024 + ls * 32 + lane]; cy = sm[2048 + ls * 32 + lane]; }
            y = __fma_rn(-cx, y, __fma_rn(-cy, shq[0], r));
#pragma unroll
            for (int i = 0; i + 1 < SIGMA; ++i) shq[i] = shq[i + 1];
            shq[SIGMA - 1] = __shfl_up_sync(0xffffffffu, y, 1);
            if (V >= 2) sm[ls * 32 + lane] = y;
        }
    } else {
        double v[2][8][3];
        for (int e = 0; e < 8; ++e) for (int c = 0; c < 3; ++c) v[0][e][c] = sm[c * 1024 + e * 32 + lane];
#pragma unroll 1
        for (int s0 = 0; s0 < STEPS; s0 += 16) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int base = (s0 + 8 * h + 8) & 31;
#pragma unroll
                for (int e = 0; e < 8; ++e)
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[h ^ 1][e][c] = sm[c * 1024 + (base + e) * 32 + lane];
                const int cur = (s0 + 8 * h) & 31;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    y = __fma_rn(-v[h][e][1], y, __fma_rn(-v[h][e][2], shq[0], v[h][e][0]));
#pragma unroll
                    for (int i = 0; i + 1 < SIGMA; ++i) shq[i] = shq[i + 1];
                    shq[SIGMA - 1] = __shfl_up_sync(0xffffffffu, y, 1);
                    if (V >= 5) sm[(cur + e) * 32 + lane] = y;
                }
            }
        }
    }
    long long t1 = clk();
    if (lane == 0) cyc[0] = t1 - t0;
    out[lane] = y + shq[0];
}
template <int V, int SIGMA> void run(const char* name, double* out, long long* cyc) {
    for (int r = 0; r < 2; ++r) { k<V, SIGMA><<<1, 32, 3 * 32 * 32 * 8>>>(out, cyc, 0.3, 0.2); cudaDeviceSynchronize(); }
    printf("%-44s sigma %d: %.1f cyc/step\n", name, SIGMA, cyc[0] / (double)STEPS);
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 8);
    run<1, 1>("chain + shuffle (regs only)", out, cyc);
    run<1, 2>("chain + shuffle (regs only)", out, cyc);
    run<1, 3>("chain + shuffle (regs only)", out, cyc);
    run<2, 2>("+ STS y", out, cyc);
    run<3, 2>("+ STS y + 3 LDS direct", out, cyc);
    run<3, 3>("+ STS y + 3 LDS direct", out, cyc);
    run<4, 2>("reg-prefetched LDS (8 steps), no STS", out, cyc);
    run<5, 2>("reg-prefetched LDS (8 steps) + STS", out, cyc);
    run<5, 3>("reg-prefetched LDS (8 steps) + STS", out, cyc);
    run<5, 4>("reg-prefetched LDS (8 steps) + STS", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
