// micro-latency probe for the wavefront design (B200): dependent DFMA chain, SHFL, LDS, DADD, DMUL.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk(){ long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
__global__ void k(double* out, long long* cyc, double a, double b) {
    __shared__ double sm[64];
    sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
    __syncwarp();
    double x = a + threadIdx.x;
    long long t0, t1;
    // DFMA chain
    asm volatile("" : "+d"(x)); t0 = clk(); asm volatile("" : "+d"(x));
#pragma unroll
    for (int i = 0; i < 256; ++i) x = __fma_rn(x, b, a);
    asm volatile("" : "+d"(x)); t1 = clk(); asm volatile("" : "+d"(x));
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DADD chain
    asm volatile("" : "+d"(x)); t0 = clk(); asm volatile("" : "+d"(x));
#pragma unroll
    for (int i = 0; i < 256; ++i) x = __dadd_rn(x, b);
    asm volatile("" : "+d"(x)); t1 = clk(); asm volatile("" : "+d"(x));
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // DMUL chain
    asm volatile("" : "+d"(x)); t0 = clk(); asm volatile("" : "+d"(x));
#pragma unroll
    for (int i = 0; i < 256; ++i) x = __dmul_rn(x, b);
    asm volatile("" : "+d"(x)); t1 = clk(); asm volatile("" : "+d"(x));
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // SHFL (double = 2x32) chain
    asm volatile("" : "+d"(x)); t0 = clk(); asm volatile("" : "+d"(x));
#pragma unroll
    for (int i = 0; i < 256; ++i) x = __shfl_up_sync(0xffffffffu, x, 1);
    asm volatile("" : "+d"(x)); t1 = clk(); asm volatile("" : "+d"(x));
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // SHFL + DFMA chain
    asm volatile("" : "+d"(x)); t0 = clk(); asm volatile("" : "+d"(x));
#pragma unroll
    for (int i = 0; i < 256; ++i) x = __fma_rn(__shfl_up_sync(0xffffffffu, x, 1), b, a);
    asm volatile("" : "+d"(x)); t1 = clk(); asm volatile("" : "+d"(x));
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // LDS pointer chase
    int idx = threadIdx.x;
    reinterpret_cast<int*>(sm)[threadIdx.x] = (threadIdx.x + 1) & 31;
    __syncwarp();
    asm volatile("" : "+d"(x)); t0 = clk(); asm volatile("" : "+d"(x));
#pragma unroll
    for (int i = 0; i < 256; ++i) idx = reinterpret_cast<volatile int*>(sm)[idx];
    asm volatile("" : "+d"(x)); t1 = clk(); asm volatile("" : "+d"(x));
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // independent DFMA throughput (8 chains)
    double y[8];
    for (int i = 0; i < 8; ++i) y[i] = x + i;
    asm volatile("" : "+d"(x)); t0 = clk(); asm volatile("" : "+d"(x));
#pragma unroll
    for (int i = 0; i < 256; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = __fma_rn(y[j], b, a);
    for (int i = 0; i < 8; ++i) x += y[i];
    asm volatile("" : "+d"(x)); t1 = clk(); asm volatile("" : "+d"(x));
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    for (int i = 0; i < 8; ++i) x += y[i];
    // FFMA chain for reference
    float f = (float)x; asm volatile("" : "+f"(f));
    asm volatile("" : "+d"(x)); t0 = clk(); asm volatile("" : "+d"(x));
#pragma unroll
    for (int i = 0; i < 256; ++i) f = __fmaf_rn(f, (float)b, (float)a);
    x += f;
    asm volatile("" : "+d"(x)); t1 = clk(); asm volatile("" : "+d"(x));
    if (threadIdx.x == 0) cyc[7] = t1 - t0;
    out[threadIdx.x] = x + idx + f;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 8 * 8);
    for (int r = 0; r < 2; ++r) { k<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); cudaDeviceSynchronize(); }
    const char* names[] = {"DFMA chain", "DADD chain", "DMUL chain", "SHFL.f64 chain", "SHFL+DFMA chain", "LDS chase", "DFMA x8 indep (per 8)", "FFMA chain"};
    for (int i = 0; i < 8; ++i) printf("%-24s %.2f cyc/op\n", names[i], cyc[i] / 256.0);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
