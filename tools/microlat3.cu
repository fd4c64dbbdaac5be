// single-warp issue rates of MIO instructions on B200: LDS.64 / LDS.128 / STS.64 / STS.128 / SHFL, alone and with
// a second warp running the same loop on another SMSP.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk(){ long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
constexpr int IT = 512;
template <int V>
__global__ void k(double* out, long long* cyc) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 1e-3;
    __syncthreads();
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const double* base = sm + warp * 2048;
    long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < IT; ++it) {
        if (V == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += base[((it + j) & 31) * 32 + lane];
        } else if (V == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { double2 v = reinterpret_cast<const double2*>(base)[((it + j) & 15) * 32 + lane]; acc[j] += v.x + v.y; }
        } else if (V == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) const_cast<double*>(base)[((it + j) & 31) * 32 + lane] = acc[j] + it;
        } else if (V == 3) {
#pragma unroll
            for (int j = 0; j < 8; ++j) reinterpret_cast<double2*>(const_cast<double*>(base))[((it + j) & 15) * 32 + lane] = make_double2(acc[j] + it, 1.0);
        } else if (V == 4) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { float f = __shfl_up_sync(0xffffffffu, (float)acc[j] + it, 1); acc[j] = f; }
        } else if (V == 5) {  // LDS.64 with 4 loads per address computation, immediate offsets
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += base[j * 32 + lane];
        }
    }
    long long t1 = clk();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0; for (int j = 0; j < 8; ++j) s += acc[j];
    out[threadIdx.x] = s;
}
template <int V> void run(const char* name, double* out, long long* cyc) {
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int nw = 1; nw <= 4; nw *= 2) {
        for (int r = 0; r < 2; ++r) { k<V><<<1, 32 * nw, 8192 * 8>>>(out, cyc); cudaDeviceSynchronize(); }
        printf("%-28s warps %d: %.2f cyc/instr (warp 0)\n", name, nw, cyc[0] / (double)(IT * 8));
    }
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 128 * 8); cudaMallocManaged(&cyc, 8);
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    run<0>("LDS.64", out, cyc); run<1>("LDS.128", out, cyc); run<2>("STS.64", out, cyc); run<3>("STS.128", out, cyc); run<4>("SHFL.32", out, cyc); run<5>("LDS.64 imm", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
