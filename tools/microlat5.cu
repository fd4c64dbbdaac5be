// Step-time probe for the two-rows-per-lane MIC(0) solve (tools only): one warp, registers only, the dependency
// structure of one solver step in three formulations.
//   A  today's chain:       shuffle -> select -> DFMA (row F) -> DFMA (row L) -> shuffle ...
//   B  flattened:           row L written as C*down + j with j, kp, B2 prepared before the shuffle arrives;
//                           select on the chain
//   C  flattened, the lane-LC case as a predicated pair of DFMA instead of a select
// Prints cycles per step; the arithmetic of B / C is the algebra used by sd::solveKernelR's flat solver loop.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }
__device__ __forceinline__ double shup(double v) {
    double r;
    asm volatile("{ .reg .b32 lo, hi; mov.b64 {lo,hi}, %1; shfl.sync.up.b32 lo, lo, 1, 0, 0xffffffff; "
                 "shfl.sync.up.b32 hi, hi, 1, 0, 0xffffffff; mov.b64 %0, {lo,hi}; }" : "=d"(r) : "d"(v) : "memory");
    return r;
}
__device__ __forceinline__ double vfma(double a, double b, double c) {
    double r;
    asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c));
    return r;
}
__device__ __forceinline__ double fmaPair(double c, double vLC, double vSh, double add, unsigned isLC) {
    double r;
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %5, 0; @p fma.rn.f64 %0, %1, %2, %4; @!p fma.rn.f64 %0, %1, %3, %4; }"
                 : "=d"(r) : "d"(c), "d"(vLC), "d"(vSh), "d"(add), "r"(isLC));
    return r;
}
constexpr int N = 4096;
__global__ void k(double* out, long long* cyc, const double* coef) {
    const int lane = threadIdx.x;
    const bool isLC = lane == 0;
    // per-lane "coefficients": kept in registers, opaque to the compiler
    double a0 = coef[lane], a1 = coef[32 + lane], x0 = coef[64 + lane], x1 = coef[96 + lane], c0 = coef[128 + lane], c1 = coef[160 + lane];
    double rv = coef[192 + lane];
    long long t0, t1;
    // ---- A
    {
        double y0 = 0, y1 = 0, sh = 0;
        t0 = clk();
#pragma unroll 16
        for (int s = 0; s < N; ++s) {
            double in0 = vfma(-x0, y0, a0), in1 = vfma(-x1, y1, a1);
            double down = isLC ? rv : sh;
            y0 = vfma(-c0, down, in0);
            y1 = vfma(-c1, y0, in1);
            sh = shup(y1);
        }
        t1 = clk();
        if (lane == 0) cyc[0] = t1 - t0;
        out[lane] = y0 + y1;
    }
    // ---- B
    {
        double yF = 0, yL = 0, sh = 0, kp = a1;
        const double X = c1 * x0, C = c1 * c0, kk = vfma(-c1, a0, a1), K2 = vfma(X, a0, kk), P = X * x0, Q = X * c0;
        t0 = clk();
#pragma unroll 16
        for (int s = 0; s < N; ++s) {
            double j = vfma(-x1, yL, kp);
            double inF = vfma(-x0, yF, a0);
            double B2 = vfma(-P, yF, K2);
            double down = isLC ? rv : sh;
            yL = vfma(C, down, j);
            sh = shup(yL);
            yF = vfma(-c0, down, inF);
            kp = vfma(-Q, down, B2);
        }
        t1 = clk();
        if (lane == 0) cyc[1] = t1 - t0;
        out[32 + lane] = yF + yL;
    }
    // ---- C
    {
        double yF = 0, yL = 0, sh = 0, kp = a1;
        const double X = c1 * x0, C = c1 * c0, kk = vfma(-c1, a0, a1), K2 = vfma(X, a0, kk), P = X * x0, Q = X * c0;
        const unsigned lc = isLC;
        t0 = clk();
#pragma unroll 16
        for (int s = 0; s < N; ++s) {
            double j = vfma(-x1, yL, kp);
            double inF = vfma(-x0, yF, a0);
            double B2 = vfma(-P, yF, K2);
            yL = fmaPair(C, rv, sh, j, lc);
            double shn = shup(yL);
            yF = fmaPair(-c0, rv, sh, inF, lc);
            kp = fmaPair(-Q, rv, sh, B2, lc);
            sh = shn;
        }
        t1 = clk();
        if (lane == 0) cyc[2] = t1 - t0;
        out[64 + lane] = yF + yL;
    }
    // ---- D: B plus the per-step coefficient work (6 FP64 ops) and a 16-byte LDS/STS pair, as in the real loop
    {
        __shared__ double sm[64 * 4];
        sm[lane] = a0; sm[lane + 32] = a1; sm[lane + 64] = x0; sm[lane + 96] = x1; sm[lane + 128] = c0; sm[lane + 160] = c1;
        __syncwarp();
        double yF = 0, yL = 0, sh = 0, kp = a1;
        double X = c1 * x0, C = c1 * c0, kk = vfma(-c1, a0, a1), K2 = vfma(X, a0, kk), P = X * x0, Q = X * c0;
        const unsigned sa = (unsigned)__cvta_generic_to_shared(sm) + lane * 16;
        t0 = clk();
#pragma unroll 16
        for (int s = 0; s < N; ++s) {
            double j = vfma(-x1, yL, kp);
            double inF = vfma(-x0, yF, a0);
            double B2 = vfma(-P, yF, K2);
            double l0, l1, l2, l3, l4, l5;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(l0), "=d"(l1) : "r"(sa) : "memory");
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(l2), "=d"(l3) : "r"(sa + 512) : "memory");
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(l4), "=d"(l5) : "r"(sa + 1024) : "memory");
            double down = isLC ? rv : sh;
            yL = vfma(C, down, j);
            sh = shup(yL);
            yF = vfma(-c0, down, inF);
            kp = vfma(-Q, down, B2);
            // next step's derived coefficients from the loaded values (off the chain)
            X = l5 * l2; C = l5 * l4; kk = vfma(-l5, l0, l1); K2 = vfma(X, a0, kk); P = X * x0; Q = X * c0;
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(sa + 1536), "d"(yF), "d"(yL) : "memory");
        }
        t1 = clk();
        if (lane == 0) cyc[3] = t1 - t0;
        out[96 + lane] = yF + yL + X;
    }
    // ---- raw latencies
    {
        double x = a0;
        t0 = clk();
#pragma unroll 16
        for (int s = 0; s < N; ++s) x = vfma(x, c0, a1);
        t1 = clk();
        if (lane == 0) cyc[4] = t1 - t0;
        t0 = clk();
#pragma unroll 16
        for (int s = 0; s < N; ++s) x = shup(x);
        t1 = clk();
        if (lane == 0) cyc[5] = t1 - t0;
        t0 = clk();
#pragma unroll 16
        for (int s = 0; s < N; ++s) x = vfma(shup(x), c0, a1);
        t1 = clk();
        if (lane == 0) cyc[6] = t1 - t0;
        t0 = clk();
#pragma unroll 16
        for (int s = 0; s < N; ++s) { double d = isLC ? rv : shup(x); x = vfma(d, c0, a1); }
        t1 = clk();
        if (lane == 0) cyc[7] = t1 - t0;
        out[128 + lane] = x;
    }
}
int main() {
    double *out, *coef; long long* cyc;
    cudaMalloc(&out, 256 * 8); cudaMallocManaged(&coef, 256 * 8); cudaMallocManaged(&cyc, 8 * 8);
    for (int i = 0; i < 256; ++i) coef[i] = 0.1 + 0.001 * (i % 37);
    for (int r = 0; r < 2; ++r) { k<<<1, 32>>>(out, cyc, coef); cudaDeviceSynchronize(); }
    const char* names[] = {"A today (shfl, sel, 2 DFMA)", "B flat, select", "C flat, predicated pair", "D flat + coefficient work + LDS/STS",
                           "DFMA chain", "SHFL.f64 chain", "SHFL+DFMA chain", "SHFL+SEL+DFMA chain"};
    for (int i = 0; i < 8; ++i) printf("%-40s %.2f cyc/step\n", names[i], cyc[i] / (double)N);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
