// issue cost of back-to-back INDEPENDENT instructions from one warp on B200 (inline PTX so nothing is fused away)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk(){ long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
#define REP4(x) x x x x
#define REP16(x) REP4(x) REP4(x) REP4(x) REP4(x)
template <int V>
__global__ void k(double* out, long long* cyc, int zero) {
    __shared__ double sm[4096];
    const int lane = threadIdx.x;
    for (int i = lane; i < 4096; i += 32) sm[i] = i * 1e-3;
    __syncwarp();
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + lane * 8 + zero;
    unsigned base16 = (unsigned)__cvta_generic_to_shared(sm) + lane * 16 + zero;
    double a0 = lane, a1 = 1, a2 = 2, a3 = 3, b0 = 1.0 + zero, b1 = 0.5;
    float f0 = lane, f1 = 1, f2 = 2, f3 = 3;
    long long best = 1 << 30;
    for (int rep = 0; rep < 5; ++rep) {
        long long t0 = clk();
        if (V == 0) {  // 64 x LDS.64, 4 destination registers round-robin
            REP16(asm volatile("ld.shared.f64 %0, [%4];\n\tld.shared.f64 %1, [%4+256];\n\tld.shared.f64 %2, [%4+512];\n\tld.shared.f64 %3, [%4+768];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3) : "r"(base));)
        } else if (V == 1) {  // 64 x LDS.128
            REP16(asm volatile("ld.shared.v2.f64 {%0,%1}, [%4];\n\tld.shared.v2.f64 {%2,%3}, [%4+512];\n\tld.shared.v2.f64 {%0,%1}, [%4+1024];\n\tld.shared.v2.f64 {%2,%3}, [%4+1536];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3) : "r"(base16));)
        } else if (V == 2) {  // 64 x STS.64
            REP16(asm volatile("st.shared.f64 [%4], %0;\n\tst.shared.f64 [%4+256], %1;\n\tst.shared.f64 [%4+512], %2;\n\tst.shared.f64 [%4+768], %3;" :: "d"(a0), "d"(a1), "d"(a2), "d"(a3), "r"(base) : "memory");)
        } else if (V == 3) {  // 64 x SHFL.32 independent
            REP16(asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;\n\tshfl.sync.up.b32 %1, %1, 1, 0, 0xffffffff;\n\tshfl.sync.up.b32 %2, %2, 1, 0, 0xffffffff;\n\tshfl.sync.up.b32 %3, %3, 1, 0, 0xffffffff;" : "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3));)
        } else if (V == 4) {  // 64 x DFMA, 4 independent chains
            REP16(asm volatile("fma.rn.f64 %0, %0, %4, %5;\n\tfma.rn.f64 %1, %1, %4, %5;\n\tfma.rn.f64 %2, %2, %4, %5;\n\tfma.rn.f64 %3, %3, %4, %5;" : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3) : "d"(b0), "d"(b1));)
        } else if (V == 5) {  // 32 x (LDS.64 + DFMA) alternating, independent
            REP16(asm volatile("ld.shared.f64 %0, [%4];\n\tfma.rn.f64 %2, %2, %5, %6;\n\tld.shared.f64 %1, [%4+256];\n\tfma.rn.f64 %3, %3, %5, %6;" : "=d"(a0), "=d"(a1), "+d"(a2), "+d"(a3) : "r"(base), "d"(b0), "d"(b1));)
        } else if (V == 6) {  // 64 x FFMA independent
            REP16(asm volatile("fma.rn.f32 %0, %0, %4, %5;\n\tfma.rn.f32 %1, %1, %4, %5;\n\tfma.rn.f32 %2, %2, %4, %5;\n\tfma.rn.f32 %3, %3, %4, %5;" : "+f"(f0), "+f"(f1), "+f"(f2), "+f"(f3) : "f"((float)b0), "f"((float)b1));)
        } else if (V == 7) {  // 64 x LDS.32
            REP16(asm volatile("ld.shared.f32 %0, [%4];\n\tld.shared.f32 %1, [%4+256];\n\tld.shared.f32 %2, [%4+512];\n\tld.shared.f32 %3, [%4+768];" : "=f"(f0), "=f"(f1), "=f"(f2), "=f"(f3) : "r"(base));)
        } else if (V == 8) {  // mix per "step": 3 LDS.64, 2 DFMA (dependent pair), 2 SHFL.32, 1 STS.64 ; 16 steps
            REP16(asm volatile("ld.shared.f64 %0, [%4];\n\tld.shared.f64 %1, [%4+256];\n\tld.shared.f64 %2, [%4+512];\n\tfma.rn.f64 %3, %3, %5, %6;\n\tfma.rn.f64 %3, %3, %5, %6;\n\tshfl.sync.up.b32 %7, %7, 1, 0, 0xffffffff;\n\tshfl.sync.up.b32 %8, %8, 1, 0, 0xffffffff;\n\tst.shared.f64 [%4+1024], %3;" : "=d"(a0), "=d"(a1), "=d"(a2), "+d"(a3), "+r"(base), "+d"(b0), "+d"(b1), "+f"(f0), "+f"(f1) :: "memory");)
        }
        long long t1 = clk();
        if (t1 - t0 < best) best = t1 - t0;
    }
    if (lane == 0) cyc[0] = best;
    out[lane] = a0 + a1 + a2 + a3 + f0 + f1 + f2 + f3 + b0;
}
template <int V> void run(const char* name, int n, double* out, long long* cyc) {
    k<V><<<1, 32>>>(out, cyc, 0); cudaDeviceSynchronize();
    printf("%-44s %.2f cyc/instr (%lld total)\n", name, cyc[0] / (double)n, cyc[0]);
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 8);
    run<0>("LDS.64 x64", 64, out, cyc); run<1>("LDS.128 x64", 64, out, cyc); run<7>("LDS.32 x64", 64, out, cyc); run<2>("STS.64 x64", 64, out, cyc);
    run<3>("SHFL.32 x64", 64, out, cyc); run<4>("DFMA x64 (4 chains)", 64, out, cyc); run<6>("FFMA x64 (4 chains)", 64, out, cyc);
    run<5>("LDS.64+DFMA alternating x64", 64, out, cyc); run<8>("solver step mix (8 instr) x16", 128, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
