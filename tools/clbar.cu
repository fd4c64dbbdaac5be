// micro-benchmark: cost per iteration of cluster-wide layer steps (tools only)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int CL = 8, TH = 512;
constexpr unsigned long long SENT = 0x7FF8F51D0DEAD002ULL;
template <int MODE>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(TH, 1) k(double* g, int iters) {
    __shared__ unsigned long long ring[2][TH];
    unsigned rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    ring[0][threadIdx.x] = SENT; ring[1][threadIdx.x] = SENT;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    unsigned la = (unsigned)__cvta_generic_to_shared(&ring[0][0]), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"((rank + 1) % CL));
    double v = threadIdx.x;
    const int tid = rank * TH + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const int par = it & 1;
        if (MODE >= 5 && it > 0) {  // read what the previous iteration pushed
            unsigned long long b; unsigned a = la + ((par ^ 1) * TH + threadIdx.x) * 8;
            do { asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(b) : "r"(a) : "memory"); } while (b == SENT);
            v += __longlong_as_double((long long)b) * 1e-9;
        }
        if (MODE == 3 || MODE == 4 || MODE == 6) __stcg(g + (size_t)(it & 1023) * CL * TH + tid, v);
        if (MODE >= 5) {
            asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra + (par * TH + threadIdx.x) * 8), "d"(v) : "memory");
            if (it > 0) { __syncthreads(); ring[par ^ 1][threadIdx.x] = SENT; }
        }
        if (MODE == 2 || MODE == 4) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        else asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
    }
    g[tid] = v;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int MODE> int run(double* g, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<CL, TH>>>(g, 10); CK(cudaDeviceSynchronize());
    cudaEventRecord(a); k<MODE><<<CL, TH>>>(g, iters); cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("mode %d: %.3f us per iteration\n", MODE, ms * 1e3 / iters);
    return 0;
}
int main() {
    double* g; CK(cudaMalloc(&g, sizeof(double) * 1024 * CL * TH));
    int iters = 5000;
    run<1>(g, iters); run<2>(g, iters); run<3>(g, iters); run<4>(g, iters); run<5>(g, iters); run<6>(g, iters);
    return 0;
}
