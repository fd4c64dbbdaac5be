// Standalone timing/validation harness for sd::waveKernel (tools only, not part of the library).
// usage: wavebench [nx ny sigma reps]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>
#include "../fluid-sim_b200/csrc/sdwave.cuh"
void fsim_set_error(const char*, ...) {}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

struct OpFwd {
    static constexpr int NIN = 4, NOUT = 1;
    const double* in[4]; double* out[1]; double* partials;
    template <int SIGMA>
    __device__ __forceinline__ void cell(const double (&v)[4], double left, double down, double (&o)[1], double& y, double& acc) const {
        double t = SIGMA == 1 ? __fma_rn(-v[2], down, __fma_rn(-v[1], left, v[0])) : __fma_rn(-v[1], left, __fma_rn(-v[2], down, v[0]));
        double w = v[3] * t; o[0] = w; acc = __fma_rn(t, w, acc); y = t;
    }
    __device__ void stripDone(int k, double acc) const { partials[k] = acc; }
    __device__ void allDone(int) const {}
};
struct OpBwd {
    static constexpr int NIN = 3, NOUT = 1;
    const double* in[3]; double* out[1];
    template <int SIGMA>
    __device__ __forceinline__ void cell(const double (&v)[3], double left, double down, double (&o)[1], double& y, double& acc) const {
        double z = SIGMA == 1 ? __fma_rn(-v[2], down, __fma_rn(-v[1], left, v[0])) : __fma_rn(-v[1], left, __fma_rn(-v[2], down, v[0])); o[0] = z; y = z;
    }
    __device__ void stripDone(int, double) const {}
    __device__ void allDone(int) const {}
};
struct SFwd {
    static constexpr int NIN = 4, KIND = 1;
    __device__ double postScalar() const { return 0.0; }
    const double* in[4]; double* out; double* partials;
    __device__ void stripDone(int k, double acc) const { partials[k] = acc; }
    __device__ void allDone(int) const {}
};
struct SBwd {
    static constexpr int NIN = 3, KIND = 0;
    __device__ double postScalar() const { return 0.0; }
    const double* in[3]; double* out;
    __device__ void stripDone(int, double) const {}
    __device__ void allDone(int) const {}
};
static int g_cl = 8;
template <class Op, int R, int SG, int DIR, int SUBS>
float runSolveR(const Op& f, const sd::Geom& g, sd::Control c, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    CK((sd::launchSolveR<Op, R, SG, DIR, SUBS>(f, g, c, 0, g_cl)));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) CK((sd::launchSolveR<Op, R, SG, DIR, SUBS>(f, g, c, 0, g_cl)));
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}
template <class Op, int SG, int DIR, int SUBS>
float runSolve(const Op& f, const sd::Geom& g, sd::Control c, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    CK((sd::launchSolve<Op, SG, DIR, SUBS>(f, g, c, 0, g_cl)));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) CK((sd::launchSolve<Op, SG, DIR, SUBS>(f, g, c, 0, g_cl)));
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}
__global__ void fillSent(unsigned long long* p, size_t n) { for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) p[k] = sd::SENT; }

template <int SG>
float runFwd(const OpFwd& f, const sd::Geom& g, sd::Control c, int reps) {
    size_t bytes = sd::Layout<OpFwd>::BYTES;
    CK(cudaFuncSetAttribute(sd::waveKernel<OpFwd, SG, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    sd::waveKernel<OpFwd, SG, 1><<<g.nstrips, 32, bytes>>>(f, g, c);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) sd::waveKernel<OpFwd, SG, 1><<<g.nstrips, 32, bytes>>>(f, g, c);
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}
template <int SG>
float runBwd(const OpBwd& f, const sd::Geom& g, sd::Control c, int reps) {
    size_t bytes = sd::Layout<OpBwd>::BYTES;
    CK(cudaFuncSetAttribute(sd::waveKernel<OpBwd, SG, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    sd::waveKernel<OpBwd, SG, -1><<<g.nstrips, 32, bytes>>>(f, g, c);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) sd::waveKernel<OpBwd, SG, -1><<<g.nstrips, 32, bytes>>>(f, g, c);
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main(int argc, char** argv) {
    int nx = argc > 1 ? atoi(argv[1]) : 4096, ny = argc > 2 ? atoi(argv[2]) : 4096, sigma = argc > 3 ? atoi(argv[3]) : 2, reps = argc > 4 ? atoi(argv[4]) : 20;
    const int rpl = argc > 8 ? atoi(argv[8]) : 1;  // rows per lane (solveKernelR, sigma 1)
    sd::Geom g = sd::makeGeom(nx, ny, sigma, rpl);
    printf("nx %d ny %d sigma %d: nstrips %d Sp %d elems %zu\n", nx, ny, sigma, g.nstrips, g.Sp, g.elems);
    // random coefficients on the logical grid (row-major host), packed to SD on the host
    std::vector<double> r((size_t)nx * ny), lx(r.size()), ly(r.size()), d(r.size());
    srand(1);
    for (size_t k = 0; k < r.size(); ++k) {
        r[k] = rand() / (double)RAND_MAX - 0.5; lx[k] = -0.3 * (rand() / (double)RAND_MAX); ly[k] = -0.3 * (rand() / (double)RAND_MAX);
        d[k] = 0.5 + rand() / (double)RAND_MAX;
    }
    // optional: random fluid column range per strip (everything outside is zero), argv[7] = 1
    const bool useRanges = argc > 7 && atoi(argv[7]) != 0;
    std::vector<int> hrange(2 * g.nstrips);
    for (int k = 0; k < g.nstrips; ++k) { hrange[2 * k] = 0; hrange[2 * k + 1] = g.nchunks - 1; }
    if (useRanges) {
        for (int k = 0; k < g.nstrips; ++k) {
            int lo = rand() % nx, hi = rand() % nx;
            if (lo > hi) std::swap(lo, hi);
            int kind = rand() % 8;
            if (atoi(argv[7]) == 2) { kind = 3; lo = 0; hi = 3 * nx / 4 - 32 * rpl * k - 1; if (hi < 0) { lo = 1; hi = 0; } }  // dam break: i + j < 3n/4
            if (kind == 0) { lo = 1; hi = 0; }                 // empty strip
            else if (kind == 1) { lo = 0; hi = nx - 1; }       // full strip
            else if (kind == 2) { hi = lo + rand() % 40; if (hi >= nx) hi = nx - 1; }  // narrow
            for (int j = 32 * rpl * k; j < 32 * rpl * (k + 1) && j < ny; ++j)
                for (int i = 0; i < nx; ++i)
                    if (i < lo || i > hi) { size_t o = (size_t)j * nx + i; r[o] = 0; lx[o] = 0; ly[o] = 0; d[o] = 0; }
            if (lo > hi) { hrange[2 * k] = 1; hrange[2 * k + 1] = 0; }
            else { hrange[2 * k] = lo / sd::CH; hrange[2 * k + 1] = (hi + 31 * sigma) / sd::CH; }
        }
    }
    auto pack = [&](const std::vector<double>& a) { std::vector<double> s(g.elems, 0.0); for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) s[sd::sdIndex(g, i, j)] = a[(size_t)j * nx + i]; return s; };
    std::vector<double> sr = pack(r), slx = pack(lx), sly = pack(ly), sdd = pack(d);
    double *dR, *dLx, *dLy, *dD, *dT, *dZ, *dPart; unsigned long long* hand; int* tick;
    size_t B = g.elems * 8;
    CK(cudaMalloc(&dR, B)); CK(cudaMalloc(&dLx, B)); CK(cudaMalloc(&dLy, B)); CK(cudaMalloc(&dD, B)); CK(cudaMalloc(&dT, B)); CK(cudaMalloc(&dZ, B));
    CK(cudaMalloc(&dPart, 8 * 4096)); CK(cudaMalloc(&hand, sd::handWords(g) * 8)); CK(cudaMalloc(&tick, 16)); CK(cudaMemset(tick, 0, 16));
    fillSent<<<256, 256>>>(hand, sd::handWords(g));
    CK(cudaMemcpy(dR, sr.data(), B, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dLx, slx.data(), B, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dLy, sly.data(), B, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dD, sdd.data(), B, cudaMemcpyHostToDevice));
    CK(cudaMemset(dT, 0, B)); CK(cudaMemset(dZ, 0, B));
    long long* prof; CK(cudaMallocManaged(&prof, 4 * 8 * (g.nstrips + 1))); memset(prof, 0, 4 * 8 * (g.nstrips + 1));
    int* dRange; CK(cudaMalloc(&dRange, hrange.size() * sizeof(int))); CK(cudaMemcpy(dRange, hrange.data(), hrange.size() * sizeof(int), cudaMemcpyHostToDevice));
    sd::Control c{tick, tick + 1, hand, nullptr, prof, useRanges ? dRange : nullptr};
    c.dbg = getenv("WB_DBG") ? atoi(getenv("WB_DBG")) : 0;  // 64: the legacy solver step
    OpFwd f; f.in[0] = dR; f.in[1] = dLx; f.in[2] = dLy; f.in[3] = dD; f.out[0] = dT; f.partials = dPart;
    OpBwd b; b.in[0] = dT; b.in[1] = dLx; b.in[2] = dLy; b.out[0] = dZ;
    float msF = 0, msB = 0;
    int mode = argc > 5 ? atoi(argv[5]) : 0; g_cl = argc > 6 ? atoi(argv[6]) : 8;  // 0: one-warp kernel; 8 / 4 / 16: warp-specialised kernel with that SUBS
    if (mode) {
        SFwd sf; sf.in[0] = dR; sf.in[1] = dLx; sf.in[2] = dLy; sf.in[3] = dD; sf.out = dT; sf.partials = dPart;
        SBwd sb; sb.in[0] = dT; sb.in[1] = dLx; sb.in[2] = dLy; sb.out = dZ;
#define RUNSR(RR, SG, SU) { msF = runSolveR<SFwd, RR, SG, 1, SU>(sf, g, c, reps); msB = runSolveR<SBwd, RR, SG, -1, SU>(sb, g, c, reps); }
#define RUNS(SG, SU) { msF = runSolve<SFwd, SG, 1, SU>(sf, g, c, reps); msB = runSolve<SBwd, SG, -1, SU>(sb, g, c, reps); }
        if (rpl == 2 && sigma == 1 && mode == 16) RUNSR(2, 1, 16) else if (rpl == 2 && sigma == 1 && mode == 8) RUNSR(2, 1, 8) else if (rpl == 2 && sigma == 2 && mode == 16) RUNSR(2, 2, 16)
        else if (rpl == 2 && sigma == 2 && mode == 8) RUNSR(2, 2, 8)

        else if (sigma == 2 && mode == 8) RUNS(2, 8) else if (sigma == 3 && mode == 8) RUNS(3, 8)

        else if (sigma == 2 && mode == 16) RUNS(2, 16) else if (sigma == 3 && mode == 16) RUNS(3, 16)
        else { printf("unsupported sigma/mode\n"); return 1; }
    } else
    switch (sigma) {
        case 1: msF = runFwd<1>(f, g, c, reps); msB = runBwd<1>(b, g, c, reps); break;
        case 2: msF = runFwd<2>(f, g, c, reps); msB = runBwd<2>(b, g, c, reps); break;
        case 3: msF = runFwd<3>(f, g, c, reps); msB = runBwd<3>(b, g, c, reps); break;
        case 4: msF = runFwd<4>(f, g, c, reps); msB = runBwd<4>(b, g, c, reps); break;
    }
#ifdef SD_PROFILE
    if (mode) {
        CK(cudaDeviceSynchronize());
        long long t0 = prof[3];
        for (int q : {0, 1, 2, 7, 8, 9, g.nstrips / 2, g.nstrips - 1}) if (q < g.nstrips) printf("  bwd strip#%d: total %lld cyc, tma-wait %lld, ready-wait %lld, start +%lld, end +%lld\n", q, prof[4 * q], prof[4 * q + 1], prof[4 * q + 2], prof[4 * q + 3] - t0, prof[4 * q + 3] - t0 + prof[4 * q]);
    }
    if (!mode) {
    // the profile holds the last launch (backward); rerun forward once for its profile
    auto show = [&](const char* nm) { long long t0 = prof[3]; for (int q : {0, 1, g.nstrips / 2, g.nstrips - 1}) if (q < g.nstrips) printf("  %s strip#%d: total %lld cyc, tma-wait %lld, poll-wait %lld, start +%lld\n", nm, q, prof[4 * q], prof[4 * q + 1], prof[4 * q + 2], prof[4 * q + 3] - t0); };
    show("bwd");
    switch (sigma) { case 1: runFwd<1>(f, g, c, 1); break; case 2: runFwd<2>(f, g, c, 1); break; case 3: runFwd<3>(f, g, c, 1); break; case 4: runFwd<4>(f, g, c, 1); break; }
    show("fwd");
    }
#endif
    // validate forward on the host
    std::vector<double> t((size_t)nx * ny), w(t.size()), z(t.size());
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
        size_t o = (size_t)j * nx + i;
        double left = i > 0 ? t[o - 1] : 0.0, down = j > 0 ? t[o - nx] : 0.0;
        t[o] = (sigma == 1 || rpl > 1) ? fma(-ly[o], down, fma(-lx[o], left, r[o])) : fma(-lx[o], left, fma(-ly[o], down, r[o])); w[o] = d[o] * t[o];
    }
    for (int j = ny - 1; j >= 0; --j) for (int i = nx - 1; i >= 0; --i) {
        size_t o = (size_t)j * nx + i;
        double left = i < nx - 1 ? z[o + 1] : 0.0, down = j < ny - 1 ? z[o + nx] : 0.0;
        z[o] = (sigma == 1 || rpl > 1) ? fma(-ly[o], down, fma(-lx[o], left, w[o])) : fma(-lx[o], left, fma(-ly[o], down, w[o]));
    }
    std::vector<double> gt(g.elems), gz(g.elems);
    CK(cudaMemcpy(gt.data(), dT, B, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(gz.data(), dZ, B, cudaMemcpyDeviceToHost));
    size_t badF = 0, badB = 0;
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) { size_t o = (size_t)j * nx + i, q = sd::sdIndex(g, i, j); if (gt[q] != w[o]) { if (badF < 12) printf("  fwd mismatch i=%d j=%d got %.17g want %.17g\n", i, j, gt[q], w[o]); ++badF; } if (gz[q] != z[o]) { if (badB < 10) printf("  bwd mismatch i=%d j=%d got %.17g want %.17g\n", i, j, gz[q], z[o]); ++badB; } }
    {   // largest backward errors
        struct E { double e; int i, j; };
        std::vector<E> es;
        for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) { size_t o = (size_t)j * nx + i, q = sd::sdIndex(g, i, j); double e = fabs(gz[q] - z[o]); if (e > 1e-9) es.push_back({e, i, j}); }
        printf("  bwd cells with |err| > 1e-9: %zu\n", es.size());
        { int cnt = 0; for (auto& e : es) if ((e.j & 31) == 31 && cnt < 40) { printf("    LC row: i=%d j=%d step %d ls %d err %.3e\n", e.i, e.j, e.i + sigma * 31, (e.i + sigma * 31) & 31, e.e); ++cnt; } }
        for (size_t k = 0; k < es.size() && k < 12; ++k) printf("    i=%d j=%d (lane %d, step %d, ls %d) err %.3e\n", es[k].i, es[k].j, es[k].j & 31, es[k].i + sigma * (es[k].j & 31), (es[k].i + sigma * (es[k].j & 31)) & 31, es[k].e);
        std::vector<unsigned long long> hh(sd::handWords(g));
        CK(cudaMemcpy(hh.data(), hand, hh.size() * 8, cudaMemcpyDeviceToHost));
        size_t hs = sd::handStride(g); size_t dirty = 0;
        for (int q = 0; q < g.nstrips - 1; ++q) for (int sl = 31 * sigma; sl < nx + 31 * sigma; ++sl) if (hh[q * hs + sl] != sd::SENT) { if (dirty < 5) printf("    dirty slot region %d slot %d\n", q, sl); ++dirty; }
        printf("  dirty polled slots after the run: %zu\n", dirty);
    }
    {   // the flat solver step (solveKernelR, R = 2, SIGMA = 1) differs from the host order by rounding: bound it
        double eF = 0, eB = 0; size_t tolF = 0, tolB = 0;
        for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
            size_t o = (size_t)j * nx + i, q = sd::sdIndex(g, i, j);
            double a = fabs(gt[q] - w[o]) / (1.0 + fabs(w[o])), b = fabs(gz[q] - z[o]) / (1.0 + fabs(z[o]));
            if (!(a <= 1e-13)) ++tolF; if (!(b <= 1e-13)) ++tolB;
            if (a > eF || a != a) eF = a; if (b > eB || b != b) eB = b;
        }
        printf("  max error / (1 + |want|): forward %.3e backward %.3e; cells above 1e-13: %zu %zu\n", eF, eB, tolF, tolB);
    }
    double cells = (double)nx * ny;
    printf("forward  %.4f ms  (%.0f GB/s at 40 B/cell)  mismatches %zu\n", msF, cells * 40 / msF / 1e6, badF);
    printf("backward %.4f ms  (%.0f GB/s at 32 B/cell)  mismatches %zu\n", msB, cells * 32 / msB / 1e6, badB);
    return 0;
}
