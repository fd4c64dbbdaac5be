// Device-side samplers of the staggered grid: Catmull-Rom bicubic (reference include/Array2D.h:244-359),
// bilinear gather (include/Array2D.h:402-420), MAC position mapping (include/MACGrid2D.h:80-98),
// clampPos (src/FluidSim2D.cpp:645-651) and the Ralston RK3 used by both advections (:213-216, :591-595).
#pragma once

#include "common.cuh"

struct GridView {
    const double* u;  // (nx+1) x ny, pointer at (0,0)
    const double* v;  // nx x (ny+1)
    int nx, ny, pitch;
    double dx;
};

template <bool CG>
__device__ __forceinline__ double ldGrid(const double* p) { return CG ? __ldcg(p) : __ldg(p); }

// a(x,y) sampled with Catmull-Rom weights; NX,NY are the array's own extents (index clamp), values outside
// the array give 0 exactly like the reference.
template <bool CG>
__device__ __forceinline__ double bicubic(const double* a, int pitch, int NX, int NY, double px, double py) {
    int x = (int)px, y = (int)py;  // truncation toward zero, as in the reference
    if (x < 0 || x >= NX || y < 0 || y >= NY) return 0.0;
    double fx = px - (double)x, fy = py - (double)y;
    double fx2 = fx * fx, fx3 = fx * fx * fx, fy2 = fy * fy, fy3 = fy * fy * fy;
    double wu[4], wv[4];
    wu[0] = -0.5 * fx3 + fx2 - 0.5 * fx;
    wu[1] = 1.5 * fx3 - 2.5 * fx2 + 1;
    wu[2] = -1.5 * fx3 + 2 * fx2 + 0.5 * fx;
    wu[3] = 0.5 * fx3 - 0.5 * fx2;
    wv[0] = -0.5 * fy3 + fy2 - 0.5 * fy;
    wv[1] = 1.5 * fy3 - 2.5 * fy2 + 1;
    wv[2] = -1.5 * fy3 + 2 * fy2 + 0.5 * fy;
    wv[3] = 0.5 * fy3 - 0.5 * fy2;
    int xs[4], ys[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        xs[k] = iclampd(x - 1 + k, 0, NX - 1);
        ys[k] = iclampd(y - 1 + k, 0, NY - 1);
    }
    double row[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double* r = a + (long long)ys[j] * pitch;
        double a0 = ldGrid<CG>(r + xs[0]), a1 = ldGrid<CG>(r + xs[1]), a2 = ldGrid<CG>(r + xs[2]), a3 = ldGrid<CG>(r + xs[3]);
        row[j] = (wu[0] * a0 + wu[1] * a1) + (wu[2] * a2 + wu[3] * a3);
    }
    return (row[0] * wv[0] + row[2] * wv[2]) + (row[1] * wv[1] + row[3] * wv[3]);
}

template <bool CG>
__device__ __forceinline__ double sampleU(const GridView& g, double x, double y) {
    x /= g.dx; y /= g.dx;
    y -= 0.5;
    x = amlClamp(x, 1e-6, (double)(g.nx - 1) - 1e-6);
    y = amlClamp(y, 1e-6, (double)(g.ny - 1) - 1e-6);
    return bicubic<CG>(g.u, g.pitch, g.nx + 1, g.ny, x, y);
}
template <bool CG>
__device__ __forceinline__ double sampleV(const GridView& g, double x, double y) {
    x /= g.dx; y /= g.dx;
    x -= 0.5;
    x = amlClamp(x, 1e-6, (double)(g.nx - 1) - 1e-6);
    y = amlClamp(y, 1e-6, (double)(g.ny - 1) - 1e-6);
    return bicubic<CG>(g.v, g.pitch, g.nx, g.ny + 1, x, y);
}

__device__ __forceinline__ void clampPos(int nx, int ny, double dx, double& x, double& y) {
    const double offset = 1e-3;
    x = amlClamp(x, (1.0 + offset) * dx, (nx - 1.0 - offset) * dx);
    y = amlClamp(y, (1.0 + offset) * dx, (ny - 1.0 - offset) * dx);
}

// Ralston RK3 through the bicubic sampler. SIGN=-1: backtrace, SIGN=+1: forward. Stage positions are only
// clamped by the sampler; the caller applies clampPos to the result.
template <int SIGN, bool CG>
__device__ __forceinline__ void rk3(const GridView& g, double dt, double x, double y, double& ox, double& oy) {
    double k1x = sampleU<CG>(g, x, y), k1y = sampleV<CG>(g, x, y);
    double x2 = x + SIGN * (0.5 * dt * k1x), y2 = y + SIGN * (0.5 * dt * k1y);
    double k2x = sampleU<CG>(g, x2, y2), k2y = sampleV<CG>(g, x2, y2);
    double x3 = x + SIGN * (0.75 * dt * k2x), y3 = y + SIGN * (0.75 * dt * k2y);
    double k3x = sampleU<CG>(g, x3, y3), k3y = sampleV<CG>(g, x3, y3);
    if (SIGN < 0) {
        ox = x - ((2. / 9.) * dt * k1x + (3. / 9.) * dt * k2x + (4. / 9.) * dt * k3x);
        oy = y - ((2. / 9.) * dt * k1y + (3. / 9.) * dt * k2y + (4. / 9.) * dt * k3y);
    } else {
        ox = x + (2. / 9.) * dt * k1x + (3. / 9.) * dt * k2x + (4. / 9.) * dt * k3x;
        oy = y + (2. / 9.) * dt * k1y + (3. / 9.) * dt * k2y + (4. / 9.) * dt * k3y;
    }
}

// bilinear footprint shared by splat and gather (include/Array2D.h:361-376, 402-420)
struct Bilinear {
    int x1, x2, y1, y2;
    double fx, fy;
};
__device__ __forceinline__ Bilinear bilinearAt(double px, double py, int NX, int NY) {
    Bilinear b;
    int ui = (int)px, uj = (int)py;
    b.fx = px - ui; b.fy = py - uj;
    b.x1 = iclampd(ui, 0, NX - 1); b.x2 = iclampd(ui + 1, 0, NX - 1);
    b.y1 = iclampd(uj, 0, NY - 1); b.y2 = iclampd(uj + 1, 0, NY - 1);
    return b;
}
