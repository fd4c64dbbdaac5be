/* TEST INFRASTRUCTURE ONLY (see fsim_oracle.h).  Plain-C restatement of the fluid-sim hot path.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 * Arithmetic is written in the same association order as the reference so that, compiled with
 * the reference's flags (-O2 -mavx -mfma), results agree to rounding; tests/test_oracle.py pins
 * it against the stock reference build and the golden fixtures.
 *
 * Conventions: a(i,j) = data[j*NX+i] (include/Array2D.h:43,87); u is (nx+1) x ny, v is nx x (ny+1)
 * (include/MACGrid2D.h:19-25); particles are AoS {x,y} doubles (deps/altmath/src/vec2.h:11-28).
 */
#include "fsim_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { CELL_EMPTY = 0, CELL_FLUID = 1, CELL_SOLID = 2 }; /* include/FluidSim2D.h:44-46 */

typedef struct {
    int nx, ny, ppcSqrt, mode; /* mode 0 = FS_SEMILAGRANGIAN, 1 = FS_PICFLIP */
    double dt, dx, dr, rho, gx, gy, alpha;
    double *u, *v, *nu, *nv, *p, *phi;
    uint8_t* cell;
    double *pos, *vel;
    long np;
    double *Adiag, *Ax, *Ay, *rhs, *precon;
    double waterVolume, totalEnergy, particleTotalEnergy, currentTime;
    int lastIters;
    float stageMs[8];
    int nStageMs;
} Sim;

static double g_tol = 1e-12; /* src/FluidSim2D.cpp:453 */
static int g_maxIters = 200; /* src/FluidSim2D.cpp:429 */
static int g_slDoubleBuffer = 0;

static inline double dmin(double a, double b) { return (a < b) ? a : b; }   /* deps/altmath/src/math_utils.h:21-23 */
static inline double dmax(double a, double b) { return (a > b) ? a : b; }   /* :16-18 */
static inline double dclamp(double v, double lo, double hi) { return dmax(lo, dmin(v, hi)); } /* :46-48 */
static inline int imin(int a, int b) { return (a < b) ? a : b; }
static inline int imax(int a, int b) { return (a > b) ? a : b; }
static inline int iclamp(int v, int lo, int hi) { return imax(lo, imin(v, hi)); }

static double* dalloc(size_t n) {
    double* p = (double*)aligned_alloc(32, ((n * sizeof(double) + 31) / 32) * 32);
    memset(p, 0, n * sizeof(double));
    return p;
}

/* ---------------------------------------------------------------- Array2D<double> kernels */

/* include/Array2D.h:244-327 (AVX path): Catmull-Rom weights, index clamp, 4x4 gather.
 * Summation order of the AVX code: row_j = (w0 a0 + w1 a1) + (w2 a2 + w3 a3) (hadd pairs, then add),
 * result = (row0 v0 + row2 v2) + (row1 v1 + row3 v3) (low/high 128-bit halves added, then swapped add). */
static double bicubic(const double* a, int NX, int NY, double px, double py) {
    int x = (int)px, y = (int)py;
    if (x < 0 || x >= NX || y < 0 || y >= NY) return 0;
    double fx = px - (double)x, fy = py - (double)y;
    double wu[4], wv[4];
    wu[0] = -0.5 * (fx * fx * fx) + (fx * fx) - 0.5 * fx;
    wu[1] = 1.5 * (fx * fx * fx) - 2.5 * (fx * fx) + 1;
    wu[2] = -1.5 * (fx * fx * fx) + 2 * (fx * fx) + 0.5 * fx;
    wu[3] = 0.5 * (fx * fx * fx) - 0.5 * (fx * fx);
    wv[0] = -0.5 * (fy * fy * fy) + (fy * fy) - 0.5 * fy;
    wv[1] = 1.5 * (fy * fy * fy) - 2.5 * (fy * fy) + 1;
    wv[2] = -1.5 * (fy * fy * fy) + 2 * (fy * fy) + 0.5 * fy;
    wv[3] = 0.5 * (fy * fy * fy) - 0.5 * (fy * fy);
    int xs[4], ys[4];
    for (int k = 0; k < 4; k++) {
        xs[k] = iclamp(x - 1 + k, 0, NX - 1);
        ys[k] = iclamp(y - 1 + k, 0, NY - 1);
    }
    volatile double row[4]; /* volatile: keep each product/sum individually rounded like the vector code */
    for (int j = 0; j < 4; j++) {
        const double* r = a + (size_t)ys[j] * NX;
        volatile double m0 = wu[0] * r[xs[0]], m1 = wu[1] * r[xs[1]], m2 = wu[2] * r[xs[2]], m3 = wu[3] * r[xs[3]];
        volatile double s01 = m0 + m1, s23 = m2 + m3;
        row[j] = s01 + s23;
    }
    volatile double f0 = row[0] * wv[0], f1 = row[1] * wv[1], f2 = row[2] * wv[2], f3 = row[3] * wv[3];
    volatile double a02 = f0 + f2, a13 = f1 + f3;
    return a02 + a13;
}

/* include/Array2D.h:361-376 */
static void splat(double* a, int NX, int NY, double px, double py, double value) {
    int ui = (int)px, uj = (int)py;
    double fx = px - ui, fy = py - uj;
    int x1 = iclamp(ui, 0, NX - 1), x2 = iclamp(ui + 1, 0, NX - 1);
    int y1 = iclamp(uj, 0, NY - 1), y2 = iclamp(uj + 1, 0, NY - 1);
    a[(size_t)y1 * NX + x1] += (1 - fx) * (1 - fy) * value;
    a[(size_t)y1 * NX + x2] += fx * (1 - fy) * value;
    a[(size_t)y2 * NX + x1] += (1 - fx) * fy * value;
    a[(size_t)y2 * NX + x2] += fx * fy * value;
}

/* include/Array2D.h:402-420 */
static double gather(const double* a, int NX, int NY, double px, double py) {
    int ui = (int)px, uj = (int)py;
    double fx = px - ui, fy = py - uj;
    int x1 = iclamp(ui, 0, NX - 1), x2 = iclamp(ui + 1, 0, NX - 1);
    int y1 = iclamp(uj, 0, NY - 1), y2 = iclamp(uj + 1, 0, NY - 1);
    double value = 0.0;
    value += a[(size_t)y1 * NX + x1] * (1 - fx) * (1 - fy);
    value += a[(size_t)y1 * NX + x2] * fx * (1 - fy);
    value += a[(size_t)y2 * NX + x1] * (1 - fx) * fy;
    value += a[(size_t)y2 * NX + x2] * fx * fy;
    return value;
}

/* include/Array2D.h:195-210 (AVX path): four lane accumulators, (l0+l1)+(l2+l3) */
static double dot(const double* a, const double* b, size_t n) {
    double l[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < n; i += 4)
        for (int k = 0; k < 4; k++) l[k] = l[k] + a[i + k] * b[i + k];
    return (l[0] + l[1]) + (l[2] + l[3]);
}

/* include/Array2D.h:220-233 */
static double norm_inf(const double* a, size_t n) {
    double l[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < n; i += 4)
        for (int k = 0; k < 4; k++) {
            double v = a[i + k], m = -v;
            double ab = (m > v) ? m : v;
            l[k] = (l[k] > ab) ? l[k] : ab;
        }
    return dmax(dmax(l[0], l[1]), dmax(l[2], l[3]));
}

/* include/Array2D.h:174-187: dst = a + b*c with one fused multiply-add per element */
static void axpy_to(double* dst, const double* a, double b, const double* c, size_t n) {
    for (size_t i = 0; i < n; i++) dst[i] = fma(b, c[i], a[i]);
}

/* include/Array2D.h:552-591: breadth-first extrapolation into the unknown (mask != 0) region */
static void extrapolate(double* a, uint32_t* mask, int NX, int NY) {
    size_t cap = (size_t)NX * NY + 1;
    int* qx = (int*)malloc(cap * sizeof(int));
    int* qy = (int*)malloc(cap * sizeof(int));
    size_t head = 0, tail = 0;
    for (int j = 0; j < NY; j++)
        for (int i = 0; i < NX; i++) {
            if (mask[(size_t)j * NX + i] != 0 &&
                ((i > 0 && mask[(size_t)j * NX + i - 1] == 0) || (i < NX - 1 && mask[(size_t)j * NX + i + 1] == 0) ||
                 (j > 0 && mask[(size_t)(j - 1) * NX + i] == 0) || (j < NY - 1 && mask[(size_t)(j + 1) * NX + i] == 0))) {
                mask[(size_t)j * NX + i] = 1;
                qx[tail] = i; qy[tail] = j; tail++;
            }
        }
    static const int ox[4] = {-1, 1, 0, 0}, oy[4] = {0, 0, -1, 1};
    while (head < tail) {
        int x = qx[head], y = qy[head]; head++;
        uint32_t mine = mask[(size_t)y * NX + x];
        double sum = 0.0;
        int count = 0;
        for (int k = 0; k < 4; k++) {
            int xx = x + ox[k], yy = y + oy[k];
            if (xx < 0 || xx >= NX || yy < 0 || yy >= NY) continue;
            if (mask[(size_t)yy * NX + xx] < mine) { sum += a[(size_t)yy * NX + xx]; count++; }
        }
        a[(size_t)y * NX + x] = count == 0 ? 0 : sum / count;
        for (int k = 0; k < 4; k++) {
            int xx = x + ox[k], yy = y + oy[k];
            if (xx < 0 || xx >= NX || yy < 0 || yy >= NY) continue;
            if (mask[(size_t)yy * NX + xx] == UINT32_MAX) {
                mask[(size_t)yy * NX + xx] = mine + 1;
                qx[tail] = xx; qy[tail] = yy; tail++;
            }
        }
    }
    free(qx); free(qy);
}

/* ---------------------------------------------------------------- MACGrid2D sampling */

/* include/MACGrid2D.h:80-86 */
static double sample_u(const Sim* s, const double* u, double x, double y) {
    x /= s->dx; y /= s->dx;
    y -= 0.5;
    x = dclamp(x, 1e-6, (double)((size_t)s->nx - 1) - 1e-6);
    y = dclamp(y, 1e-6, (double)((size_t)s->ny - 1) - 1e-6);
    return bicubic(u, s->nx + 1, s->ny, x, y);
}
/* include/MACGrid2D.h:88-94 */
static double sample_v(const Sim* s, const double* v, double x, double y) {
    x /= s->dx; y /= s->dx;
    x -= 0.5;
    x = dclamp(x, 1e-6, (double)((size_t)s->nx - 1) - 1e-6);
    y = dclamp(y, 1e-6, (double)((size_t)s->ny - 1) - 1e-6);
    return bicubic(v, s->nx, s->ny + 1, x, y);
}

/* src/FluidSim2D.cpp:645-651 */
static void clamp_pos(const Sim* s, double* x, double* y) {
    const double offset = 1e-3;
    *x = dclamp(*x, (1.0 + offset) * s->dx, (s->nx - 1.0 - offset) * s->dx);
    *y = dclamp(*y, (1.0 + offset) * s->dx, (s->ny - 1.0 - offset) * s->dx);
}

/* Ralston RK3 through the bicubic MAC sampler; sign = -1 backtrace (src/FluidSim2D.cpp:213-216,
 * 226-229), sign = +1 particle advection (:591-595).  Stage positions are not clamp_pos'ed. */
static void rk3(const Sim* s, const double* u, const double* v, double x, double y, double sign, double* ox, double* oy) {
    double dt = s->dt;
    double k1x = sample_u(s, u, x, y), k1y = sample_v(s, v, x, y);
    if (sign < 0) {
        double x2 = x - 0.5 * dt * k1x, y2 = y - 0.5 * dt * k1y;
        double k2x = sample_u(s, u, x2, y2), k2y = sample_v(s, v, x2, y2);
        double x3 = x - 0.75 * dt * k2x, y3 = y - 0.75 * dt * k2y;
        double k3x = sample_u(s, u, x3, y3), k3y = sample_v(s, v, x3, y3);
        /* x - ((2/9 dt k1 + 3/9 dt k2) + 4/9 dt k3) */
        *ox = x - ((2. / 9.) * dt * k1x + (3. / 9.) * dt * k2x + (4. / 9.) * dt * k3x);
        *oy = y - ((2. / 9.) * dt * k1y + (3. / 9.) * dt * k2y + (4. / 9.) * dt * k3y);
    } else {
        double x2 = x + 0.5 * dt * k1x, y2 = y + 0.5 * dt * k1y;
        double k2x = sample_u(s, u, x2, y2), k2y = sample_v(s, v, x2, y2);
        double x3 = x + 0.75 * dt * k2x, y3 = y + 0.75 * dt * k2y;
        double k3x = sample_u(s, u, x3, y3), k3y = sample_v(s, v, x3, y3);
        /* ((x + 2/9 dt k1) + 3/9 dt k2) + 4/9 dt k3 */
        *ox = x + (2. / 9.) * dt * k1x + (3. / 9.) * dt * k2x + (4. / 9.) * dt * k3x;
        *oy = y + (2. / 9.) * dt * k1y + (3. / 9.) * dt * k2y + (4. / 9.) * dt * k3y;
    }
}

/* ---------------------------------------------------------------- level set */

/* one visit of the closest-particle propagation, src/FluidSim2D.cpp:775-792 */
static inline void lsVisit(Sim* s, size_t* t, int i, int j) {
    const int il[4] = {i - 1, i + 1, i, i}, jl[4] = {j, j, j - 1, j + 1};
    int nx = s->nx, ny = s->ny;
    double dx = s->dx;
    for (int k = 0; k < 4; k++) {
        if (il[k] < 0 || il[k] >= nx || jl[k] < 0 || jl[k] >= ny) continue;
        size_t e = t[(size_t)jl[k] * nx + il[k]];
        if (e != (size_t)-1) {
            double px = s->pos[2 * e], py = s->pos[2 * e + 1];
            double d = sqrt((px - i * dx) * (px - i * dx) + (py - j * dx) * (py - j * dx)) - s->dr;
            if (d < s->phi[(size_t)j * nx + i]) {
                s->phi[(size_t)j * nx + i] = d;
                t[(size_t)j * nx + i] = e;
            }
        }
    }
}

/* src/FluidSim2D.cpp:752-794 with the sweep order of include/FluidSim2D.h:178-203 */
static void levelset_construct(Sim* s) {
    int nx = s->nx, ny = s->ny;
    double dx = s->dx;
    size_t n = (size_t)nx * ny;
    size_t* t = (size_t*)malloc(n * sizeof(size_t));
    for (size_t k = 0; k < n; k++) { s->phi[k] = HUGE_VAL; t[k] = (size_t)-1; }
    for (long e = 0; e < s->np; e++) {
        double px = s->pos[2 * e], py = s->pos[2 * e + 1];
        int x = (int)(px / dx), y = (int)(py / dx);
        double d = sqrt((px - x * dx) * (px - x * dx) + (py - y * dx) * (py - y * dx)) - s->dr;
        if (d < s->phi[(size_t)y * nx + x]) { s->phi[(size_t)y * nx + x] = d; t[(size_t)y * nx + x] = (size_t)e; }
    }
    for (int k = 0; k < 4; k++) {
        for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) lsVisit(s, t, i, j);
        for (int j = 0; j < ny; j++) for (int i = nx; i-- > 0;) lsVisit(s, t, i, j);
        for (int j = ny; j-- > 0;) for (int i = 0; i < nx; i++) lsVisit(s, t, i, j);
        for (int j = ny; j-- > 0;) for (int i = nx; i-- > 0;) lsVisit(s, t, i, j);
    }
    free(t);
}

/* eikonal update of one negative cell from two upwind neighbours, src/FluidSim2D.cpp:848-857 */
static inline void eikonal(Sim* s, int i, int j, int ia, int ja, int ib, int jb) {
    int nx = s->nx;
    double dx = s->dx;
    double* phi = s->phi;
    if (phi[(size_t)j * nx + i] >= 0) return;
    double a = fabs(phi[(size_t)ja * nx + ia]), b = fabs(phi[(size_t)jb * nx + ib]);
    double phi0 = dmin(a, b), phi1 = dmax(a, b);
    double d = phi0 + dx;
    if (d > phi1) d = 0.5 * (phi0 + phi1 + sqrt(2 * dx * dx - (phi1 - phi0) * (phi1 - phi0)));
    if (d < -phi[(size_t)j * nx + i]) phi[(size_t)j * nx + i] = -d;
}

/* src/FluidSim2D.cpp:796-920 */
static void levelset_redistance(Sim* s) {
    int nx = s->nx, ny = s->ny;
    size_t n = (size_t)nx * ny;
    double* phi = s->phi;
    uint8_t* surf = (uint8_t*)calloc(n, 1);
    double* old = (double*)malloc(n * sizeof(double));
    memcpy(old, phi, n * sizeof(double));
    /* :803-834 -- sign change across the +x / +y edge marks both cells as surface; the phi
     * assignments in the reference re-store the value already there, so they are omitted */
    for (int j = 0; j < ny - 1; j++)
        for (int i = 0; i < nx - 1; i++) {
            double p0 = old[(size_t)j * nx + i], p1 = old[(size_t)j * nx + i + 1], p2 = old[(size_t)(j + 1) * nx + i];
            if (p0 * p1 < 0) { surf[(size_t)j * nx + i] = 1; surf[(size_t)j * nx + i + 1] = 1; }
            if (p0 * p2 < 0) { surf[(size_t)j * nx + i] = 1; surf[(size_t)(j + 1) * nx + i] = 1; }
        }
    /* :836-842 */
    for (size_t k = 0; k < n; k++)
        if (!surf[k] && phi[k] < 0) phi[k] = -HUGE_VAL;
    /* :845-902 */
    for (int k = 0; k < 4; k++) {
        for (int j = 1; j < ny; j++) for (int i = 1; i < nx; i++) eikonal(s, i, j, i - 1, j, i, j - 1);
        for (int j = 1; j < ny; j++) for (int i = nx - 1; i-- > 0;) eikonal(s, i, j, i + 1, j, i, j - 1);
        for (int j = ny - 1; j-- > 0;) for (int i = 1; i < nx; i++) eikonal(s, i, j, i - 1, j, i, j + 1);
        for (int j = ny - 1; j-- > 0;) for (int i = nx - 1; i-- > 0;) eikonal(s, i, j, i + 1, j, i, j + 1);
    }
    /* :905-919 two Jacobi smoothing passes on the interior */
    for (int k = 0; k < 2; k++) {
        memcpy(old, phi, n * sizeof(double));
        for (int j = 1; j < ny - 1; j++)
            for (int i = 1; i < nx - 1; i++) {
                double avg = 0.25 * (old[(size_t)j * nx + i - 1] + old[(size_t)j * nx + i + 1] +
                                     old[(size_t)(j - 1) * nx + i] + old[(size_t)(j + 1) * nx + i]);
                if (avg < old[(size_t)j * nx + i]) phi[(size_t)j * nx + i] = avg;
            }
    }
    free(surf); free(old);
}

/* src/FluidSim2D.cpp:653-732 */
static void stage_water_level_set(Sim* s) {
    int nx = s->nx, ny = s->ny;
    double dx = s->dx;
    levelset_construct(s);
    levelset_redistance(s);
    for (size_t k = 0; k < (size_t)nx * ny; k++)
        if (s->cell[k] != CELL_SOLID) s->cell[k] = s->phi[k] < 0.0 ? CELL_FLUID : CELL_EMPTY;
    /* :709-731 statistics; note they sample the grid left by the previous step */
    double vol = 0.0, en = 0.0, pen = 0.0;
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++)
            if (s->cell[(size_t)j * nx + i] == CELL_FLUID) {
                vol += dx * dx;
                double vx = sample_u(s, s->u, i * dx, j * dx), vy = sample_v(s, s->v, i * dx, j * dx);
                en += 0.5 * (s->rho * dx * dx) * (vx * vx + vy * vy);
                en -= (s->rho * dx * dx) * (s->gx * (i * dx) + s->gy * (j * dx));
            }
    int pp = s->ppcSqrt * s->ppcSqrt;
    for (long e = 0; e < s->np; e++) {
        double vx = s->vel[2 * e], vy = s->vel[2 * e + 1];
        pen += 0.5 * (s->rho * dx * dx / pp) * (vx * vx + vy * vy);
        pen -= (s->rho * dx * dx / pp) * (s->gx * s->pos[2 * e] + s->gy * s->pos[2 * e + 1]);
    }
    s->waterVolume = vol; s->totalEnergy = en; s->particleTotalEnergy = pen;
}

/* ---------------------------------------------------------------- particle <-> grid */

/* src/FluidSim2D.cpp:144-204 */
static void stage_p2g(Sim* s) {
    int nx = s->nx, ny = s->ny;
    double dx = s->dx;
    size_t nu = (size_t)(nx + 1) * ny, nv = (size_t)nx * (ny + 1);
    memset(s->u, 0, nu * sizeof(double));
    memset(s->v, 0, nv * sizeof(double));
    double* uw = dalloc(nu);
    double* vw = dalloc(nv);
    for (long e = 0; e < s->np; e++) {
        double px = s->pos[2 * e], py = s->pos[2 * e + 1];
        double ux = px / dx, uy = py / dx - 0.5;
        splat(s->u, nx + 1, ny, ux, uy, s->vel[2 * e]);
        splat(uw, nx + 1, ny, ux, uy, 1.0);
        double vx = px / dx - 0.5, vy = py / dx;
        splat(s->v, nx, ny + 1, vx, vy, s->vel[2 * e + 1]);
        splat(vw, nx, ny + 1, vx, vy, 1.0);
    }
    for (size_t k = 0; k < nu; k++) if (uw[k] > 0) s->u[k] /= uw[k];
    for (size_t k = 0; k < nv; k++) if (vw[k] > 0) s->v[k] /= vw[k];
    uint32_t* um = (uint32_t*)malloc(nu * sizeof(uint32_t));
    uint32_t* vm = (uint32_t*)malloc(nv * sizeof(uint32_t));
    for (size_t k = 0; k < nu; k++) um[k] = s->u[k] == 0.0 ? UINT32_MAX : 0;
    for (size_t k = 0; k < nv; k++) vm[k] = s->v[k] == 0.0 ? UINT32_MAX : 0;
    extrapolate(s->u, um, nx + 1, ny);
    extrapolate(s->v, vm, nx, ny + 1);
    free(um); free(vm); free(uw); free(vw);
}

/* src/FluidSim2D.cpp:552-568 */
static void stage_g2p(Sim* s) {
    int nx = s->nx, ny = s->ny;
    double dx = s->dx;
    size_t nu = (size_t)(nx + 1) * ny, nv = (size_t)nx * (ny + 1);
    double* du = dalloc(nu);
    double* dv = dalloc(nv);
    for (size_t k = 0; k < nu; k++) du[k] = s->nu[k] - s->u[k];
    for (size_t k = 0; k < nv; k++) dv[k] = s->nv[k] - s->v[k];
    double al = s->alpha;
    for (long e = 0; e < s->np; e++) {
        double px = s->pos[2 * e], py = s->pos[2 * e + 1];
        double ux = px / dx, uy = py / dx - 0.5;
        double vx = px / dx - 0.5, vy = py / dx;
        double picx = gather(s->nu, nx + 1, ny, ux, uy), picy = gather(s->nv, nx, ny + 1, vx, vy);
        double flipx = s->vel[2 * e] + gather(du, nx + 1, ny, ux, uy);
        double flipy = s->vel[2 * e + 1] + gather(dv, nx, ny + 1, vx, vy);
        s->vel[2 * e] = al * picx + (1 - al) * flipx;
        s->vel[2 * e + 1] = al * picy + (1 - al) * flipy;
    }
    memcpy(s->u, s->nu, nu * sizeof(double));
    memcpy(s->v, s->nv, nv * sizeof(double));
    free(du); free(dv);
}

/* src/FluidSim2D.cpp:570-604 (the CFL diagnostic only logs; omitted) */
static void stage_advect(Sim* s) {
    for (long e = 0; e < s->np; e++) {
        double x, y;
        rk3(s, s->u, s->v, s->pos[2 * e], s->pos[2 * e + 1], +1.0, &x, &y);
        clamp_pos(s, &x, &y);
        s->pos[2 * e] = x; s->pos[2 * e + 1] = y;
    }
}

/* src/FluidSim2D.cpp:206-235: in place, raster order; optional snapshot variant */
static void stage_sl_advect(Sim* s) {
    int nx = s->nx, ny = s->ny;
    double dx = s->dx;
    size_t nu = (size_t)(nx + 1) * ny, nv = (size_t)nx * (ny + 1);
    double *du = s->u, *dv = s->v;
    if (g_slDoubleBuffer) { du = s->nu; dv = s->nv; }
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx + 1; i++) {
            double x0 = (double)i * dx, y0 = ((double)j + 0.5) * dx, x, y;
            rk3(s, s->u, s->v, x0, y0, -1.0, &x, &y);
            clamp_pos(s, &x, &y);
            du[(size_t)j * (nx + 1) + i] = sample_u(s, s->u, x, y);
        }
    for (int j = 0; j < ny + 1; j++)
        for (int i = 0; i < nx; i++) {
            double x0 = ((double)i + 0.5) * dx, y0 = (double)j * dx, x, y;
            rk3(s, s->u, s->v, x0, y0, -1.0, &x, &y);
            clamp_pos(s, &x, &y);
            dv[(size_t)j * nx + i] = sample_v(s, s->v, x, y);
        }
    if (g_slDoubleBuffer) { memcpy(s->u, s->nu, nu * sizeof(double)); memcpy(s->v, s->nv, nv * sizeof(double)); }
}

/* src/FluidSim2D.cpp:237-250 */
static void stage_gravity(Sim* s) {
    size_t nu = (size_t)(s->nx + 1) * s->ny, nv = (size_t)s->nx * (s->ny + 1);
    for (size_t k = 0; k < nu; k++) s->u[k] += s->dt * s->gx;
    for (size_t k = 0; k < nv; k++) s->v[k] += s->dt * s->gy;
}

/* ---------------------------------------------------------------- projection */

#define C(i, j) s->cell[(size_t)(j) * nx + (i)]
#define PHI(i, j) s->phi[(size_t)(j) * nx + (i)]
#define AT(a, i, j) (a)[(size_t)(j) * nx + (i)]
#define U(a, i, j) (a)[(size_t)(j) * (nx + 1) + (i)]
#define V(a, i, j) (a)[(size_t)(j) * nx + (i)]

/* ghost-pressure matrix, src/FluidSim2D.cpp:260-303 */
static void assemble(Sim* s) {
    int nx = s->nx, ny = s->ny;
    size_t n = (size_t)nx * ny;
    memset(s->Adiag, 0, n * 8); memset(s->Ax, 0, n * 8); memset(s->Ay, 0, n * 8);
    double scaleA = s->dt / (s->rho * s->dx * s->dx);
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            if (C(i, j) != CELL_FLUID) continue;
            double d = 0.0;
            if (C(i - 1, j) == CELL_FLUID) d += scaleA;
            else if (C(i - 1, j) == CELL_EMPTY) d -= scaleA * dmax(PHI(i - 1, j) / PHI(i, j), -1e3);
            if (C(i + 1, j) == CELL_FLUID) { d += scaleA; AT(s->Ax, i, j) = -scaleA; }
            else if (C(i + 1, j) == CELL_EMPTY) d += scaleA * (1 - dmax(PHI(i + 1, j) / PHI(i, j), -1e3));
            if (C(i, j - 1) == CELL_FLUID) d += scaleA;
            else if (C(i, j - 1) == CELL_EMPTY) d -= scaleA * dmax(PHI(i, j - 1) / PHI(i, j), -1e3);
            if (C(i, j + 1) == CELL_FLUID) { d += scaleA; AT(s->Ay, i, j) = -scaleA; }
            else if (C(i, j + 1) == CELL_EMPTY) d += scaleA * (1 - dmax(PHI(i, j + 1) / PHI(i, j), -1e3));
            AT(s->Adiag, i, j) = d;
        }
}

/* negative divergence with solid-wall corrections, src/FluidSim2D.cpp:334-362 */
static void build_rhs(Sim* s) {
    int nx = s->nx, ny = s->ny;
    memset(s->rhs, 0, (size_t)nx * ny * 8);
    double scale = 1.0 / s->dx;
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            if (C(i, j) != CELL_FLUID) continue;
            double r = -scale * (U(s->u, i + 1, j) - U(s->u, i, j) + V(s->v, i, j + 1) - V(s->v, i, j));
            if (C(i - 1, j) == CELL_SOLID) r -= scale * (U(s->u, i, j) - 0);
            if (C(i + 1, j) == CELL_SOLID) r += scale * (U(s->u, i + 1, j) - 0);
            if (C(i, j - 1) == CELL_SOLID) r -= scale * (V(s->v, i, j) - 0);
            if (C(i, j + 1) == CELL_SOLID) r += scale * (V(s->v, i, j + 1) - 0);
            AT(s->rhs, i, j) = r;
        }
}

/* MIC(0) factor, tau = 0.999, sigma = 0.25, src/FluidSim2D.cpp:364-388 */
static void mic_factor(Sim* s) {
    int nx = s->nx, ny = s->ny;
    memset(s->precon, 0, (size_t)nx * ny * 8);
    const double tau = 0.999, sigma = 0.25;
    for (int j = 1; j < ny; j++)
        for (int i = 1; i < nx; i++) {
            if (C(i, j) != CELL_FLUID) continue;
            double axl = AT(s->Ax, i - 1, j), ayl = AT(s->Ay, i - 1, j), pl = AT(s->precon, i - 1, j);
            double ayd = AT(s->Ay, i, j - 1), axd = AT(s->Ax, i, j - 1), pd = AT(s->precon, i, j - 1);
            double e = AT(s->Adiag, i, j) - (axl * pl) * (axl * pl) - (ayd * pd) * (ayd * pd) -
                       tau * (axl * ayl * (pl) * (pl) + ayd * axd * (pd) * (pd));
            if (e < sigma * AT(s->Adiag, i, j)) e = AT(s->Adiag, i, j);
            AT(s->precon, i, j) = 1 / sqrt(e);
        }
}

/* z = M^-1 r by forward then backward substitution, src/FluidSim2D.cpp:397-421 */
static void mic_apply(Sim* s, const double* r, double* z) {
    int nx = s->nx, ny = s->ny;
    double* q = dalloc((size_t)nx * ny);
    for (int j = 1; j < ny; j++)
        for (int i = 1; i < nx; i++) {
            if (C(i, j) != CELL_FLUID) continue;
            double t = AT(r, i, j) - AT(s->Ax, i - 1, j) * AT(s->precon, i - 1, j) * AT(q, i - 1, j) -
                       AT(s->Ay, i, j - 1) * AT(s->precon, i, j - 1) * AT(q, i, j - 1);
            AT(q, i, j) = t * AT(s->precon, i, j);
        }
    for (int j = ny - 1; j-- > 0;)
        for (int i = nx - 1; i-- > 0;) {
            if (C(i, j) != CELL_FLUID) continue;
            double t = AT(q, i, j) - AT(s->Ax, i, j) * AT(s->precon, i, j) * AT(z, i + 1, j) -
                       AT(s->Ay, i, j) * AT(s->precon, i, j) * AT(z, i, j + 1);
            AT(z, i, j) = t * AT(s->precon, i, j);
        }
    free(q);
}

/* 5-point operator on fluid cells, src/FluidSim2D.cpp:433-444 */
static void apply_A(Sim* s, const double* x, double* y) {
    int nx = s->nx, ny = s->ny;
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            if (C(i, j) != CELL_FLUID) continue;
            AT(y, i, j) = AT(s->Adiag, i, j) * AT(x, i, j) + AT(s->Ax, i - 1, j) * AT(x, i - 1, j) +
                          AT(s->Ax, i, j) * AT(x, i + 1, j) + AT(s->Ay, i, j - 1) * AT(x, i, j - 1) +
                          AT(s->Ay, i, j) * AT(x, i, j + 1);
        }
}

/* src/FluidSim2D.cpp:252-467 */
static void stage_project(Sim* s) {
    int nx = s->nx, ny = s->ny;
    size_t n = (size_t)nx * ny;
    assemble(s);
    build_rhs(s);
    mic_factor(s);
    double *r = dalloc(n), *z = dalloc(n), *sv = dalloc(n);
    memset(s->p, 0, n * 8);
    memcpy(r, s->rhs, n * 8);
    mic_apply(s, r, z);
    memcpy(sv, z, n * 8);
    double sigma = dot(z, r, n);
    int iter = 0;
    while (iter < g_maxIters) {
        apply_A(s, sv, z);
        double rhsNorm = norm_inf(s->rhs, n);
        if (rhsNorm <= 1e-12) break;
        double alpha = sigma / dot(z, sv, n);
        axpy_to(s->p, s->p, alpha, sv, n);
        axpy_to(r, r, -alpha, z, n);
        if (norm_inf(r, n) <= g_tol * rhsNorm) break;
        mic_apply(s, r, z);
        double sigmaNew = dot(z, r, n);
        double beta = sigmaNew / sigma;
        axpy_to(sv, z, beta, sv, n);
        sigma = sigmaNew;
        iter++;
    }
    s->lastIters = iter;
    free(r); free(z); free(sv);
}

/* src/FluidSim2D.cpp:469-550 */
static void stage_update_velocity(Sim* s) {
    int nx = s->nx, ny = s->ny;
    size_t nu = (size_t)(nx + 1) * ny, nv = (size_t)nx * (ny + 1);
    uint32_t* mu = (uint32_t*)calloc(nu, sizeof(uint32_t));
    uint32_t* mv = (uint32_t*)calloc(nv, sizeof(uint32_t));
    memcpy(s->nu, s->u, nu * 8);
    memcpy(s->nv, s->v, nv * 8);
    double scale = s->dt / (s->rho * s->dx);
    const double* p = s->p;
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            if ((i > 0 && C(i - 1, j) == CELL_FLUID) || C(i, j) == CELL_FLUID) {
                if ((i == 0 || C(i - 1, j) == CELL_SOLID) || C(i, j) == CELL_SOLID) U(s->nu, i, j) = 0;
                else if (i > 0 && C(i - 1, j) == CELL_EMPTY)
                    U(s->nu, i, j) -= scale * (1 - dmax(PHI(i - 1, j) / PHI(i, j), -1e3)) * AT(p, i, j);
                else if (C(i, j) == CELL_EMPTY)
                    U(s->nu, i, j) -= scale * (dmax(PHI(i, j) / PHI(i - 1, j), -1e3) - 1) * AT(p, i - 1, j);
                else if (i > 0) U(s->nu, i, j) -= scale * (AT(p, i, j) - AT(p, i - 1, j));
                else U(s->nu, i, j) = 0;
            } else U(mu, i, j) = UINT32_MAX;
            if ((j > 0 && C(i, j - 1) == CELL_FLUID) || C(i, j) == CELL_FLUID) {
                if ((j == 0 || C(i, j - 1) == CELL_SOLID) || C(i, j) == CELL_SOLID) V(s->nv, i, j) = 0;
                else if (j > 0 && C(i, j - 1) == CELL_EMPTY)
                    V(s->nv, i, j) -= scale * (1 - dmax(PHI(i, j - 1) / PHI(i, j), -1e3)) * AT(p, i, j);
                else if (C(i, j) == CELL_EMPTY)
                    V(s->nv, i, j) -= scale * (dmax(PHI(i, j) / PHI(i, j - 1), -1e3) - 1) * AT(p, i, j - 1);
                else if (j > 0) V(s->nv, i, j) -= scale * (AT(p, i, j) - AT(p, i, j - 1));
                else V(s->nv, i, j) = 0;
            } else V(mv, i, j) = UINT32_MAX;
        }
    extrapolate(s->nu, mu, nx + 1, ny);
    extrapolate(s->nv, mv, nx, ny + 1);
    if (s->mode == 0) { memcpy(s->u, s->nu, nu * 8); memcpy(s->v, s->nv, nv * 8); }
    free(mu); free(mv);
}

/* ---------------------------------------------------------------- driver */

static void run_stage(Sim* s, int stage) {
    switch (stage) {
        case FSO_STAGE_WATER_LEVEL_SET: stage_water_level_set(s); break;
        case FSO_STAGE_P2G: stage_p2g(s); break;
        case FSO_STAGE_SL_ADVECT: stage_sl_advect(s); break;
        case FSO_STAGE_GRAVITY: stage_gravity(s); break;
        case FSO_STAGE_SOLID_LEVEL_SET: break; /* src/FluidSim2D.cpp:734-736 is an empty TODO */
        case FSO_STAGE_PROJECT: stage_project(s); break;
        case FSO_STAGE_UPDATE_VELOCITY: stage_update_velocity(s); break;
        case FSO_STAGE_G2P: stage_g2p(s); break;
        case FSO_STAGE_ADVECT: stage_advect(s); break;
        default: break;
    }
}

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* stage order of src/FluidSim2D.cpp:94-138 */
static void run_frame(Sim* s) {
    static const int sl[] = {1, 3, 4, 5, 6, 7, 9};
    static const int pf[] = {1, 2, 4, 5, 6, 7, 8, 9};
    const int* order = s->mode == 0 ? sl : pf;
    int n = s->mode == 0 ? 7 : 8;
    for (int k = 0; k < n; k++) {
        double t0 = now_ms();
        run_stage(s, order[k]);
        s->stageMs[k] = (float)(now_ms() - t0);
    }
    s->nStageMs = n;
    s->currentTime += s->dt;
}

const char* fso_kind(void) { return "port"; }

/* src/FluidSim2D.cpp:19-76: particle seeding draws from libc rand(); seed made explicit (1) */
void* fso_create(int sizeX, int sizeY, int ppcSqrt, double dt, double dx, double rho,
                 double gx, double gy, int mode, double alpha, const uint8_t* cells) {
    Sim* s = (Sim*)calloc(1, sizeof(Sim));
    s->nx = sizeX; s->ny = sizeY; s->ppcSqrt = ppcSqrt; s->mode = mode;
    s->dt = dt; s->dx = dx; s->dr = 0.9 * dx; s->rho = rho; s->gx = gx; s->gy = gy; s->alpha = alpha;
    size_t n = (size_t)sizeX * sizeY, nu = (size_t)(sizeX + 1) * sizeY, nv = (size_t)sizeX * (sizeY + 1);
    s->u = dalloc(nu); s->nu = dalloc(nu); s->v = dalloc(nv); s->nv = dalloc(nv);
    s->p = dalloc(n); s->phi = dalloc(n);
    s->Adiag = dalloc(n); s->Ax = dalloc(n); s->Ay = dalloc(n); s->rhs = dalloc(n); s->precon = dalloc(n);
    s->cell = (uint8_t*)malloc(n);
    memcpy(s->cell, cells, n);
    long fluid = 0;
    for (size_t k = 0; k < n; k++) fluid += s->cell[k] == CELL_FLUID;
    int ppc = ppcSqrt * ppcSqrt;
    s->np = fluid * ppc;
    s->pos = (double*)malloc((size_t)(s->np ? s->np : 1) * 16);
    s->vel = (double*)calloc((size_t)(s->np ? s->np : 1), 16);
    srand(1);
    float dist = 1.0f / ppcSqrt;
    long e = 0;
    for (int j = 0; j < sizeY; j++)
        for (int i = 0; i < sizeX; i++)
            if (s->cell[(size_t)j * sizeX + i] == CELL_FLUID)
                for (int k = 0; k < ppc; k++) {
                    int x = k % ppcSqrt, y = k / ppcSqrt;
                    double rx = ((double)rand() / (double)RAND_MAX) * dist;
                    s->pos[2 * e] = ((double)i + dist * x + rx) * dx;
                    double ry = ((double)rand() / (double)RAND_MAX) * dist;
                    s->pos[2 * e + 1] = ((double)j + dist * y + ry) * dx;
                    e++;
                }
    s->waterVolume = fluid * dx * dx;
    s->lastIters = -1;
    return s;
}

void fso_destroy(void* h) {
    Sim* s = (Sim*)h;
    free(s->u); free(s->v); free(s->nu); free(s->nv); free(s->p); free(s->phi); free(s->cell);
    free(s->pos); free(s->vel); free(s->Adiag); free(s->Ax); free(s->Ay); free(s->rhs); free(s->precon);
    free(s);
}

long fso_num_particles(void* h) { return ((Sim*)h)->np; }

static size_t field_ptr(Sim* s, int field, void** p) {
    size_t n = (size_t)s->nx * s->ny, nu = (size_t)(s->nx + 1) * s->ny, nv = (size_t)s->nx * (s->ny + 1);
    switch (field) {
        case FSO_U: *p = s->u; return nu * 8;
        case FSO_V: *p = s->v; return nv * 8;
        case FSO_NEWU: *p = s->nu; return nu * 8;
        case FSO_NEWV: *p = s->nv; return nv * 8;
        case FSO_P: *p = s->p; return n * 8;
        case FSO_CELL: *p = s->cell; return n;
        case FSO_PHI: *p = s->phi; return n * 8;
        case FSO_PARTICLES: *p = s->pos; return (size_t)s->np * 16;
        case FSO_PARTICLE_VELS: *p = s->vel; return (size_t)s->np * 16;
        case FSO_ADIAG: *p = s->Adiag; return n * 8;
        case FSO_AX: *p = s->Ax; return n * 8;
        case FSO_AY: *p = s->Ay; return n * 8;
        case FSO_RHS: *p = s->rhs; return n * 8;
        case FSO_PRECON: *p = s->precon; return n * 8;
        default: *p = NULL; return 0;
    }
}

int fso_get(void* h, int field, void* dst) {
    void* p; size_t n = field_ptr((Sim*)h, field, &p);
    if (!p) return -1;
    memcpy(dst, p, n);
    return 0;
}

int fso_set(void* h, int field, const void* src) {
    void* p; size_t n = field_ptr((Sim*)h, field, &p);
    if (!p) return -1;
    memcpy(p, src, n);
    return 0;
}

int fso_set_particles(void* h, long n, const double* pos, const double* vel) {
    Sim* s = (Sim*)h;
    free(s->pos); free(s->vel);
    s->np = n;
    s->pos = (double*)malloc((size_t)(n ? n : 1) * 16);
    s->vel = (double*)malloc((size_t)(n ? n : 1) * 16);
    memcpy(s->pos, pos, (size_t)n * 16);
    memcpy(s->vel, vel, (size_t)n * 16);
    return 0;
}

int fso_stage(void* h, int stage) {
    if (stage < 1 || stage > 9) return -1;
    run_stage((Sim*)h, stage);
    return 0;
}

int fso_step(void* h, int n) {
    for (int k = 0; k < n; k++) run_frame((Sim*)h);
    return 0;
}

void fso_set_params(void* h, double gx, double gy, double alpha, double dt) {
    Sim* s = (Sim*)h;
    s->gx = gx; s->gy = gy; s->alpha = alpha; s->dt = dt;
}

double fso_stat(void* h, int which) {
    Sim* s = (Sim*)h;
    switch (which) {
        case 0: return s->waterVolume;
        case 1: return s->totalEnergy;
        case 2: return s->particleTotalEnergy;
        case 3: return s->currentTime;
        default: break;
    }
    /* 4 avgPressure, 5 avgPressureInFluid, 6 maxVelocity (src/FluidSim2D.cpp:607-638): raster-order loops (FluidSim2D.h:153-159) */
    double acc = 0.0;
    size_t count = 0;
    for (int j = 0; j < s->ny; j++)
        for (int i = 0; i < s->nx; i++) {
            size_t o = (size_t)j * s->nx + i;
            if (which == 4) acc += s->p[o];
            else if (which == 5) { if (s->cell[o] == CELL_FLUID) { acc += s->p[o]; count++; } }
            else if (which == 6 && s->cell[o] == CELL_FLUID) {
                double x = (double)i * s->dx, y = (double)j * s->dx;
                double vx = sample_u(s, s->u, x, y), vy = sample_v(s, s->v, x, y);
                double vel = sqrt(vx * vx + vy * vy);
                if (vel > acc) acc = vel;
            }
        }
    if (which == 4) return acc / (s->nx * s->ny);
    if (which == 5) return acc / count;
    if (which == 6) return acc;
    return 0.0;
}

int fso_stage_times(void* h, float* out, int maxStages) {
    Sim* s = (Sim*)h;
    int n = s->nStageMs < maxStages ? s->nStageMs : maxStages;
    for (int i = 0; i < n; i++) out[i] = s->stageMs[i];
    return n;
}

/* include/MACGrid2D.h:96-98 at n positions */
int fso_vel_interp(void* h, long n, const double* pos, double* out) {
    Sim* s = (Sim*)h;
    for (long k = 0; k < n; k++) {
        out[2 * k] = sample_u(s, s->u, pos[2 * k], pos[2 * k + 1]);
        out[2 * k + 1] = sample_v(s, s->v, pos[2 * k], pos[2 * k + 1]);
    }
    return 0;
}

int fso_set_pcg(double tol, int maxIters) { g_tol = tol; g_maxIters = maxIters; return 0; }
int fso_last_pcg_iters(void* h) { return ((Sim*)h)->lastIters; }
int fso_set_sl_double_buffer(int enable) { g_slDoubleBuffer = enable; return 0; }
