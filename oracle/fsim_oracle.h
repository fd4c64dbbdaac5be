/* TEST INFRASTRUCTURE ONLY -- the product path (fluid-sim_b200/, include/fsim.h) never includes,
 * links or calls this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / reported baseline.
 *
 * fsim_oracle: a plain-C, single-threaded restatement of the per-step hot path of
 * lasagnaphil/fluid-sim @ 29962de (src/FluidSim2D.cpp, include/Array2D.h, include/MACGrid2D.h).
 * Parity is PINNED: tests/test_oracle.py checks every stage of this file against the stock
 * reference compiled from /root/reference (oracle/_ref/libfsim_ref.so) and against the golden
 * fixtures under tests/golden/ that were dumped from that build (tests/golden/make_golden.py).
 * The reference itself ships no tests or golden vectors for the solver.
 */
#ifndef FSIM_ORACLE_H
#define FSIM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* field ids shared with oracle/ref_harness.cpp and include/fsim.h */
enum {
    FSO_U = 0, FSO_V = 1, FSO_NEWU = 2, FSO_NEWV = 3, FSO_P = 4, FSO_CELL = 5, FSO_PHI = 6,
    FSO_PARTICLES = 7, FSO_PARTICLE_VELS = 8,
    /* projection internals (locals of applyProjection in the reference; exported by the port and by the patched reference build) */
    FSO_ADIAG = 9, FSO_AX = 10, FSO_AY = 11, FSO_RHS = 12, FSO_PRECON = 13
};

/* stage ids = FluidSim2D::StageType (include/FluidSim2D.h:93-97) */
enum {
    FSO_STAGE_WATER_LEVEL_SET = 1, FSO_STAGE_P2G = 2, FSO_STAGE_SL_ADVECT = 3, FSO_STAGE_GRAVITY = 4,
    FSO_STAGE_SOLID_LEVEL_SET = 5, FSO_STAGE_PROJECT = 6, FSO_STAGE_UPDATE_VELOCITY = 7,
    FSO_STAGE_G2P = 8, FSO_STAGE_ADVECT = 9
};

const char* fso_kind(void);
void* fso_create(int sizeX, int sizeY, int ppcSqrt, double dt, double dx, double rho,
                 double gx, double gy, int mode, double alpha, const uint8_t* cells);
void fso_destroy(void* h);
long fso_num_particles(void* h);
int fso_get(void* h, int field, void* dst);
int fso_set(void* h, int field, const void* src);
int fso_set_particles(void* h, long n, const double* pos, const double* vel);
int fso_stage(void* h, int stage);
int fso_step(void* h, int n);
void fso_set_params(void* h, double gx, double gy, double alpha, double dt);
double fso_stat(void* h, int which);
int fso_stage_times(void* h, float* out, int maxStages);
int fso_vel_interp(void* h, long n, const double* pos, double* out); /* MACGrid2D::velInterp at n positions */
int fso_set_pcg(double tol, int maxIters);
int fso_last_pcg_iters(void* h);
int fso_set_sl_double_buffer(int enable);

#ifdef __cplusplus
}
#endif
#endif
