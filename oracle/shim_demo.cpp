// TEST INFRASTRUCTURE ONLY.  Headless driver for the drop-in shim (fluid-sim_b200/shim/FluidSim2D_b200.cpp): the
// same calls demo/App.cpp makes (FluidSim2D::create, update() per frame, saveStats-free teardown), compiled against
// the reference's own headers, linked with libfsim_b200.so.  Dumps the public fields after N frames so that
// tests/test_gpu_shim.py can compare them with the stock reference (oracle/_ref/libfsim_ref.so) run beside it.
// usage: shim_demo N mode(0 SL / 1 PICFLIP) steps out.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "FluidSim2D.h"

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: shim_demo N mode steps out.bin\n"); return 2; }
    int n = atoi(argv[1]), mode = atoi(argv[2]), steps = atoi(argv[3]);
    std::vector<FluidCellType> cells((size_t)n * n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            FluidCellType c = (i + j < n * 3 / 4) ? FS_FLUID : FS_EMPTY;   // scene of demo/App.cpp:147-160, row-major
            if (i == 0 || j == 0 || i == n - 1 || j == n - 1) c = FS_SOLID;
            cells[(size_t)j * n + i] = c;
        }
    FluidSim2DConfig config;
    config.sizeX = n; config.sizeY = n; config.particlesPerCellSqrt = 2;
    config.dt = 0.005; config.dx = 1.28 / n; config.rho = 997.0; config.gravityX = 0.0; config.gravityY = -9.81;
    config.mode = mode ? FS_PICFLIP : FS_SEMILAGRANGIAN; config.picFlipAlpha = 0.05; config.initialValues = cells.data();
    FluidSim2D sim = FluidSim2D::create(config);
    for (int k = 0; k < steps; ++k) sim.update();
    FILE* f = fopen(argv[4], "wb");
    if (!f) return 3;
    long long np = (long long)sim.particles.size;
    fwrite(&np, 8, 1, f);
    fwrite(sim.mac.u.data, 8, (size_t)(n + 1) * n, f);
    fwrite(sim.mac.v.data, 8, (size_t)n * (n + 1), f);
    fwrite(sim.p.data, 8, (size_t)n * n, f);
    fwrite(sim.waterLevelSet.phi.data, 8, (size_t)n * n, f);
    fwrite(sim.cell.data, 1, (size_t)n * n, f);
    fwrite(sim.particles.data, 16, (size_t)np, f);
    fwrite(sim.particleVels.data, 16, (size_t)np, f);
    fclose(f);
    printf("shim_demo: %d steps, %lld particles, volume %.6f, avgP(fluid) %.6f, time %.4f\n", steps, np, sim.waterVolume,
           sim.avgPressureInFluid(), sim.currentTime);
    sim.free();
    return 0;
}
