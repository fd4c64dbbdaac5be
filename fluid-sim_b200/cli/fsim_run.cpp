// fsim_run -- headless driver of the B200 hot path.
//
// The reference can only be run through its SDL2/OpenGL demo (demo/App.cpp:133-213: fill FluidSim2DConfig, create(),
// update() once per rendered frame, saveStats() at exit).  This is the same loop without a window, over the C ABI of
// include/fsim.h: it builds the demo's dam-break scene (demo/App.cpp:147-160, row-major as SURVEY.md D12 notes), steps it,
// and writes the two files FluidSim2D::saveStats produces, in the reference's formats:
//   perf.csv          PerformanceCounter::saveToFile (src/PerformanceCounter.cpp:58-74): one row per frame after the first
//                     30, each the 30-frame rolling mean of every stage's time in ms, std::to_string floats, ',' after each
//   conservation.csv  src/FluidSim2D.cpp:738-750: header line, then "%f, %f, %f\n" of volume, grid energy, particle energy
// Stage times are CUDA-event times of the device stages.  --save / --load write and read FSIMCKP1 checkpoints
// (fsim_checkpoint_save / fsim_checkpoint_load).
//
//   fsim_run [--size N] [--size-y M] [--mode flip|sl] [--steps K] [--alpha A] [--dt DT] [--dx DX] [--rho R] [--ppc S]
//            [--gravity GX GY] [--tol T] [--max-iters I] [--device D] [--sl-snapshot] [--out DIR] [--load FILE]
//            [--save FILE] [--save-every K] [--quiet]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fsim.h"

namespace {

// rolling means like the reference's PerformanceCounter (src/PerformanceCounter.cpp:16-56): a ring of 30 frames per stage;
// once the ring has filled, every frame appends the per-stage means to the store (zeros before that)
struct StageMeans {
    static const int kRing = 30;  // PerformanceCounter::SampleCount (src/PerformanceCounter.h)
    int stages = 0, frame = 0;
    bool filled = false;
    std::vector<std::vector<float>> ring;
    std::vector<float> mean, store;
    void init(int n) { stages = n; ring.assign(n, std::vector<float>(kRing, 0.f)); mean.assign(n, 0.f); }
    void push(const float* ms) {
        for (int i = 0; i < stages; ++i) ring[i][frame] = ms[i];
        if (filled)
            for (int i = 0; i < stages; ++i) {
                float a = 0.f;
                for (int k = 0; k < kRing; ++k) a += ring[i][k];
                mean[i] = a / kRing;
            }
        if (frame == kRing - 1) filled = true;
        frame = (frame + 1) % kRing;
        for (int i = 0; i < stages; ++i) store.push_back(mean[i]);
    }
    bool save(const std::string& path) const {
        FILE* f = fopen(path.c_str(), "w+");
        if (!f) return false;
        const size_t rows = stages ? store.size() / stages : 0;
        for (size_t r = kRing; r < rows; ++r) {
            for (int j = 0; j < stages; ++j) fprintf(f, "%s,", std::to_string(store[r * stages + j]).c_str());
            fputc('\n', f);
        }
        fclose(f);
        return true;
    }
};

[[noreturn]] void die(const char* what) {
    fprintf(stderr, "fsim_run: %s: %s\n", what, fsim_last_error());
    exit(EXIT_FAILURE);
}
#define CHECK(call) do { if ((call) != FSIM_OK) die(#call); } while (0)

}  // namespace

int main(int argc, char** argv) {
    int nx = 128, ny = 0, steps = 100, ppc = 2, saveEvery = 0;
    bool flip = true, quiet = false, dtSet = false, dxSet = false;
    double alpha = 0.05, dt = 0.005, dx = 0.01, rho = 997.0, gx = 0.0, gy = -9.81;
    std::string outDir = ".", loadPath, savePath;
    fsim_options opt;
    fsim_default_options(&opt);
    for (int a = 1; a < argc; ++a) {
        const std::string k = argv[a];
        auto need = [&](int n) { if (a + n >= argc) { fprintf(stderr, "fsim_run: %s needs %d value(s)\n", k.c_str(), n); exit(2); } };
        if (k == "--size") { need(1); nx = atoi(argv[++a]); }
        else if (k == "--size-y") { need(1); ny = atoi(argv[++a]); }
        else if (k == "--mode") { need(1); flip = std::string(argv[++a]) != "sl"; }
        else if (k == "--steps") { need(1); steps = atoi(argv[++a]); }
        else if (k == "--alpha") { need(1); alpha = atof(argv[++a]); }
        else if (k == "--dt") { need(1); dt = atof(argv[++a]); dtSet = true; }
        else if (k == "--dx") { need(1); dx = atof(argv[++a]); dxSet = true; }
        else if (k == "--rho") { need(1); rho = atof(argv[++a]); }
        else if (k == "--ppc") { need(1); ppc = atoi(argv[++a]); }
        else if (k == "--gravity") { need(2); gx = atof(argv[++a]); gy = atof(argv[++a]); }
        else if (k == "--tol") { need(1); opt.pcgTol = atof(argv[++a]); }
        else if (k == "--max-iters") { need(1); opt.pcgMaxIters = atoi(argv[++a]); }
        else if (k == "--device") { need(1); opt.device = atoi(argv[++a]); }
        else if (k == "--sl-snapshot") opt.slDoubleBuffer = 1;
        else if (k == "--out") { need(1); outDir = argv[++a]; }
        else if (k == "--load") { need(1); loadPath = argv[++a]; }
        else if (k == "--save") { need(1); savePath = argv[++a]; }
        else if (k == "--save-every") { need(1); saveEvery = atoi(argv[++a]); }
        else if (k == "--quiet") quiet = true;
        else { fprintf(stderr, "fsim_run: unknown option %s\n", k.c_str()); return 2; }
    }
    if (ny <= 0) ny = nx;
    // the demo's numbers at 128 cells (demo/App.cpp:136-146); larger grids keep the domain (1.28 m) and the CFL number
    if (!dxSet) dx = 1.28 / nx;
    if (!dtSet) dt = nx > 128 ? 0.005 * 128.0 / nx : 0.005;

    fsim_handle h = nullptr;
    if (!loadPath.empty()) {
        CHECK(fsim_checkpoint_load(loadPath.c_str(), &opt, &h));
    } else {
        std::vector<uint8_t> cells((size_t)nx * ny);
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                uint8_t c = (i + j < ny * 3 / 4) ? 1 : 0;  // FLUID below the diagonal, EMPTY above
                if (i == 0 || j == 0 || i == nx - 1 || j == ny - 1) c = 2;  // SOLID border
                cells[(size_t)j * nx + i] = c;
            }
        fsim_config cfg;
        cfg.sizeX = nx; cfg.sizeY = ny; cfg.particlesPerCellSqrt = ppc;
        cfg.dt = dt; cfg.dx = dx; cfg.rho = rho; cfg.gravityX = gx; cfg.gravityY = gy;
        cfg.mode = flip ? FSIM_PICFLIP : FSIM_SEMILAGRANGIAN;
        cfg.picFlipAlpha = alpha;
        cfg.initialValues = cells.data();
        CHECK(fsim_create(&cfg, &opt, &h));
    }
    size_t np = 0;
    CHECK(fsim_num_particles(h, &np));
    if (!quiet) fprintf(stderr, "fsim_run: %s, %zu particles, %d steps (%s)\n", fsim_version(), np, steps, loadPath.empty() ? "new scene" : "resumed");

    StageMeans perf;
    std::vector<double> vol, eGrid, ePart;
    double totalMs = 0.0;
    for (int k = 0; k < steps; ++k) {
        double ms = 0.0;
        CHECK(fsim_step_timed(h, 1, &ms));
        totalMs += ms;
        fsim_stats st;
        CHECK(fsim_get_stats(h, &st));
        if (k == 0) perf.init(st.numStages);
        perf.push(st.stageMs);
        vol.push_back(st.waterVolume); eGrid.push_back(st.totalEnergy); ePart.push_back(st.particleTotalEnergy);
        if (st.pcgHitMaxIters && !quiet) fprintf(stderr, "Maximum iteration limit exceeded!\n");          // src/FluidSim2D.cpp:464-466
        if (st.cflMax > 5.0) fprintf(stderr, "CFL condition broken: %f > 5\n", st.cflMax);                // :582-585
        if (st.nanPositions) { fprintf(stderr, "Error: particle position is NaN\n"); return EXIT_FAILURE; }  // :598-601
        if (saveEvery > 0 && !savePath.empty() && (k + 1) % saveEvery == 0) CHECK(fsim_checkpoint_save(h, savePath.c_str()));
    }
    if (!savePath.empty()) CHECK(fsim_checkpoint_save(h, savePath.c_str()));

    if (!perf.save(outDir + "/perf.csv")) { fprintf(stderr, "fsim_run: cannot write %s/perf.csv\n", outDir.c_str()); return EXIT_FAILURE; }
    FILE* f = fopen((outDir + "/conservation.csv").c_str(), "w+");
    if (!f) { fprintf(stderr, "fsim_run: cannot write %s/conservation.csv\n", outDir.c_str()); return EXIT_FAILURE; }
    fputs("Total Volume, Total Energy, Total Energy (Particle) \n", f);
    for (size_t k = 0; k < vol.size(); ++k) fprintf(f, "%f, %f, %f\n", vol[k], eGrid[k], ePart[k]);
    fclose(f);

    double avgP = 0, avgPF = 0, maxV = 0;
    CHECK(fsim_diagnostics(h, &avgP, &avgPF, &maxV));
    if (!quiet && steps > 0)
        printf("{\"steps\": %d, \"ms_per_step\": %.4f, \"mcell_steps_per_s\": %.4f, \"avg_pressure\": %.9g, \"avg_pressure_in_fluid\": %.9g, "
               "\"max_velocity\": %.9g}\n", steps, totalMs / steps, (double)nx * ny / (totalMs / steps) / 1e3, avgP, avgPF, maxV);
    CHECK(fsim_destroy(h));
    return 0;
}
