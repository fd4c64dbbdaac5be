// FluidSim2D_b200.cpp -- drop-in replacement for the reference's src/FluidSim2D.cpp.
//
// Compile this file INSTEAD of src/FluidSim2D.cpp inside the reference tree (it includes the reference's own
// include/FluidSim2D.h, so the struct layout, field names and method signatures callers see are unchanged) and
// link libfsim_b200.so.  Every method below forwards to the C ABI of include/fsim.h; the heavy state lives in
// HBM and the public host arrays (mac, newMac, p, cell, particles, particleVels, waterLevelSet.phi) are mirrors:
//
//   update()/runFrame()   upload mac.u / mac.v (the demo's renderer writes them, demo/FluidRenderer2D.cpp:305-308),
//                         push gravity / picFlipAlpha / dt (re-read every frame like the reference does), run one
//                         device step, download every public field (renderer reads them, :436-485)
//   stage methods         upload the whole host state, run that one stage on the device, download it again
//                         (slow, but each public stage method keeps its stand-alone meaning)
//
// The mirrored arrays are page-locked in place (fsim_host_register) so that the per-frame copies run at full PCIe rate
// beside the device stages.
//
// Environment: FSIM_B200_NO_MIRROR=1 skips the per-frame downloads (headless runs; call fsimShimSync() before
// reading fields), FSIM_B200_NO_PIN=1 leaves the host arrays pageable, FSIM_B200_DEVICE selects the CUDA device.  There is no CPU fallback: if the library cannot
// create a device simulation the process exits like the reference does on allocation failure (Array2D.h:69-72).
//
// The handle cannot be stored in the struct (its layout belongs to the reference and it is copied by value), so
// it is kept in a side table keyed by the address of the pressure array, which is unique per instance and
// travels with every copy of the struct.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "FluidSim2D.h"   // the reference's header (include/FluidSim2D.h)
#include "fsim.h"         // this repository's include/fsim.h

namespace {

struct ShimState {
    fsim_handle h = nullptr;
    bool mirror = true;
    std::vector<void*> pinned;  // host arrays page-locked by pinMirrors
    void* pinnedParticles = nullptr;  // particles.data at the time of pinning
};

std::unordered_map<const void*, ShimState>& table() {
    static std::unordered_map<const void*, ShimState> t;
    return t;
}

ShimState& stateOf(const FluidSim2D* sim) {
    auto it = table().find(sim->p.data);
    if (it == table().end()) {
        fprintf(stderr, "FluidSim2D (b200): this object was not made by FluidSim2D::create\n");
        exit(EXIT_FAILURE);
    }
    return it->second;
}

void check(int rc, const char* what) {
    if (rc != FSIM_OK) {
        fprintf(stderr, "FluidSim2D (b200): %s failed: %s\n", what, fsim_last_error());
        exit(EXIT_FAILURE);
    }
}

size_t bytesOf(const Array2D<double>& a) { return sizeof(double) * (size_t)a.NX * a.NY; }

void unpinMirrors(ShimState& st) {
    for (void* q : st.pinned) fsim_host_unregister(st.h, q);
    st.pinned.clear();
    st.pinnedParticles = nullptr;
}

// page-locks the arrays fsim_step_host mirrors every frame; a refusal (locked-memory limit) only costs the overlap
void pinMirrors(FluidSim2D* sim, ShimState& st) {
    const char* np = getenv("FSIM_B200_NO_PIN");
    if (!st.mirror || (np && np[0] == '1')) return;
    struct { void* ptr; size_t bytes; } bufs[] = {
        {sim->mac.u.data, bytesOf(sim->mac.u)}, {sim->mac.v.data, bytesOf(sim->mac.v)}, {sim->p.data, bytesOf(sim->p)},
        {sim->waterLevelSet.phi.data, bytesOf(sim->waterLevelSet.phi)},
        {sim->cell.data, (size_t)sim->cell.NX * sim->cell.NY * sizeof(*sim->cell.data)},
        {sim->particles.data, sim->particles.size * sizeof(vec2d)},
        {sim->particleVels.data, sim->particleVels.size * sizeof(vec2d)}};
    for (auto& b : bufs) {
        if (!b.ptr || !b.bytes) continue;
        if (fsim_host_register(st.h, b.ptr, b.bytes) != FSIM_OK) {
            fprintf(stderr, "FluidSim2D (b200): host arrays stay pageable: %s\n", fsim_last_error());
            unpinMirrors(st);
            return;
        }
        st.pinned.push_back(b.ptr);
    }
    st.pinnedParticles = sim->particles.data;
}

void resizeParticles(FluidSim2D* sim, size_t n) {
    if (sim->particles.capacity < n || sim->particleVels.capacity < n) {
        auto it = table().find(sim->p.data);
        if (it != table().end()) unpinMirrors(it->second);  // reserve() moves the arrays
        if (sim->particles.capacity < n) sim->particles.reserve(n);
        if (sim->particleVels.capacity < n) sim->particleVels.reserve(n);
    }
    sim->particles.size = n;
    sim->particleVels.size = n;
}

void downloadAll(FluidSim2D* sim, ShimState& st) {
    fsim_handle h = st.h;
    check(fsim_download(h, FSIM_U, sim->mac.u.data, bytesOf(sim->mac.u)), "download u");
    check(fsim_download(h, FSIM_V, sim->mac.v.data, bytesOf(sim->mac.v)), "download v");
    check(fsim_download(h, FSIM_NEWU, sim->newMac.u.data, bytesOf(sim->newMac.u)), "download newMac.u");
    check(fsim_download(h, FSIM_NEWV, sim->newMac.v.data, bytesOf(sim->newMac.v)), "download newMac.v");
    check(fsim_download(h, FSIM_P, sim->p.data, bytesOf(sim->p)), "download p");
    check(fsim_download(h, FSIM_PHI, sim->waterLevelSet.phi.data, bytesOf(sim->waterLevelSet.phi)), "download phi");
    check(fsim_download(h, FSIM_CELL, sim->cell.data, (size_t)sim->cell.NX * sim->cell.NY), "download cell");
    size_t n = 0;
    check(fsim_num_particles(h, &n), "particle count");
    resizeParticles(sim, n);
    if (n) {
        check(fsim_download(h, FSIM_PARTICLES, sim->particles.data, n * sizeof(vec2d)), "download particles");
        check(fsim_download(h, FSIM_PARTICLE_VELS, sim->particleVels.data, n * sizeof(vec2d)), "download particleVels");
    }
}

void uploadAll(FluidSim2D* sim, ShimState& st) {
    fsim_handle h = st.h;
    check(fsim_upload(h, FSIM_U, sim->mac.u.data, bytesOf(sim->mac.u)), "upload u");
    check(fsim_upload(h, FSIM_V, sim->mac.v.data, bytesOf(sim->mac.v)), "upload v");
    check(fsim_upload(h, FSIM_NEWU, sim->newMac.u.data, bytesOf(sim->newMac.u)), "upload newMac.u");
    check(fsim_upload(h, FSIM_NEWV, sim->newMac.v.data, bytesOf(sim->newMac.v)), "upload newMac.v");
    check(fsim_upload(h, FSIM_P, sim->p.data, bytesOf(sim->p)), "upload p");
    check(fsim_upload(h, FSIM_PHI, sim->waterLevelSet.phi.data, bytesOf(sim->waterLevelSet.phi)), "upload phi");
    check(fsim_upload(h, FSIM_CELL, sim->cell.data, (size_t)sim->cell.NX * sim->cell.NY), "upload cell");
    check(fsim_set_particles(h, sim->particles.size, reinterpret_cast<const double*>(sim->particles.data),
                             reinterpret_cast<const double*>(sim->particleVels.data)), "upload particles");
}

void pushParams(FluidSim2D* sim, ShimState& st) {
    check(fsim_set_params(st.h, sim->gravity.x, sim->gravity.y, sim->picFlipAlpha, sim->dt), "set_params");
}

void pullStats(FluidSim2D* sim, ShimState& st, bool record) {
    fsim_stats s;
    check(fsim_get_stats(st.h, &s), "get_stats");
    sim->waterVolume = s.waterVolume;
    sim->totalEnergy = s.totalEnergy;
    sim->particleTotalEnergy = s.particleTotalEnergy;
    if (record) {
        sim->waterVolumeData.push(s.waterVolume);
        sim->totalEnergyData.push(s.totalEnergy);
        sim->particleTotalEnergyData.push(s.particleTotalEnergy);
    }
    if (s.pcgHitMaxIters) fprintf(stderr, "Maximum iteration limit exceeded!\n");  // src/FluidSim2D.cpp:464-466
    if (s.cflMax > 5.0) fprintf(stderr, "CFL condition broken: %f > 5\n", s.cflMax);  // :582-585
    if (s.nanPositions) {  // :598-601
        fprintf(stderr, "Error: particle position is NaN (%d particles)\n", s.nanPositions);
        exit(1);
    }
}

void runStage(FluidSim2D* sim, int stage, FluidSim2D::StageType tag) {
    ShimState& st = stateOf(sim);
    sim->stage = tag;
    uploadAll(sim, st);
    pushParams(sim, st);
    check(fsim_stage(st.h, stage), "stage");
    downloadAll(sim, st);
    if (stage == FSIM_STAGE_CREATE_WATER_LEVEL_SET) pullStats(sim, st, true);
}

}  // namespace

// Explicit device -> host refresh for callers that run with FSIM_B200_NO_MIRROR=1.
extern "C" void fsimShimSync(FluidSim2D* sim) {
    ShimState& st = stateOf(sim);
    downloadAll(sim, st);
}
extern "C" fsim_handle fsimShimHandle(FluidSim2D* sim) { return stateOf(sim).h; }

FluidSim2D FluidSim2D::create(const FluidSim2DConfig& config) {
    FluidSim2D sim;
    sim.sizeX = config.sizeX;
    sim.sizeY = config.sizeY;
    sim.particlesPerCellSqrt = config.particlesPerCellSqrt;
    sim.particlesPerCell = config.particlesPerCellSqrt * config.particlesPerCellSqrt;
    sim.dt = config.dt;
    sim.dx = config.dx;
    sim.dr = 0.9 * config.dx;
    sim.rho = config.rho;
    sim.gravity = vec2d{config.gravityX, config.gravityY};
    sim.origGravity = sim.gravity;
    sim.mode = config.mode;
    sim.picFlipAlpha = config.picFlipAlpha;
    sim.numStages = config.mode == FS_SEMILAGRANGIAN ? 7 : 8;
    sim.mac = MACGrid2D::create(config.sizeX, config.sizeY, config.dx);
    sim.newMac = MACGrid2D::create(config.sizeX, config.sizeY, config.dx);
    sim.p = Array2D<double>::create(config.sizeX, config.sizeY);
    sim.cell = Array2D<FluidCellType>::create(config.sizeX, config.sizeY);
    sim.waterLevelSet = LevelSet::create(config.sizeX, config.sizeY, config.dx);
    sim.perfCounter = PerformanceCounter::create(sim.numStages);

    fsim_config cfg;
    cfg.sizeX = config.sizeX; cfg.sizeY = config.sizeY; cfg.particlesPerCellSqrt = config.particlesPerCellSqrt;
    cfg.dt = config.dt; cfg.dx = config.dx; cfg.rho = config.rho;
    cfg.gravityX = config.gravityX; cfg.gravityY = config.gravityY;
    cfg.mode = config.mode == FS_SEMILAGRANGIAN ? FSIM_SEMILAGRANGIAN : FSIM_PICFLIP;
    cfg.picFlipAlpha = config.picFlipAlpha;
    cfg.initialValues = reinterpret_cast<const uint8_t*>(config.initialValues);  // borrowed, one byte per cell
    fsim_options opt;
    fsim_default_options(&opt);
    if (const char* d = getenv("FSIM_B200_DEVICE")) opt.device = atoi(d);
    ShimState st;
    const char* nm = getenv("FSIM_B200_NO_MIRROR");
    st.mirror = !(nm && nm[0] == '1');
    check(fsim_create(&cfg, &opt, &st.h), "fsim_create");
    table()[sim.p.data] = st;

    // the device seeded the particles exactly like the reference (glibc rand(), seed 1); mirror the start state
    sim.particles = Vec<vec2d>::create(0);
    sim.particleVels = Vec<vec2d>::create(0);
    downloadAll(&sim, table()[sim.p.data]);
    pinMirrors(&sim, table()[sim.p.data]);
    size_t fluidCells = 0;
    for (size_t k = 0; k < (size_t)sim.sizeX * sim.sizeY; ++k) fluidCells += sim.cell.data[k] == FS_FLUID;
    sim.origWaterVolume = sim.waterVolume = (double)fluidCells * sim.dx * sim.dx;
    return sim;
}

void FluidSim2D::free() {
    auto it = table().find(p.data);
    if (it != table().end()) {
        unpinMirrors(it->second);
        fsim_destroy(it->second.h);
        table().erase(it);
    }
    mac.free();
    newMac.free();
    p.free();
    cell.free();
    particles.free();
    particleVels.free();
    waterLevelSet.free();
    perfCounter.free();
    waterVolumeData.free();
    totalEnergyData.free();
    particleTotalEnergyData.free();
}

void FluidSim2D::runFrame() {
    ShimState& st = stateOf(this);
    pushParams(this, st);
    // Mirror contract: mac.u / mac.v go up every frame; a caller that replaced or resized the particle
    // Vecs has them re-uploaded (and re-pinned: a page lock on the old allocation would dangle); every public field comes down.
    {
        size_t devN = 0;
        check(fsim_num_particles(st.h, &devN), "particle count");
        const bool moved = !st.pinned.empty() && st.pinnedParticles != static_cast<void*>(particles.data);
        if (devN != particles.size || moved) {
            unpinMirrors(st);
            if (particleVels.size != particles.size) { fprintf(stderr, "FluidSim2D (b200): particles / particleVels sizes differ\n"); exit(EXIT_FAILURE); }
            check(fsim_set_particles(st.h, particles.size, reinterpret_cast<const double*>(particles.data),
                                     reinterpret_cast<const double*>(particleVels.data)), "upload particles");
            pinMirrors(this, st);
        }
    }
    fsim_host_mirror io;
    memset(&io, 0, sizeof(io));
    io.u_in = mac.u.data;  // callers may have written the velocity field since the last frame
    io.v_in = mac.v.data;
    if (st.mirror) {
        io.u = mac.u.data; io.v = mac.v.data; io.p = p.data;
        io.cell = reinterpret_cast<uint8_t*>(cell.data);
        io.phi = waterLevelSet.phi.data;
        io.particles = reinterpret_cast<double*>(particles.data);
        io.particleVels = reinterpret_cast<double*>(particleVels.data);
    }
    check(fsim_step_host(st.h, &io), "fsim_step_host");
    if (st.mirror) {  // the reference ends every frame with mac == newMac (:547-549, :566)
        memcpy(newMac.u.data, mac.u.data, bytesOf(mac.u));
        memcpy(newMac.v.data, mac.v.data, bytesOf(mac.v));
    }
    pullStats(this, st, true);
    // per-stage device times (CUDA events) go where the reference keeps its chrono samples
    fsim_stats s;
    check(fsim_get_stats(st.h, &s), "get_stats");
    for (int k = 0; k < s.numStages && k < numStages; ++k)
        perfCounter.samples[k][perfCounter.currentFrame] = s.stageMs[k];  // milliseconds, like PerformanceCounter::endStage
    perfCounter.currentStage = numStages;
    perfCounter.endFrame();
    stage = StageType::ApplyAdvection;
    rendered = false;
    currentTime += dt;
}

void FluidSim2D::update() { runFrame(); }

void FluidSim2D::createWaterLevelSet() { runStage(this, FSIM_STAGE_CREATE_WATER_LEVEL_SET, StageType::CreateWaterLevelSet); }
void FluidSim2D::transferVelocityToGrid() { runStage(this, FSIM_STAGE_TRANSFER_VELOCITY_TO_GRID, StageType::TransferVelocityToGrid); }
void FluidSim2D::applySemiLagrangianAdvection() { runStage(this, FSIM_STAGE_APPLY_SEMI_LAGRANGIAN_ADVECTION, StageType::ApplySemiLagrangianAdvection); }
void FluidSim2D::applyGravity() { runStage(this, FSIM_STAGE_APPLY_GRAVITY, StageType::ApplyGravity); }
void FluidSim2D::createSolidLevelSet() { stage = StageType::CreateSolidLevelSet; }  // empty in the reference (:734-736)
void FluidSim2D::applyProjection() { runStage(this, FSIM_STAGE_APPLY_PROJECTION, StageType::ApplyProjection); }
void FluidSim2D::updateVelocity() { runStage(this, FSIM_STAGE_UPDATE_VELOCITY, StageType::UpdateVelocity); }
void FluidSim2D::updateParticleVelocities() { runStage(this, FSIM_STAGE_UPDATE_PARTICLE_VELOCITIES, StageType::UpdateParticleVelocities); }
void FluidSim2D::applyAdvection() { runStage(this, FSIM_STAGE_APPLY_ADVECTION, StageType::ApplyAdvection); }

// Queries of the reference (:607-638) as one fused device reduction over the resident state (fsim_diagnostics): they
// stay valid with FSIM_B200_NO_MIRROR=1.  With mirrors on, the host's mac.u / mac.v are pushed first -- the renderer may
// have edited them since the last frame (demo/FluidRenderer2D.cpp:305-308) and maxVelocity samples them.
namespace {
void deviceDiagnostics(FluidSim2D* sim, double* avgP, double* avgPFluid, double* maxVel) {
    ShimState& st = stateOf(sim);
    if (st.mirror) {
        check(fsim_upload(st.h, FSIM_U, sim->mac.u.data, bytesOf(sim->mac.u)), "upload u");
        check(fsim_upload(st.h, FSIM_V, sim->mac.v.data, bytesOf(sim->mac.v)), "upload v");
    }
    check(fsim_diagnostics(st.h, avgP, avgPFluid, maxVel), "fsim_diagnostics");
}
}  // namespace

double FluidSim2D::avgPressure() {
    double v = 0.0;
    deviceDiagnostics(this, &v, nullptr, nullptr);
    return v;
}

double FluidSim2D::avgPressureInFluid() {
    double v = 0.0;
    deviceDiagnostics(this, nullptr, &v, nullptr);
    return v;
}

double FluidSim2D::maxVelocity() {
    double v = 0.0;
    deviceDiagnostics(this, nullptr, nullptr, &v);
    return v;
}

vec2d FluidSim2D::getGridCenter() { return vec2d{sizeX * dx / 2, sizeY * dx / 2}; }

vec2d FluidSim2D::clampPos(vec2d to) {
    const double lo = (1.0 + 1e-3) * dx;
    return vec2d{aml::clamp<double>(to.x, lo, (sizeX - 1.0 - 1e-3) * dx), aml::clamp<double>(to.y, lo, (sizeY - 1.0 - 1e-3) * dx)};
}

// perf.csv through the reference's PerformanceCounter, conservation.csv in the reference's format (:738-750)
void FluidSim2D::saveStats() {
    perfCounter.saveToFile("perf.csv");
    FILE* f = fopen("conservation.csv", "w+");
    if (!f) { fprintf(stderr, "cannot open conservation.csv\n"); exit(EXIT_FAILURE); }
    fputs("Total Volume, Total Energy, Total Energy (Particle) \n", f);
    for (size_t k = 0; k < waterVolumeData.size; ++k)
        fprintf(f, "%f, %f, %f\n", waterVolumeData[k], totalEnergyData[k], particleTotalEnergyData[k]);
    fclose(f);
}

// LevelSet's two public methods are sub-steps of createWaterLevelSet on the device and are not offered
// separately; callers in the reference only reach them through FluidSim2D::createWaterLevelSet (:653-656).
void LevelSet::constructFromParticles(Vec<vec2d>, double) {
    fprintf(stderr, "LevelSet::constructFromParticles: use FluidSim2D::createWaterLevelSet with the b200 library\n");
    exit(EXIT_FAILURE);
}
void LevelSet::redistance() {
    fprintf(stderr, "LevelSet::redistance: use FluidSim2D::createWaterLevelSet with the b200 library\n");
    exit(EXIT_FAILURE);
}
