// State checkpoint files: everything FluidSim2D carries from one update() to the next -- the public
// data members of reference include/FluidSim2D.h:71-78 (mac, newMac, p, cell, waterLevelSet.phi, particles, particleVels)
// plus the scalars update() re-reads (dt, gravity, picFlipAlpha, currentTime) -- as one little-endian binary file:
//
//   char magic[8] = "FSIMCKP1"; int32 sizeX, sizeY, particlesPerCellSqrt, mode; double dt, dx, rho, gravityX, gravityY,
//   picFlipAlpha, currentTime; uint64 numParticles; then the dense arrays in field-id order (include/fsim.h):
//   u[(sizeX+1)*sizeY] v[sizeX*(sizeY+1)] newU newV p[sizeX*sizeY] (doubles) cell[sizeX*sizeY] (bytes) phi (doubles)
//   particles[2*np] particleVels[2*np] (doubles).
//
// The reference has no checkpointing; a resumed run continues bit for bit (tests/test_gpu_cli.py).
#include <stdio.h>
#include <string.h>

#include <vector>

#include "sim.h"

namespace {

const char MAGIC[8] = {'F', 'S', 'I', 'M', 'C', 'K', 'P', '1'};

struct Header {
    char magic[8];
    int32_t sizeX, sizeY, ppcSqrt, mode;
    double dt, dx, rho, gx, gy, alpha, currentTime;
    uint64_t np;
};

size_t fieldBytes(int field, int nx, int ny, size_t np) {
    switch (field) {
        case FSIM_U: case FSIM_NEWU: return (size_t)(nx + 1) * ny * 8;
        case FSIM_V: case FSIM_NEWV: return (size_t)nx * (ny + 1) * 8;
        case FSIM_CELL: return (size_t)nx * ny;
        case FSIM_PARTICLES: case FSIM_PARTICLE_VELS: return np * 16;
        default: return (size_t)nx * ny * 8;
    }
}

}  // namespace

extern "C" int fsim_checkpoint_save(fsim_handle h, const char* path) {
    if (!h || !path) { fsim_set_error("null argument"); return FSIM_E_INVALID; }
    Sim* s = reinterpret_cast<Sim*>(h);
    FILE* f = fopen(path, "wb");
    if (!f) { fsim_set_error("cannot open %s for writing", path); return FSIM_E_STATE; }
    Header hd;
    memset(&hd, 0, sizeof(hd));
    memcpy(hd.magic, MAGIC, 8);
    hd.sizeX = s->nx; hd.sizeY = s->ny; hd.ppcSqrt = s->ppcSqrt; hd.mode = s->mode;
    hd.dt = s->dt; hd.dx = s->dx; hd.rho = s->rho; hd.gx = s->gx; hd.gy = s->gy; hd.alpha = s->alpha;
    hd.currentTime = s->currentTime; hd.np = s->np;
    int rc = FSIM_OK;
    if (fwrite(&hd, sizeof(hd), 1, f) != 1) rc = FSIM_E_STATE;
    std::vector<unsigned char> buf;
    for (int field = FSIM_U; field <= FSIM_PARTICLE_VELS && rc == FSIM_OK; ++field) {
        const size_t bytes = fieldBytes(field, s->nx, s->ny, s->np);
        if (!bytes) continue;
        buf.resize(bytes);
        if ((rc = fsim_download(h, field, buf.data(), bytes))) break;
        if (fwrite(buf.data(), 1, bytes, f) != bytes) rc = FSIM_E_STATE;
    }
    if (fclose(f) != 0 && rc == FSIM_OK) rc = FSIM_E_STATE;
    if (rc == FSIM_E_STATE) fsim_set_error("short write to %s", path);
    return rc;
}

extern "C" int fsim_checkpoint_load(const char* path, const fsim_options* optIn, fsim_handle* out) {
    if (!path || !out) { fsim_set_error("null argument"); return FSIM_E_INVALID; }
    FILE* f = fopen(path, "rb");
    if (!f) { fsim_set_error("cannot open %s", path); return FSIM_E_STATE; }
    Header hd;
    if (fread(&hd, sizeof(hd), 1, f) != 1 || memcmp(hd.magic, MAGIC, 8) != 0) {
        fclose(f);
        fsim_set_error("%s is not an FSIMCKP1 checkpoint", path);
        return FSIM_E_INVALID;
    }
    if (hd.sizeX < 4 || hd.sizeY < 4 || hd.sizeX > 32768 || hd.sizeY > 32768) {
        fclose(f);
        fsim_set_error("%s: bad grid size", path);
        return FSIM_E_INVALID;
    }
    const int nx = hd.sizeX, ny = hd.sizeY;
    std::vector<std::vector<unsigned char>> data(FSIM_PARTICLE_VELS + 1);
    for (int field = FSIM_U; field <= FSIM_PARTICLE_VELS; ++field) {
        const size_t bytes = fieldBytes(field, nx, ny, hd.np);
        data[field].resize(bytes);
        if (bytes && fread(data[field].data(), 1, bytes, f) != bytes) {
            fclose(f);
            fsim_set_error("%s is truncated", path);
            return FSIM_E_INVALID;
        }
    }
    fclose(f);
    fsim_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.sizeX = nx; cfg.sizeY = ny; cfg.particlesPerCellSqrt = hd.ppcSqrt; cfg.mode = hd.mode;
    cfg.dt = hd.dt; cfg.dx = hd.dx; cfg.rho = hd.rho; cfg.gravityX = hd.gx; cfg.gravityY = hd.gy; cfg.picFlipAlpha = hd.alpha;
    // the labels of a running simulation keep the SOLID cells of the initial ones (relabelling only touches non-SOLID cells,
    // src/FluidSim2D.cpp:657-664), so they pass fsim_create's border check
    cfg.initialValues = data[FSIM_CELL].data();
    fsim_options opt;
    if (optIn) opt = *optIn; else fsim_default_options(&opt);
    opt.seedParticles = 0;
    fsim_handle h = nullptr;
    int rc = fsim_create(&cfg, &opt, &h);
    if (rc) return rc;
    rc = fsim_set_particles(h, (size_t)hd.np, reinterpret_cast<const double*>(data[FSIM_PARTICLES].data()),
                            reinterpret_cast<const double*>(data[FSIM_PARTICLE_VELS].data()));
    for (int field = FSIM_U; field <= FSIM_PHI && rc == FSIM_OK; ++field)
        rc = fsim_upload(h, field, data[field].data(), data[field].size());
    if (rc) { fsim_destroy(h); return rc; }
    reinterpret_cast<Sim*>(h)->currentTime = hd.currentTime;
    *out = h;
    return FSIM_OK;
}
