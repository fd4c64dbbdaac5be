// dist.cu -- the NCCL side of the y-slab pressure projection (see stageApplyProjectionDist in projection.cu).
//
// One process per GPU.  NCCL is resolved at run time with dlopen("libnccl.so.2"): the library then shares the copy
// a host application (e.g. torch) has already loaded, and single-GPU users need no NCCL at all.  Only point-to-point
// halo rows (ncclSend/ncclRecv, NVLink P2P under the hood), 8-byte allreduces of the PCG scalars and the final
// exchange of pressure rows go through it; everything is enqueued on the simulation's own stream.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "sim.h"

int distPackHalo(Sim* s, int unpack);

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int loadNccl() {
    if (g_nccl.lib) return FSIM_OK;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { fsim_set_error("cannot load libnccl.so.2: %s", dlerror()); return FSIM_E_STATE; }
#define SYM(field, name)                                                                         \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(lib, name);                                 \
    if (!g_nccl.field) { fsim_set_error("libnccl: missing symbol %s", name); return FSIM_E_STATE; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(Broadcast, "ncclBroadcast")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return FSIM_OK;
}

#define NCCL_TRY(expr)                                                                                   \
    do {                                                                                                 \
        ncclResult_t _r = (expr);                                                                        \
        if (_r != ncclSuccess) {                                                                         \
            fsim_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r));     \
            return FSIM_E_CUDA;                                                                          \
        }                                                                                                \
    } while (0)

}  // namespace

int distGetUniqueId(void* out128) {
    int rc = loadNccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return FSIM_OK;
}

// own strips of rank r when `ns` strips are dealt to `world` ranks in contiguous blocks as even as possible
void distSlabOf(int ns, int world, int r, int* strip0, int* nOwn) {
    const int base = ns / world, extra = ns % world;
    *strip0 = r * base + (r < extra ? r : extra);
    *nOwn = base + (r < extra ? 1 : 0);
}
static void slabOf(int ns, int world, int r, int* strip0, int* nOwn) { distSlabOf(ns, world, r, strip0, nOwn); }

int distInit(Sim* s, int rank, int world, const void* uniqueId) {
    if (world < 1 || rank < 0 || rank >= world || !uniqueId) { fsim_set_error("bad rank/world"); return FSIM_E_INVALID; }
    if (s->dist.on) { fsim_set_error("fsim_dist_init called twice"); return FSIM_E_STATE; }
    const int SR = 32 * s->sdg.rpl;
    const int ns = (s->ny + SR - 1) / SR;
    if (world > ns) { fsim_set_error("more ranks (%d) than %d-row strips (%d)", world, SR, ns); return FSIM_E_INVALID; }
    Sim::Dist& d = s->dist;
    d.rank = rank; d.world = world;
    // (the slabs themselves are chosen every step from the fluid cells' bounding box, see stageApplyProjectionDist)
    int rc = loadNccl();
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(s->device));
    void* raw = nullptr;
    CUDA_TRY(cudaMalloc(&raw, (size_t)4 * s->nx * sizeof(double)));
    s->rawAllocs.push_back(raw);
    CUDA_TRY(cudaMemset(raw, 0, (size_t)4 * s->nx * sizeof(double)));
    d.haloSend = reinterpret_cast<double*>(raw);
    d.haloRecv = d.haloSend + 2 * (size_t)s->nx;
    ncclUniqueId id;
    memcpy(&id, uniqueId, sizeof(id));
    ncclComm_t comm;
    NCCL_TRY(g_nccl.CommInitRank(&comm, world, id, rank));
    d.comm = comm;
    d.on = true;
    return FSIM_OK;
}

void distDestroy(Sim* s) {
    if (s->dist.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(s->dist.comm));
    s->dist.comm = nullptr;
    s->dist.on = false;
}

// in-place allreduce of one device double (sum or max)
int distAllReduce(Sim* s, double* devPtr, int isMax) {
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(s->dist.comm);
    NCCL_TRY(g_nccl.AllReduce(devPtr, devPtr, 1, ncclDouble, isMax ? ncclMax : ncclSum, comm, s->stream));
    LAUNCH_COUNT(s);
    return FSIM_OK;
}

// first / last own row of the search direction to the neighbours' halo strips
int distHaloExchange(Sim* s) {
    const Sim::Dist& d = s->dist;
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(d.comm);
    int rc = distPackHalo(s, 0);
    if (rc) return rc;
    const size_t n = (size_t)d.gExt.nx;
    NCCL_TRY(g_nccl.GroupStart());
    if (d.rank > 0) {
        NCCL_TRY(g_nccl.Send(d.haloSend, n, ncclDouble, d.rank - 1, comm, s->stream));
        NCCL_TRY(g_nccl.Recv(d.haloRecv, n, ncclDouble, d.rank - 1, comm, s->stream));
    }
    if (d.rank < d.world - 1) {
        NCCL_TRY(g_nccl.Send(d.haloSend + n, n, ncclDouble, d.rank + 1, comm, s->stream));
        NCCL_TRY(g_nccl.Recv(d.haloRecv + n, n, ncclDouble, d.rank + 1, comm, s->stream));
    }
    NCCL_TRY(g_nccl.GroupEnd());
    LAUNCH_COUNT(s);
    return distPackHalo(s, 1);
}

// every rank's rows of a frame-shaped array to every rank (in place)
int distShareRows(Sim* s, double* frame) {
    const Sim::Dist& d = s->dist;
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(d.comm);
    NCCL_TRY(g_nccl.GroupStart());
    for (int r = 0; r < d.world; ++r) {
        int st0, n;
        slabOf(d.boxStrips, d.world, r, &st0, &n);
        const int SR = 32 * s->sdg.rpl;
        const int j0 = SR * (st0 + d.boxStrip0);
        const int j1 = j0 + SR * n < s->ny ? j0 + SR * n : s->ny;
        if (j1 <= j0) continue;
        double* p = frame + (long long)j0 * s->fr.pitch;
        NCCL_TRY(g_nccl.Broadcast(p, p, (size_t)(j1 - j0) * s->fr.pitch, ncclDouble, r, comm, s->stream));
    }
    NCCL_TRY(g_nccl.GroupEnd());
    LAUNCH_COUNT(s);
    return FSIM_OK;
}
