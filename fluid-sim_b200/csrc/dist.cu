// dist.cu -- set-up of the y-slab pressure projection (see stageApplyProjectionDist in projection.cu).
//
// One process per GPU.  The PCG iteration itself communicates through peer memory only (distpeer.cuh): fsim_dist_init
// allocates this rank's PeerBlock, exports it with CUDA IPC, exchanges the 64-byte handles and maps every other rank's
// block -- NVSwitch gives every GPU a direct store path to every peer.  NCCL is used for the plumbing around that: the
// all-gather of the IPC handles at init and the exchange of the pressure rows once per step.  It is resolved at run
// time with dlopen("libnccl.so.2"): the library then shares the copy a host application (e.g. torch) has already
// loaded, and single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <vector>

#include "distpeer.cuh"
#include "sim.h"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
    SYM(AllGather, "ncclAllGather")
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
    if (world < 1 || world > DIST_MAXW || rank < 0 || rank >= world || !uniqueId) { fsim_set_error("bad rank/world (at most %d ranks)", DIST_MAXW); return FSIM_E_INVALID; }
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
    ncclUniqueId id;
    memcpy(&id, uniqueId, sizeof(id));
    ncclComm_t comm;
    NCCL_TRY(g_nccl.CommInitRank(&comm, world, id, rank));
    d.comm = comm;
    // this rank's PeerBlock: control words (stamps all-ones = "nothing yet") + two ghost rows
    d.ghostPitch = ((s->nx + 1 + 15) / 16) * 16;
    const size_t bytes = DIST_GHOST_OFF + (size_t)2 * d.ghostPitch * sizeof(double);
    void* raw = nullptr;
    CUDA_TRY(cudaMalloc(&raw, bytes));  // (plain cudaMalloc: exportable with cudaIpcGetMemHandle)
    s->rawAllocs.push_back(raw);
    CUDA_TRY(cudaMemset(raw, 0xff, DIST_GHOST_OFF));
    CUDA_TRY(cudaMemset(static_cast<char*>(raw) + DIST_GHOST_OFF, 0, bytes - DIST_GHOST_OFF));
    d.peerLocal = raw;
    // exchange the IPC handles (64 bytes each) and map the other ranks' blocks
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t mine;
    CUDA_TRY(cudaIpcGetMemHandle(&mine, raw));
    char* dh = nullptr;
    CUDA_TRY(cudaMalloc(&dh, (size_t)64 * world));
    CUDA_TRY(cudaMemcpy(dh + (size_t)64 * rank, &mine, 64, cudaMemcpyHostToDevice));
    NCCL_TRY(g_nccl.AllGather(dh + (size_t)64 * rank, dh, 64, ncclChar, comm, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    std::vector<cudaIpcMemHandle_t> all(world);
    CUDA_TRY(cudaMemcpy(all.data(), dh, (size_t)64 * world, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaFree(dh));
    for (int r = 0; r < world; ++r) {
        if (r == rank) { d.peerBlk[r] = raw; continue; }
        void* mapped = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&mapped, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            fsim_set_error("cannot map rank %d's peer block (cudaIpcOpenMemHandle -> %s): the GPUs of a y-slab run need peer access", r, cudaGetErrorString(e));
            return FSIM_E_CUDA;
        }
        d.peerBlk[r] = mapped;
    }
    // nobody may store into a block before its owner has initialised it, nor unmap while others still store: one more
    // collective on the stream serves as the barrier
    CUDA_TRY(cudaMalloc(&dh, 64 * (size_t)world));
    NCCL_TRY(g_nccl.AllGather(dh + (size_t)64 * rank, dh, 64, ncclChar, comm, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaFree(dh));
    d.epoch = 0;
    d.on = true;
    return FSIM_OK;
}

// the device-side view of the peer blocks for this projection
int distPeerView(Sim* s, PeerView* pv) {
    const Sim::Dist& d = s->dist;
    memset(pv, 0, sizeof(*pv));
    for (int r = 0; r < d.world; ++r) pv->blk[r] = static_cast<PeerBlock*>(d.peerBlk[r]);
    pv->rank = d.rank; pv->world = d.world; pv->ghostPitch = d.ghostPitch;
    pv->stampBase = d.epoch << 16;
    return FSIM_OK;
}

void distDestroy(Sim* s) {
    Sim::Dist& d = s->dist;
    for (int r = 0; r < d.world; ++r)
        if (d.on && r != d.rank && d.peerBlk[r]) cudaIpcCloseMemHandle(d.peerBlk[r]);
    if (d.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(d.comm));
    d.comm = nullptr;
    d.on = false;
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
