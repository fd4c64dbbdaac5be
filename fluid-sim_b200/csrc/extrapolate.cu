// Array2D<double>::extrapolate (reference include/Array2D.h:552-591) without a serial queue.
//
// The reference runs a breadth-first search from the known faces: a face popped with layer number d takes
// the mean of its 4-neighbours (order x-1, x+1, y-1, y+1) whose layer is smaller.  On the full rectangle the
// BFS layer of an unknown face is its Manhattan distance to the nearest known face, so:
//   1. exact L1 distance transform (row pass with warp ballots, column pass with a running minimum),
//   2. counting sort of the unknown faces by layer,
//   3. one persistent cooperative kernel that fills layer after layer (faces of one layer are independent;
//      they only read the previous layer), one grid barrier per layer, u and v grids in the same pass.
// Values are identical to the reference's: same neighbours, same summation order, same division.
#include <cooperative_groups.h>

#include "sim.h"

namespace cg = cooperative_groups;

namespace {

constexpr int DINF = 1 << 28;

struct ExtrapArray {
    double* a;
    const uint8_t* unk;
    int* dist;
    int* distTmp;
    uint32_t* cells;
    int* layerStart;  // [maxLayers+2]; doubles as the histogram before the scan
    int* layerCursor;
    int NX, NY;
};

// row pass: distance to the nearest known face in the same row; one warp per row
__global__ void distRowKernel(ExtrapArray A, ExtrapArray B, int pitch, const int* anyKnown) {
    int which = blockIdx.y;
    if (anyKnown[which] == 0) return;
    const ExtrapArray& X = which == 0 ? A : B;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= X.NY) return;
    const uint8_t* unk = X.unk + (long long)warp * pitch;
    int* d = X.dist + (long long)warp * pitch;
    int last = -DINF;
    for (int x0 = 0; x0 < X.NX; x0 += 32) {
        int idx = x0 + lane;
        bool known = idx < X.NX && unk[idx] == 0;
        unsigned m = __ballot_sync(0xffffffffu, known);
        unsigned mine = m & (0xffffffffu >> (31 - lane));
        int nearest = mine ? x0 + 31 - __clz(mine) : last;
        if (idx < X.NX) d[idx] = nearest > -DINF ? idx - nearest : DINF;
        if (m) last = x0 + 31 - __clz(m);
    }
    int next = DINF;
    for (int x0 = ((X.NX - 1) / 32) * 32; x0 >= 0; x0 -= 32) {
        int idx = x0 + lane;
        bool known = idx < X.NX && unk[idx] == 0;
        unsigned m = __ballot_sync(0xffffffffu, known);
        unsigned mine = m & (0xffffffffu << lane);
        int nearest = mine ? x0 + __ffs(mine) - 1 : next;
        if (idx < X.NX) {
            int dr = nearest < DINF ? nearest - idx : DINF;
            d[idx] = min(d[idx], dr);
        }
        if (m) next = x0 + __ffs(m) - 1;
    }
}

// column pass: d(i,j) = min_j' d_row(i,j') + |j-j'|, as an ascending then a descending running minimum.
// One thread per column (coalesced across the warp); in/out are distinct arrays so the loads pipeline.
template <bool FINAL>
__global__ void distColKernel(ExtrapArray A, ExtrapArray B, int pitch, const int* anyKnown) {
    int which = blockIdx.y;
    if (anyKnown[which] == 0) return;
    const ExtrapArray& X = which == 0 ? A : B;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= X.NX) return;
    const int* __restrict__ in = FINAL ? X.distTmp : X.dist;
    int* __restrict__ out = FINAL ? X.dist : X.distTmp;
    int run = DINF;
    if (!FINAL) {
        for (int j = 0; j < X.NY; ++j) {
            int v = in[(long long)j * pitch + i];
            run = min(v, run + 1);
            out[(long long)j * pitch + i] = run;
        }
    } else {
        for (int j = X.NY - 1; j >= 0; --j) {
            int v = in[(long long)j * pitch + i];
            run = min(v, run + 1);
            out[(long long)j * pitch + i] = run;
            if (run > 0 && run < DINF) atomicAdd(&X.layerStart[run], 1);  // histogram of layers
        }
    }
}

// exclusive scan of the layer histogram (single block), max layer
__global__ void layerScanKernel(ExtrapArray A, ExtrapArray B, int maxLayers, const int* anyKnown, int* maxLayerOut) {
    int which = blockIdx.x;
    const ExtrapArray& X = which == 0 ? A : B;
    if (anyKnown[which] == 0) { if (threadIdx.x == 0) maxLayerOut[which] = 0; return; }
    __shared__ int carry, top;
    __shared__ int buf[1024];
    if (threadIdx.x == 0) { carry = 0; top = 0; }
    __syncthreads();
    for (int base = 0; base <= maxLayers + 1; base += 1024) {
        int idx = base + threadIdx.x;
        int v = idx <= maxLayers + 1 ? X.layerStart[idx] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        // Hillis-Steele inclusive scan
        for (int o = 1; o < 1024; o <<= 1) {
            int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        int incl = buf[threadIdx.x];
        if (idx <= maxLayers + 1) {
            X.layerStart[idx] = carry + incl - v;
            X.layerCursor[idx] = 0;
        }
        if (v > 0) atomicMax(&top, idx);
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) maxLayerOut[which] = top;
}

__global__ void layerScatterKernel(ExtrapArray A, ExtrapArray B, int pitch, const int* anyKnown) {
    int which = blockIdx.z;
    if (anyKnown[which] == 0) return;
    const ExtrapArray& X = which == 0 ? A : B;
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= X.NX || j >= X.NY) return;
    int d = X.dist[(long long)j * pitch + i];
    if (d <= 0 || d >= DINF) return;
    int slot = atomicAdd(&X.layerCursor[d], 1);
    X.cells[X.layerStart[d] + slot] = (uint32_t)(j * pitch + i);
}

__device__ __forceinline__ void fillLayer(const ExtrapArray& X, int L, int pitch, int tid, int nthreads) {
    int b = X.layerStart[L], e = X.layerStart[L + 1];
    for (int k = b + tid; k < e; k += nthreads) {
        int off = (int)X.cells[k];
        int y = off / pitch, x = off - y * pitch;
        double sum = 0.0;
        int count = 0;
        // neighbours in the reference's order; a neighbour contributes iff its layer is smaller
        if (x > 0 && X.dist[off - 1] < L) { sum += __ldcg(X.a + off - 1); ++count; }
        if (x < X.NX - 1 && X.dist[off + 1] < L) { sum += __ldcg(X.a + off + 1); ++count; }
        if (y > 0 && X.dist[off - pitch] < L) { sum += __ldcg(X.a + off - pitch); ++count; }
        if (y < X.NY - 1 && X.dist[off + pitch] < L) { sum += __ldcg(X.a + off + pitch); ++count; }
        __stcg(X.a + off, count == 0 ? 0.0 : sum / count);
    }
}

__global__ void layerFillKernel(ExtrapArray A, ExtrapArray B, int pitch, const int* anyKnown, const int* maxLayer) {
    cg::grid_group grid = cg::this_grid();
    int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    int la = anyKnown[0] ? maxLayer[0] : 0, lb = anyKnown[1] ? maxLayer[1] : 0;
    int lmax = max(la, lb);
    for (int L = 1; L <= lmax; ++L) {
        if (L <= la) fillLayer(A, L, pitch, tid, nthreads);
        if (L <= lb) fillLayer(B, L, pitch, tid, nthreads);
        grid.sync();
    }
}

}  // namespace

int extrapolatePair(Sim* s, double* a, double* b, const uint8_t* unkA, const uint8_t* unkB) {
    const Frame& f = s->fr;
    ExtrapArray A{a, unkA, s->distU, s->distTmp, s->layerCellsU, s->layerStartU, s->layerStartU + (s->maxLayers + 2), s->nx + 1, s->ny};
    ExtrapArray B{b, unkB, s->distV, s->distTmp + f.elems, s->layerCellsV, s->layerStartV, s->layerStartV + (s->maxLayers + 2), s->nx, s->ny + 1};
    A.dist += f.org; A.distTmp += f.org; B.dist += f.org; B.distTmp += f.org;
    const int* anyKnown = s->ctl->anyKnown;
    int* maxLayer = s->ctl->maxLayer;
    size_t histBytes = (size_t)(s->maxLayers + 2) * 2 * sizeof(int);
    CUDA_TRY(cudaMemsetAsync(s->layerStartU, 0, histBytes, s->stream));
    CUDA_TRY(cudaMemsetAsync(s->layerStartV, 0, histBytes, s->stream));
    int rowsMax = s->ny + 1, colsMax = s->nx + 1;
    distRowKernel<<<dim3((rowsMax * 32 + 255) / 256, 2), 256, 0, s->stream>>>(A, B, f.pitch, anyKnown);
    distColKernel<false><<<dim3((colsMax + 127) / 128, 2), 128, 0, s->stream>>>(A, B, f.pitch, anyKnown);
    distColKernel<true><<<dim3((colsMax + 127) / 128, 2), 128, 0, s->stream>>>(A, B, f.pitch, anyKnown);
    layerScanKernel<<<2, 1024, 0, s->stream>>>(A, B, s->maxLayers, anyKnown, maxLayer);
    layerScatterKernel<<<dim3((colsMax + 31) / 32, (rowsMax + 7) / 8, 2), dim3(32, 8), 0, s->stream>>>(A, B, f.pitch, anyKnown);
    s->launches += 5;
    CUDA_TRY(cudaGetLastError());
    int pitch = f.pitch;
    void* args[] = {&A, &B, &pitch, (void*)&anyKnown, (void*)&maxLayer};
    int grid = 64;
    CUDA_TRY(cudaLaunchCooperativeKernel((void*)layerFillKernel, dim3(grid), dim3(256), args, 0, s->stream));
    LAUNCH_COUNT(s);
    return FSIM_OK;
}
