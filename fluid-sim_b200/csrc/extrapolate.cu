// Array2D<double>::extrapolate (reference include/Array2D.h:552-591) without a serial queue.
//
// The reference runs a breadth-first search from the known faces: a face popped with layer number d takes
// the mean of its 4-neighbours (order x-1, x+1, y-1, y+1) whose layer is smaller.  On the full rectangle the
// BFS layer of an unknown face is its Manhattan distance to the nearest known face, so:
//   1. exact L1 distance transform (row pass with warp ballots, column pass with a running minimum),
//   2. counting sort of the unknown faces by layer,
//   3. one thread-block cluster that fills layer after layer (faces of one layer are independent; they only
//      read the previous layer), one hardware cluster barrier per layer, u and v grids in the same pass.
// Values are identical to the reference's: same neighbours, same summation order, same division.
#include "sim.h"

namespace {

constexpr int DINF = 1 << 28;

struct ExtrapArray {
    double* a;
    const uint8_t* unk;
    int* dist;
    int* distTmp;
    uint32_t* cells;
    uint8_t* cmask;   // per sorted face: which 4-neighbours lie in a smaller layer
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
    // which 4-neighbours (reference order x-1, x+1, y-1, y+1) lie in a smaller layer: stored beside the frame
    // offset so that the fill kernel does not have to read the distance field again
    const long long off = (long long)j * pitch + i;
    uint32_t mask = 0;
    if (i > 0 && X.dist[off - 1] < d) mask |= 1u;
    if (i < X.NX - 1 && X.dist[off + 1] < d) mask |= 2u;
    if (j > 0 && X.dist[off - pitch] < d) mask |= 4u;
    if (j < X.NY - 1 && X.dist[off + pitch] < d) mask |= 8u;
    X.cells[X.layerStart[d] + slot] = (uint32_t)off;
    X.cmask[X.layerStart[d] + slot] = (uint8_t)mask;
}

// Layer fill: the faces of one BFS layer are independent and only read the previous layer, so the whole fill is a
// chain of (max layer) tiny steps -- latency, not bandwidth.  One thread-block cluster of 8 CTAs x 1024 threads
// walks the layers with the hardware cluster barrier between them (release/acquire at cluster scope orders the
// global stores of layer L before the loads of layer L+1); the next layer's bounds and packed face entries are
// fetched BEFORE waiting on the barrier, so only the neighbour loads and the store sit on the per-layer path.
// (A cooperative grid.sync() per layer cost ~3.5 us; this is ~0.6 us.)
constexpr int EX_CL = 8, EX_THREADS = 1024;

__device__ __forceinline__ double fillValue(const double* a, uint32_t off, uint32_t mask, int pitch) {
    double sum = 0.0;
    int count = 0;
    if (mask & 1u) { sum += __ldcg(a + off - 1); ++count; }
    if (mask & 2u) { sum += __ldcg(a + off + 1); ++count; }
    if (mask & 4u) { sum += __ldcg(a + off - pitch); ++count; }
    if (mask & 8u) { sum += __ldcg(a + off + pitch); ++count; }
    return count == 0 ? 0.0 : sum / count;
}

__global__ void __cluster_dims__(EX_CL, 1, 1) __launch_bounds__(EX_THREADS, 1)
layerFillKernel(ExtrapArray A, ExtrapArray B, int pitch, const int* anyKnown, const int* maxLayer) {
    const int tid = blockIdx.x * EX_THREADS + threadIdx.x, nthreads = EX_CL * EX_THREADS;
    const int la = anyKnown[0] ? maxLayer[0] : 0, lb = anyKnown[1] ? maxLayer[1] : 0;
    const int lmax = max(la, lb);
    // bounds and first entries of layer 1
    int bA = 0, eA = 0, bB = 0, eB = 0;
    uint32_t pA = 0, pB = 0, mA = 0, mB = 0;
    auto fetch = [&](int L) {
        bA = eA = bB = eB = 0;
        if (L <= la) { bA = __ldcg(&A.layerStart[L]); eA = __ldcg(&A.layerStart[L + 1]); }
        if (L <= lb) { bB = __ldcg(&B.layerStart[L]); eB = __ldcg(&B.layerStart[L + 1]); }
        if (bA + tid < eA) { pA = __ldcg(&A.cells[bA + tid]); mA = __ldcg(&A.cmask[bA + tid]); }
        if (bB + tid < eB) { pB = __ldcg(&B.cells[bB + tid]); mB = __ldcg(&B.cmask[bB + tid]); }
    };
    fetch(1);
    for (int L = 1; L <= lmax; ++L) {
        const int cbA = bA, ceA = eA, cbB = bB, ceB = eB;
        const uint32_t cpA = pA, cpB = pB, cmA = mA, cmB = mB;
        if (cbA + tid < ceA) __stcg(A.a + cpA, fillValue(A.a, cpA, cmA, pitch));
        if (cbB + tid < ceB) __stcg(B.a + cpB, fillValue(B.a, cpB, cmB, pitch));
        for (int k = cbA + tid + nthreads; k < ceA; k += nthreads) {  // layers wider than the cluster (rare)
            const uint32_t p = __ldcg(&A.cells[k]);
            __stcg(A.a + p, fillValue(A.a, p, __ldcg(&A.cmask[k]), pitch));
        }
        for (int k = cbB + tid + nthreads; k < ceB; k += nthreads) {
            const uint32_t p = __ldcg(&B.cells[k]);
            __stcg(B.a + p, fillValue(B.a, p, __ldcg(&B.cmask[k]), pitch));
        }
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        if (L < lmax) fetch(L + 1);  // independent of the values: overlaps the barrier
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}

}  // namespace

int extrapolatePair(Sim* s, double* a, double* b, const uint8_t* unkA, const uint8_t* unkB) {
    const Frame& f = s->fr;
    ExtrapArray A{a, unkA, s->distU, s->distTmp, s->layerCellsU, s->layerMaskU, s->layerStartU, s->layerStartU + (s->maxLayers + 2), s->nx + 1, s->ny};
    ExtrapArray B{b, unkB, s->distV, s->distTmp + f.elems, s->layerCellsV, s->layerMaskV, s->layerStartV, s->layerStartV + (s->maxLayers + 2), s->nx, s->ny + 1};
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
    layerFillKernel<<<EX_CL, EX_THREADS, 0, s->stream>>>(A, B, f.pitch, anyKnown, maxLayer);
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}
