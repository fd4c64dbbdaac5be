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
#include "sdwave.cuh"
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
    unsigned long long* cons;  // per sorted face: (CTA, slot) of the next-layer faces that read it, 16 bits per direction
    int* layerStart;  // [maxLayers+2]; doubles as the histogram before the scan
    int* layerCursor;
    int* pos;         // frame-shaped: sorted index of every unknown face (aliases distTmp, free after the column pass)
    int NX, NY;
    double* a2;       // optional second destination of the fill (same frame layout as a)
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
// The recurrence is sequential along a column, so a block takes 32 columns and walks them in chunks of 64 rows: all
// 8 warps fetch the 64x32 chunk with coalesced row segments (the DRAM latency is paid once per chunk, 2048 loads in
// flight), warp 0 runs the 32 column recurrences through shared memory, all warps write the chunk back.  (One thread
// per column reading global memory row by row took 1.3-1.8 ms at 4096^2: one DRAM round trip per row.)  The final
// pass also builds the histogram of layers, in shared memory first (one global atomic per non-empty bin and block).
constexpr int DC_ROWS = 64;
template <bool FINAL>
__global__ void __launch_bounds__(256) distColKernel(ExtrapArray A, ExtrapArray B, int pitch, const int* anyKnown, int nbins) {
    extern __shared__ int dcSmem[];
    int* tile = dcSmem;                 // [DC_ROWS][32]
    int* hist = dcSmem + DC_ROWS * 32;  // [nbins] (FINAL only)
    int which = blockIdx.y;
    if (anyKnown[which] == 0) return;
    const ExtrapArray& X = which == 0 ? A : B;
    const int i0 = blockIdx.x * 32;
    if (i0 >= X.NX) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = i0 + lane;
    const bool col = i < X.NX;
    const int* __restrict__ in = FINAL ? X.distTmp : X.dist;
    int* __restrict__ out = FINAL ? X.dist : X.distTmp;
    if (FINAL) {
        for (int k = threadIdx.x; k < nbins; k += 256) hist[k] = 0;
    }
    int run = DINF;
    const int nchunks = (X.NY + DC_ROWS - 1) / DC_ROWS;
    for (int cidx = 0; cidx < nchunks; ++cidx) {
        const int c = FINAL ? nchunks - 1 - cidx : cidx;
        const int j0 = c * DC_ROWS;
        __syncthreads();
        for (int r = w; r < DC_ROWS; r += 8) {
            const int j = j0 + r;
            tile[r * 32 + lane] = (col && j < X.NY) ? in[(long long)j * pitch + i] : DINF;
        }
        __syncthreads();
        if (w == 0) {
            if (!FINAL) {
#pragma unroll 8
                for (int r = 0; r < DC_ROWS; ++r) { run = min(tile[r * 32 + lane], run + 1); tile[r * 32 + lane] = run; }
            } else {
#pragma unroll 8
                for (int r = DC_ROWS - 1; r >= 0; --r) { run = min(tile[r * 32 + lane], run + 1); tile[r * 32 + lane] = run; }
            }
        }
        __syncthreads();
        for (int r = w; r < DC_ROWS; r += 8) {
            const int j = j0 + r;
            if (col && j < X.NY) {
                const int v = tile[r * 32 + lane];
                out[(long long)j * pitch + i] = v;
                if (FINAL && v > 0 && v < DINF) atomicAdd(&hist[v], 1);  // histogram of layers
            }
        }
    }
    if (FINAL) {
        __syncthreads();
        for (int k = threadIdx.x; k < nbins; k += 256)
            if (hist[k]) atomicAdd(&X.layerStart[k], hist[k]);
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
    // low nibble: the neighbour contributes; high nibble: it is itself an unknown face (of layer d-1, filled by the
    // same kernel) rather than a known one
    uint32_t mask = 0;
    int dn;
    if (i > 0 && (dn = X.dist[off - 1]) < d) mask |= dn > 0 ? 0x11u : 0x01u;
    if (i < X.NX - 1 && (dn = X.dist[off + 1]) < d) mask |= dn > 0 ? 0x22u : 0x02u;
    if (j > 0 && (dn = X.dist[off - pitch]) < d) mask |= dn > 0 ? 0x44u : 0x04u;
    if (j < X.NY - 1 && (dn = X.dist[off + pitch]) < d) mask |= dn > 0 ? 0x88u : 0x08u;
    X.cells[X.layerStart[d] + slot] = (uint32_t)off;
    X.cmask[X.layerStart[d] + slot] = (uint8_t)mask;
    X.pos[off] = X.layerStart[d] + slot;
}

// Layer fill: the faces of one BFS layer are independent and only read the previous layer, so the whole fill is a
// chain of (max layer) small steps -- latency, not bandwidth.  One thread-block cluster of 16 CTAs x 512 threads
// walks the layers; face q of a layer belongs to thread q mod 8192.  A value is stored to global memory (the
// result) and pushed, with plain remote shared-memory stores, straight into the private slots of the faces of
// the next layer that read it: every thread owns four slots per array (one per neighbour direction), and the
// (CTA, slot) targets of every face are precomputed.  Slots are self-validating (a reserved NaN payload = "not
// written yet"): a reader polls its own shared memory and clears the slot.  Between layers there is only the
// relaxed hardware cluster barrier (flow control: nobody runs two layers ahead; 54 ns measured) -- no memory fence
// and no global-memory round trip (a release/acquire cluster barrier behind a global store costs 0.8 us, a
// cooperative grid.sync() 3.5 us; tools/clbar.cu).  What does not depend on the previous layer (bounds, face
// entries, values of known neighbours) is fetched ahead through a software pipeline of L2 prefetches.  A layer with
// more faces than threads is read from global memory behind a fenced barrier instead.
constexpr int EX_CL = 16, EX_THREADS = 512, EX_NT = EX_CL * EX_THREADS;
constexpr int EX_SLOTS = 2 * 2 * EX_THREADS * 4;  // [parity][array][thread][direction]
constexpr unsigned long long EX_SENT = 0x7FF8F51D0DEAD002ULL;

// for every unknown face: where its value has to be pushed -- per direction n the neighbour of the next layer (if
// any) is face q of that layer; it lives in CTA (q / 512) % 16, thread q % 512, and reads us from its slot of the
// opposite direction.  16 bits per direction: cta << 11 | thread * 4 + direction, 0xFFFF = nobody.
__global__ void layerConsumersKernel(ExtrapArray A, ExtrapArray B, int pitch, const int* anyKnown) {
    int which = blockIdx.z;
    if (anyKnown[which] == 0) return;
    const ExtrapArray& X = which == 0 ? A : B;
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= X.NX || j >= X.NY) return;
    const long long off = (long long)j * pitch + i;
    const int d = X.dist[off];
    if (d <= 0 || d >= DINF) return;
    const int nextStart = X.layerStart[d + 1];
    const bool nextInSlots = X.layerStart[d + 2] - nextStart <= EX_NT;
    unsigned long long t = 0;
    auto look = [&](bool inb, long long o, int n) {
        unsigned long long tg = 0xFFFFull;
        if (inb && nextInSlots && X.dist[o] == d + 1) {
            const int q = X.pos[o] - nextStart;
            tg = (unsigned long long)((((q / EX_THREADS) % EX_CL) << 11) | ((q % EX_THREADS) * 4 + (n ^ 1)));
        }
        t |= tg << (16 * n);
    };
    look(i > 0, off - 1, 0);
    look(i < X.NX - 1, off + 1, 1);
    look(j > 0, off - pitch, 2);
    look(j < X.NY - 1, off + pitch, 3);
    X.cons[X.pos[off]] = t;
}

struct FacePre {   // one face of the layer being prepared
    uint32_t off, mask;
    unsigned long long cons;
    double kv[4];  // values of known neighbours
};

// nearPtr / part: the fill can be cut in two launches at layer K = *nearPtr -- part 0 fills the layers 1..K, part 1 the
// layers K+1.. (its first layer reads layer K from global memory, where part 0 left it).  nearPtr == nullptr: everything.
__global__ void __launch_bounds__(EX_THREADS, 1)
layerFillKernel(ExtrapArray A, ExtrapArray B, int pitch, const int* anyKnown, const int* maxLayer, const int* nearPtr, int part) {
    extern __shared__ __align__(16) unsigned char exSmem[];
    unsigned long long* slots = reinterpret_cast<unsigned long long*>(exSmem);  // [2][2][EX_THREADS][4]
    unsigned int rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tid = (int)rank * EX_THREADS + threadIdx.x;
    const int la = anyKnown[0] ? maxLayer[0] : 0, lb = anyKnown[1] ? maxLayer[1] : 0;
    const int lmax = max(la, lb);
    int L0 = 1, L1 = lmax;
    if (nearPtr) {
        const int K = max(*nearPtr, 1);
        if (part == 0) L1 = min(lmax, K); else L0 = K + 1;
    }
    if (L0 > L1) return;  // (uniform over the cluster)
    for (int i = threadIdx.x; i < EX_SLOTS; i += EX_THREADS) slots[i] = EX_SENT;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const unsigned int slotsA = sd::smemAddr(slots);
    auto bounds = [&](int L, int& bA, int& eA, int& bB, int& eB) {
        bA = eA = bB = eB = 0;
        // (read-only during the kernel and the same for all 8192 threads: through L1, not 1024 L2 requests to one line)
        if (L >= 1 && L <= la) { bA = __ldg(&A.layerStart[L]); eA = __ldg(&A.layerStart[L + 1]); }
        if (L >= 1 && L <= lb) { bB = __ldg(&B.layerStart[L]); eB = __ldg(&B.layerStart[L + 1]); }
    };
    // Software pipeline over the layers (a cold load is a ~1 us DRAM round trip, a layer takes ~0.3 us):
    //   layer L+10: its face entries are prefetched into L2
    //   layer L+5:  this thread's face entry is loaded (an L2 hit by now) ...
    //   layer L+4:  ... and, one layer later, the lines of its known neighbours' values are prefetched
    //   layer L+1:  those values are loaded (L2 hits) between the barrier's arrive and wait
    constexpr int PD = 5, PF = 10;
    struct Bnd { int bA, eA, bB, eB; };
    Bnd bnd[PF + 1];  // bounds of layers L .. L+PF
#pragma unroll
    for (int d = 0; d <= PF; ++d) bounds(L0 + d, bnd[d].bA, bnd[d].eA, bnd[d].bB, bnd[d].eB);
    struct Entry { uint32_t off, mask; unsigned long long cons; };
    const int nb[4] = {-1, 1, -pitch, pitch};
    auto loadEntry = [&](const ExtrapArray& X, int k, Entry& e) {
        e.off = __ldcg(&X.cells[k]); e.mask = __ldcg(&X.cmask[k]); e.cons = __ldcg(&X.cons[k]);
    };
    auto prefetchL2 = [&](const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); };
    auto prefetchKnown = [&](const ExtrapArray& X, const Entry& e) {
#pragma unroll
        for (int n = 0; n < 4; ++n)
            if ((e.mask & (0x11u << n)) == (1u << n)) prefetchL2(X.a + (long long)e.off + nb[n]);
    };
    auto finishFace = [&](const ExtrapArray& X, const Entry& e, FacePre& f) {
        f.off = e.off; f.mask = e.mask; f.cons = e.cons;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            f.kv[n] = 0.0;
            if ((f.mask & (0x11u << n)) == (1u << n)) f.kv[n] = __ldcg(X.a + (long long)f.off + nb[n]);
        }
    };
    // fromSlots: the unknown neighbours' values arrive in this thread's slots (else: global memory)
    auto faceValue = [&](const ExtrapArray& X, const FacePre& f, bool fromSlots, unsigned int mySlots) -> double {
        double sum = 0.0;
        int count = 0;
#pragma unroll
        for (int n = 0; n < 4; ++n) {  // the reference's neighbour order
            if (f.mask & (1u << n)) {
                double v = f.kv[n];
                if (f.mask & (0x10u << n)) {
                    if (fromSlots) {
                        unsigned long long bits;
                        const unsigned int addr = mySlots + n * 8;
                        do { asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(bits) : "r"(addr) : "memory"); } while (bits == EX_SENT);
                        asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(addr), "l"(EX_SENT) : "memory");
                        v = __longlong_as_double((long long)bits);
                    } else {
                        v = __ldcg(X.a + (long long)f.off + nb[n]);
                    }
                }
                sum += v;
                ++count;
            }
        }
        return count == 0 ? 0.0 : sum / count;
    };
    // result to global memory, and into the slots of the next layer's readers (parity `par`, array `which`)
    auto emit = [&](const ExtrapArray& X, int which, int par, const FacePre& f, double v) {
        __stcg(X.a + f.off, v);
        if (X.a2) __stcg(X.a2 + f.off, v);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const unsigned int tg = (unsigned int)(f.cons >> (16 * n)) & 0xFFFFu;
            if (tg != 0xFFFFu) {
                unsigned int remote;
                const unsigned int local = slotsA + (unsigned int)((((par * 2 + which) * EX_THREADS * 4) + (tg & 0x7FFu)) * 8);
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(tg >> 11));
                asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
            }
        }
    };
    // cold path for the faces of a layer beyond the first EX_NT (such a layer is read from global memory)
    auto coldFace = [&](const ExtrapArray& X, int which, int par, int k) {
        Entry e;
        loadEntry(X, k, e);
        FacePre f;
        finishFace(X, e, f);
        emit(X, which, par, f, faceValue(X, f, false, 0u));
    };
    FacePre fa, fb;
    fa.mask = 0; fb.mask = 0;
    {
        Entry e;
        if (bnd[0].bA + tid < bnd[0].eA) { loadEntry(A, bnd[0].bA + tid, e); finishFace(A, e, fa); }
        if (bnd[0].bB + tid < bnd[0].eB) { loadEntry(B, bnd[0].bB + tid, e); finishFace(B, e, fb); }
    }
    Entry qA[PD + 1], qB[PD + 1];  // this thread's face entries of layers L+1 .. L+PD (index = distance)
#pragma unroll
    for (int d = 1; d <= PD; ++d) {
        qA[d] = {0, 0, 0}; qB[d] = {0, 0, 0};
        if (d < PD) {
            if (bnd[d].bA + tid < bnd[d].eA) loadEntry(A, bnd[d].bA + tid, qA[d]);
            if (bnd[d].bB + tid < bnd[d].eB) loadEntry(B, bnd[d].bB + tid, qB[d]);
        }
    }
    for (int L = L0; L <= L1; ++L) {
        const int par = L & 1;
        const int bA = bnd[0].bA, eA = bnd[0].eA, bB = bnd[0].bB, eB = bnd[0].eB;
        // this layer reads the previous one from its slots iff it fits the cluster in one pass (the pushes were
        // planned with the same rule, per array)
        const bool slotsInA = L > L0 && (eA - bA) <= EX_NT, slotsInB = L > L0 && (eB - bB) <= EX_NT;
        const bool nextFits = (bnd[1].eA - bnd[1].bA) <= EX_NT && (bnd[1].eB - bnd[1].bB) <= EX_NT;
        // pipeline stages that do not depend on anything computed here
        qA[PD] = {0, 0, 0}; qB[PD] = {0, 0, 0};
        if (bnd[PD].bA + tid < bnd[PD].eA) loadEntry(A, bnd[PD].bA + tid, qA[PD]);
        if (bnd[PD].bB + tid < bnd[PD].eB) loadEntry(B, bnd[PD].bB + tid, qB[PD]);
        prefetchKnown(A, qA[PD - 1]);  // (loaded one layer ago)
        prefetchKnown(B, qB[PD - 1]);
        if (bnd[PF].bA + tid < bnd[PF].eA) { prefetchL2(&A.cells[bnd[PF].bA + tid]); prefetchL2(&A.cmask[bnd[PF].bA + tid]); prefetchL2(&A.cons[bnd[PF].bA + tid]); }
        if (bnd[PF].bB + tid < bnd[PF].eB) { prefetchL2(&B.cells[bnd[PF].bB + tid]); prefetchL2(&B.cmask[bnd[PF].bB + tid]); prefetchL2(&B.cons[bnd[PF].bB + tid]); }
        Bnd nbLast;
        bounds(L + PF + 1, nbLast.bA, nbLast.eA, nbLast.bB, nbLast.eB);
        // this layer
        const unsigned int mySlotsA = slotsA + (unsigned int)(((((par ^ 1) * 2 + 0) * EX_THREADS + threadIdx.x) * 4) * 8);
        const unsigned int mySlotsB = slotsA + (unsigned int)(((((par ^ 1) * 2 + 1) * EX_THREADS + threadIdx.x) * 4) * 8);
        if (bA + tid < eA) emit(A, 0, par, fa, faceValue(A, fa, slotsInA, mySlotsA));
        if (bB + tid < eB) emit(B, 1, par, fb, faceValue(B, fb, slotsInB, mySlotsB));
        if ((eA - bA) > EX_NT)
            for (int k = bA + tid + EX_NT; k < eA; k += EX_NT) coldFace(A, 0, par, k);
        if ((eB - bB) > EX_NT)
            for (int k = bB + tid + EX_NT; k < eB; k += EX_NT) coldFace(B, 1, par, k);
        if (nextFits) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
        else asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        // next layer's faces: values of their known neighbours (L2 hits, prefetched four layers ago)
        fa.mask = 0; fb.mask = 0;
        if (L < L1) {
            if (bnd[1].bA + tid < bnd[1].eA) finishFace(A, qA[1], fa);
            if (bnd[1].bB + tid < bnd[1].eB) finishFace(B, qB[1], fb);
        }
#pragma unroll
        for (int d = 1; d < PD; ++d) { qA[d] = qA[d + 1]; qB[d] = qB[d + 1]; }
#pragma unroll
        for (int d = 0; d < PF; ++d) bnd[d] = bnd[d + 1];
        bnd[PF] = nbLast;
        if (nextFits) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
        else asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    // no peer may still be pushing into this CTA's shared memory when it exits
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace

static void extrapArrays(Sim* s, double* a, double* b, const uint8_t* unkA, const uint8_t* unkB, ExtrapArray& A, ExtrapArray& B) {
    const Frame& f = s->fr;
    A = ExtrapArray{a, unkA, s->distU, s->distTmp, s->layerCellsU, s->layerMaskU, s->layerConsU, s->layerStartU, s->layerStartU + (s->maxLayers + 2), nullptr, s->nx + 1, s->ny};
    B = ExtrapArray{b, unkB, s->distV, s->distTmp + f.elems, s->layerCellsV, s->layerMaskV, s->layerConsV, s->layerStartV, s->layerStartV + (s->maxLayers + 2), nullptr, s->nx, s->ny + 1};
    A.dist += f.org; A.distTmp += f.org; B.dist += f.org; B.distTmp += f.org;
    A.pos = A.distTmp; B.pos = B.distTmp;
    A.a2 = nullptr; B.a2 = nullptr;
}

// Everything that depends on the unknown masks only (not on the values): distance transform, layer histogram and
// sort, neighbour masks, consumer targets.  runFrame runs it for updateVelocity's extrapolation beside the projection.
int extrapolatePrepare(Sim* s, const uint8_t* unkA, const uint8_t* unkB) {
    const Frame& f = s->fr;
    ExtrapArray A, B;
    extrapArrays(s, nullptr, nullptr, unkA, unkB, A, B);
    const int* anyKnown = s->ctl->anyKnown;
    int* maxLayer = s->ctl->maxLayer;
    size_t histBytes = (size_t)(s->maxLayers + 2) * 2 * sizeof(int);
    CUDA_TRY(cudaMemsetAsync(s->layerStartU, 0, histBytes, s->stream));
    CUDA_TRY(cudaMemsetAsync(s->layerStartV, 0, histBytes, s->stream));
    int rowsMax = s->ny + 1, colsMax = s->nx + 1;
    profBegin(s, 9);
    distRowKernel<<<dim3((rowsMax * 32 + 255) / 256, 2), 256, 0, s->stream>>>(A, B, f.pitch, anyKnown);
    const int nbins = s->maxLayers + 2;
    const size_t dcBytes0 = (size_t)DC_ROWS * 32 * sizeof(int), dcBytes1 = dcBytes0 + (size_t)nbins * sizeof(int);
    {
        static bool dcAttr[16] = {};
        if (!dcAttr[s->device & 15]) {
            CUDA_TRY(cudaFuncSetAttribute(distColKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            dcAttr[s->device & 15] = true;
        }
        if (dcBytes1 > 200 * 1024) { fsim_set_error("grid too large for the layer histogram in shared memory"); return FSIM_E_INVALID; }
    }
    distColKernel<false><<<dim3((colsMax + 31) / 32, 2), 256, dcBytes0, s->stream>>>(A, B, f.pitch, anyKnown, nbins);
    distColKernel<true><<<dim3((colsMax + 31) / 32, 2), 256, dcBytes1, s->stream>>>(A, B, f.pitch, anyKnown, nbins);
    layerScanKernel<<<2, 1024, 0, s->stream>>>(A, B, s->maxLayers, anyKnown, maxLayer);
    layerScatterKernel<<<dim3((colsMax + 31) / 32, (rowsMax + 7) / 8, 2), dim3(32, 8), 0, s->stream>>>(A, B, f.pitch, anyKnown);
    layerConsumersKernel<<<dim3((colsMax + 31) / 32, (rowsMax + 7) / 8, 2), dim3(32, 8), 0, s->stream>>>(A, B, f.pitch, anyKnown);
    profEnd(s);
    s->launches += 6;
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}

// The layer fill proper: one thread-block cluster walks the BFS layers.  part < 0: all of them; part 0 / 1: the layers up to /
// beyond DevCtl::nearLayers (stageUpdateVelocity's split, projection.cu); a2 / b2: optional second destinations.
int extrapolateFill(Sim* s, double* a, double* b, const uint8_t* unkA, const uint8_t* unkB, int part, double* a2, double* b2) {
    const Frame& f = s->fr;
    ExtrapArray A, B;
    extrapArrays(s, a, b, unkA, unkB, A, B);
    A.a2 = a2; B.a2 = b2;
    const int* anyKnown = s->ctl->anyKnown;
    int* maxLayer = s->ctl->maxLayer;
    const size_t exBytes = (size_t)EX_SLOTS * sizeof(double);
    static bool attrSet[16] = {};
    if (!attrSet[s->device & 15]) {
        CUDA_TRY(cudaFuncSetAttribute(layerFillKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)exBytes));
        CUDA_TRY(cudaFuncSetAttribute(layerFillKernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attrSet[s->device & 15] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(EX_CL);
    cfg.blockDim = dim3(EX_THREADS);
    cfg.dynamicSmemBytes = exBytes;
    cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = EX_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const int pitchArg = f.pitch;
    profBegin(s, 7);
    const int* nearPtr = part < 0 ? nullptr : &s->ctl->nearLayers;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, layerFillKernel, A, B, pitchArg, anyKnown, (const int*)maxLayer, nearPtr, part));
    profEnd(s);
    LAUNCH_COUNT(s);
    CUDA_TRY(cudaGetLastError());
    return FSIM_OK;
}

int extrapolatePair(Sim* s, double* a, double* b, const uint8_t* unkA, const uint8_t* unkB) {
    int rc = extrapolatePrepare(s, unkA, unkB);
    if (rc) return rc;
    return extrapolateFill(s, a, b, unkA, unkB, -1, nullptr, nullptr);
}
