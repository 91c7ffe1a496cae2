// K6: the acceleration-structure build ON THE DEVICE (the default builder of spb_bvh_build).
// Replaces BVHAccel::construct / constructRec (reference accelerators/bvh.cc:139-237), which is a
// single-threaded top-down recursion on the host.
//
//   (1) primBoxKernel        per-triangle bounds, world bounds, centroid bounds, "every vertex is float32-exact"
//   (2) binned SAH, top down, THE SAME ALGORITHM as the host builder (bvh_host.cpp, build_binary_sah: 32 bins on each
//       of the three axes, cost = area_L * n_L + area_R * n_R, first strict minimum in (axis, bin) order, 1 primitive per
//       leaf) in the same double arithmetic without FMA contraction (this file is compiled with -fmad=false), so the
//       tree is the host builder's tree as a set of nodes -- and the 8-wide BVH below is byte-identical to the host's
//       (tests/test_trace_gpu.py compares the two through spb_bvh_export).  Parallelisation:
//         large nodes (> 512 primitives)  level by level; a node's primitives are binned in chunks of 1024 by one CTA
//                                         each (bins in shared memory, 64-bit min / max / add atomics on order-preserving
//                                         integer images of the doubles, then one flush to the node's global bins);
//                                         one warp per node sweeps the bins (lane = bin; prefix / suffix scans by
//                                         shuffles) and creates the children; the chunks are partitioned into the other
//                                         index buffer.
//         small nodes (<= 512)            one warp builds the whole subtree in place: bins in shared memory, the same
//                                         sweep, partition through a shared-memory copy of the segment; depth-first with
//                                         the larger child pushed, so the stack never holds more than 10 entries.
//   (3) the cost-optimal 8-wide collapse (Ylitie et al. 2017, sec. 3.1) as a bottom-up pass over the binary tree
//       (dpRowKernel) and a level-by-level emission of the wide nodes (gatherKernel -> exclusive scans -> emitKernel):
//       octant slot assignment, conservative 8-bit quantisation and the leaf-ordered triangle records, in the numbering
//       the host encoder (bvh_host.cpp, encode_wide) produces.
// Nodes and triangles never leave HBM; the host sees four counters per level.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "context.h"

namespace spb {

namespace {

constexpr int kBins = 32;            // == warp size: the sweep maps one lane to one bin
constexpr int kSmall = 512;          // a node with at most this many primitives is finished by one warp
constexpr int kChunk = 1024;         // primitives per CTA of the large-node passes
constexpr double kEmptyLo = DBL_MAX, kEmptyHi = -DBL_MAX;

struct DBox { double lo[3], hi[3]; };
struct DNode {                       // == BinNode (bvh_host.h)
    double lo[3], hi[3];
    int32_t left, right, first, count;
};
static_assert(sizeof(DNode) == sizeof(BinNode), "BinNode layout");

// order-preserving image of a double in an unsigned 64-bit integer: min / max of doubles as integer atomics
__device__ __forceinline__ unsigned long long dkey(double d) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

struct Bin {                         // one (axis, bin) cell; boxes as dkey images
    unsigned long long lo[3], hi[3];     // bounds of the primitives whose centroid falls into the cell
    unsigned long long clo[3], chi[3];   // bounds of those centroids
    uint32_t cnt, pad;
};
static_assert(sizeof(Bin) == 104, "Bin layout");

__device__ __forceinline__ void binClear(Bin* b) {
    const unsigned long long L = dkey(kEmptyLo), H = dkey(kEmptyHi);
#pragma unroll
    for (int k = 0; k < 3; k++) { b->lo[k] = L; b->hi[k] = H; b->clo[k] = L; b->chi[k] = H; }
    b->cnt = 0u; b->pad = 0u;
}

__device__ __forceinline__ double boxArea(const double lo[3], const double hi[3]) {       // Box::area (bvh_host.cpp)
    const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (dx < 0) return 0.0;
    return 2.0 * (dx * dy + dy * dz + dz * dx);
}

__device__ __forceinline__ int binOf(double c, double cmin, double scale) {
    int b = (int)((c - cmin) * scale);
    return b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
}

// ---- (1) primitive boxes ------------------------------------------------------------------------------------------------
struct BuildGlobals {                // device-resident scalars of a build
    unsigned long long wlo[3], whi[3];   // world bounds (dkey)
    unsigned long long clo[3], chi[3];   // centroid bounds (dkey)
    uint32_t f32ok;                      // 1 while every vertex coordinate is float32-exact
    uint32_t nodeCounter;                // next free binary node
    uint32_t nNextLarge, nSmall;         // list sizes of the level being produced
    uint32_t taskCounter;                // dynamic task fetch of smallKernel
    uint32_t maxDepthWide;
    double   sahCost;
};

__global__ void initGlobalsKernel(BuildGlobals* g) {
    const unsigned long long L = dkey(kEmptyLo), H = dkey(kEmptyHi);
    for (int k = 0; k < 3; k++) { g->wlo[k] = L; g->whi[k] = H; g->clo[k] = L; g->chi[k] = H; }
    g->f32ok = 1u; g->nodeCounter = 1u; g->nNextLarge = 0u; g->nSmall = 0u; g->taskCounter = 0u; g->maxDepthWide = 0u; g->sahCost = 0.0;
}

__global__ void __launch_bounds__(256) primBoxKernel(const double* __restrict__ verts, int n, DBox* __restrict__ prims, int32_t* __restrict__ idx,
                                                     BuildGlobals* g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double lo[3] = {kEmptyLo, kEmptyLo, kEmptyLo}, hi[3] = {kEmptyHi, kEmptyHi, kEmptyHi}, c[3] = {0, 0, 0};
    bool exact = true;
    if (i < n) {
        const double* v = verts + (size_t)i * 9;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double a = v[k], b = v[3 + k], cc = v[6 + k];
            lo[k] = fmin(a, fmin(b, cc)); hi[k] = fmax(a, fmax(b, cc));
            c[k] = 0.5 * (lo[k] + hi[k]);
            exact = exact && (double)(float)a == a && (double)(float)b == b && (double)(float)cc == cc;
        }
        DBox bx;
#pragma unroll
        for (int k = 0; k < 3; k++) { bx.lo[k] = lo[k]; bx.hi[k] = hi[k]; }
        prims[i] = bx;
        idx[i] = i;
    }
    // warp reduction, then one set of atomics per warp
    double clo[3], chi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { clo[k] = i < n ? c[k] : kEmptyLo; chi[k] = i < n ? c[k] : kEmptyHi; }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d)); hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
            clo[k] = fmin(clo[k], __shfl_xor_sync(0xffffffffu, clo[k], d)); chi[k] = fmax(chi[k], __shfl_xor_sync(0xffffffffu, chi[k], d));
        }
    }
    const bool allExact = __all_sync(0xffffffffu, exact);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicMin(&g->wlo[k], dkey(lo[k])); atomicMax(&g->whi[k], dkey(hi[k]));
            atomicMin(&g->clo[k], dkey(clo[k])); atomicMax(&g->chi[k], dkey(chi[k]));
        }
        if (!allExact) atomicAnd(&g->f32ok, 0u);
    }
}

__global__ void rootKernel(BuildGlobals* g, DNode* nodes, DBox* cbs, int32_t* parent, int n) {
    DNode r;
    for (int k = 0; k < 3; k++) { r.lo[k] = dunkey(g->wlo[k]); r.hi[k] = dunkey(g->whi[k]); cbs[0].lo[k] = dunkey(g->clo[k]); cbs[0].hi[k] = dunkey(g->chi[k]); }
    r.left = r.right = -1; r.first = 0; r.count = n;
    nodes[0] = r;
    parent[0] = -1;
}

// ---- the sweep: one warp, lane = bin ---------------------------------------------------------------------------------
struct Split {
    int axis, bin;            // axis < 0: no valid split (all centroids coincide): split by index
    int nl;
    double llo[3], lhi[3], rlo[3], rhi[3];         // children bounds
    double lclo[3], lchi[3], rclo[3], rchi[3];     // children centroid bounds
};

// bins: 3 x kBins cells (generic address: shared or global).  cb: the node's centroid bounds.  All 32 lanes call it.
__device__ __forceinline__ void sweepBins(const Bin* bins, int count, const DBox& cb, int lane, Split* out) {
    const unsigned full = 0xffffffffu;
    double bestCost = DBL_MAX;
    int bestAxis = -1, bestBin = -1;
    for (int axis = 0; axis < 3; axis++) {
        if (!(cb.hi[axis] > cb.lo[axis])) continue;
        const Bin* b = bins + axis * kBins + lane;
        int pc = (int)b->cnt;
        double plo[3], phi[3], slo[3], shi[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { plo[k] = dunkey(b->lo[k]); phi[k] = dunkey(b->hi[k]); slo[k] = plo[k]; shi[k] = phi[k]; }
        // inclusive prefix (bins 0..lane) and suffix (bins lane..31)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int c2 = __shfl_up_sync(full, pc, d);
            double tl[3], th[3], ul[3], uh[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                tl[k] = __shfl_up_sync(full, plo[k], d); th[k] = __shfl_up_sync(full, phi[k], d);
                ul[k] = __shfl_down_sync(full, slo[k], d); uh[k] = __shfl_down_sync(full, shi[k], d);
            }
            if (lane >= d) {
                pc += c2;
#pragma unroll
                for (int k = 0; k < 3; k++) { plo[k] = fmin(plo[k], tl[k]); phi[k] = fmax(phi[k], th[k]); }
            }
            if (lane + d < 32) {
#pragma unroll
                for (int k = 0; k < 3; k++) { slo[k] = fmin(slo[k], ul[k]); shi[k] = fmax(shi[k], uh[k]); }
            }
        }
        // split after bin `lane`: left = prefix(lane), right = suffix(lane + 1)
        double rl[3], rh[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { rl[k] = __shfl_down_sync(full, slo[k], 1); rh[k] = __shfl_down_sync(full, shi[k], 1); }
        const int nl = pc, nr = count - nl;
        double cost = DBL_MAX;
        if (lane < kBins - 1 && nl > 0 && nr > 0) cost = boxArea(plo, phi) * nl + boxArea(rl, rh) * nr;
        // first minimum over the bins of this axis
        double mc = cost; int mb = lane;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const double oc = __shfl_xor_sync(full, mc, d);
            const int ob = __shfl_xor_sync(full, mb, d);
            if (oc < mc || (oc == mc && ob < mb)) { mc = oc; mb = ob; }
        }
        if (mc < bestCost) { bestCost = mc; bestAxis = axis; bestBin = mb; }
    }
    out->axis = bestAxis; out->bin = bestBin; out->nl = 0;
    if (bestAxis < 0) return;
    // children of the chosen split: bounds and centroid bounds from the cells of that axis
    const Bin* b = bins + bestAxis * kBins + lane;
    const bool left = lane <= bestBin;
    int cl = left ? (int)b->cnt : 0;
    double v[24];      // [0..5] left lo/hi, [6..11] right lo/hi, [12..17] left clo/chi, [18..23] right clo/chi
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double lo = dunkey(b->lo[k]), hi = dunkey(b->hi[k]), clo = dunkey(b->clo[k]), chi = dunkey(b->chi[k]);
        v[k] = left ? lo : kEmptyLo; v[3 + k] = left ? hi : kEmptyHi;
        v[6 + k] = left ? kEmptyLo : lo; v[9 + k] = left ? kEmptyHi : hi;
        v[12 + k] = left ? clo : kEmptyLo; v[15 + k] = left ? chi : kEmptyHi;
        v[18 + k] = left ? kEmptyLo : clo; v[21 + k] = left ? kEmptyHi : chi;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        cl += __shfl_xor_sync(full, cl, d);
#pragma unroll
        for (int q = 0; q < 24; q++) {
            const double o = __shfl_xor_sync(full, v[q], d);
            const bool isLo = (q % 6) < 3;
            v[q] = isLo ? fmin(v[q], o) : fmax(v[q], o);
        }
    }
    out->nl = cl;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        out->llo[k] = v[k]; out->lhi[k] = v[3 + k]; out->rlo[k] = v[6 + k]; out->rhi[k] = v[9 + k];
        out->lclo[k] = v[12 + k]; out->lchi[k] = v[15 + k]; out->rclo[k] = v[18 + k]; out->rchi[k] = v[21 + k];
    }
}

// bounds + centroid bounds of the primitives at positions [first, first + count) of idx, by one warp (the index-split fallback)
__device__ __forceinline__ void segmentBounds(const DBox* __restrict__ prims, const int32_t* idx, int first, int count, int lane, double lo[3], double hi[3],
                                              double clo[3], double chi[3]) {
#pragma unroll
    for (int k = 0; k < 3; k++) { lo[k] = kEmptyLo; hi[k] = kEmptyHi; clo[k] = kEmptyLo; chi[k] = kEmptyHi; }
    for (int p = lane; p < count; p += 32) {
        const DBox bx = prims[idx[first + p]];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double c = 0.5 * (bx.lo[k] + bx.hi[k]);
            lo[k] = fmin(lo[k], bx.lo[k]); hi[k] = fmax(hi[k], bx.hi[k]); clo[k] = fmin(clo[k], c); chi[k] = fmax(chi[k], c);
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d)); hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
            clo[k] = fmin(clo[k], __shfl_xor_sync(0xffffffffu, clo[k], d)); chi[k] = fmax(chi[k], __shfl_xor_sync(0xffffffffu, chi[k], d));
        }
    }
}

// Creates the two children of `node` (lane 0 writes).  Returns their ids through *l, *r (all lanes).
__device__ __forceinline__ void makeChildren(DNode* nodes, DBox* cbs, int32_t* parent, BuildGlobals* g, int node, int first, int count, const Split& s,
                                             const double llo[3], const double lhi[3], const double rlo[3], const double rhi[3], const double lclo[3],
                                             const double lchi[3], const double rclo[3], const double rchi[3], int nl, int lane, int* l, int* r) {
    int id = 0;
    if (lane == 0) id = (int)atomicAdd(&g->nodeCounter, 2u);
    id = __shfl_sync(0xffffffffu, id, 0);
    *l = id; *r = id + 1;
    if (lane == 0) {
        DNode a, b;
        DBox ca, cb;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            a.lo[k] = llo[k]; a.hi[k] = lhi[k]; b.lo[k] = rlo[k]; b.hi[k] = rhi[k];
            ca.lo[k] = lclo[k]; ca.hi[k] = lchi[k]; cb.lo[k] = rclo[k]; cb.hi[k] = rchi[k];
        }
        a.left = a.right = -1; a.first = first; a.count = nl;
        b.left = b.right = -1; b.first = first + nl; b.count = count - nl;
        nodes[id] = a; nodes[id + 1] = b;
        cbs[id] = ca; cbs[id + 1] = cb;
        parent[id] = node; parent[id + 1] = node;
        nodes[node].left = id; nodes[node].right = id + 1;
    }
    (void)s;
}

// ---- (2a) large nodes, level by level -----------------------------------------------------------------------------------
struct LargeTask { int32_t node, axis, bin, nl; uint32_t leftCur, rightCur; };

__device__ __forceinline__ int findTask(const uint32_t* __restrict__ chunkStart, int nTasks, uint32_t chunk) {
    int lo = 0, hi = nTasks;        // last task with chunkStart <= chunk
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (chunkStart[mid] <= chunk) lo = mid; else hi = mid; }
    return lo;
}

__global__ void chunkCountKernel(const LargeTask* __restrict__ tasks, int nTasks, const DNode* __restrict__ nodes, uint32_t* __restrict__ chunks) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nTasks) chunks[t] = (uint32_t)((nodes[tasks[t].node].count + kChunk - 1) / kChunk);
}

__global__ void clearBinsKernel(Bin* bins, int nCells) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nCells) binClear(bins + i);
}

__global__ void __launch_bounds__(256) largeBinKernel(const LargeTask* __restrict__ tasks, const uint32_t* __restrict__ chunkStart, int nTasks,
                                                      const DNode* __restrict__ nodes, const DBox* __restrict__ cbs, const DBox* __restrict__ prims,
                                                      const int32_t* __restrict__ idx, Bin* __restrict__ gbins) {
    __shared__ Bin s_bins[3 * kBins];
    __shared__ int s_task;
    if (threadIdx.x == 0) s_task = findTask(chunkStart, nTasks, blockIdx.x);
    for (int i = threadIdx.x; i < 3 * kBins; i += blockDim.x) binClear(&s_bins[i]);
    __syncthreads();
    const int t = s_task;
    const DNode nd = nodes[tasks[t].node];
    const DBox cb = cbs[tasks[t].node];
    const int c0 = (int)(blockIdx.x - chunkStart[t]) * kChunk;
    const int c1 = min(nd.count, c0 + kChunk);
    double scale[3]; bool use[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { use[k] = cb.hi[k] > cb.lo[k]; scale[k] = use[k] ? kBins / (cb.hi[k] - cb.lo[k]) : 0.0; }
    for (int p = c0 + threadIdx.x; p < c1; p += blockDim.x) {
        const DBox bx = prims[idx[nd.first + p]];
        double c[3];
        unsigned long long klo[3], khi[3], kc[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { c[k] = 0.5 * (bx.lo[k] + bx.hi[k]); klo[k] = dkey(bx.lo[k]); khi[k] = dkey(bx.hi[k]); kc[k] = dkey(c[k]); }
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (!use[a]) continue;
            Bin* b = &s_bins[a * kBins + binOf(c[a], cb.lo[a], scale[a])];
            atomicAdd(&b->cnt, 1u);
#pragma unroll
            for (int k = 0; k < 3; k++) { atomicMin(&b->lo[k], klo[k]); atomicMax(&b->hi[k], khi[k]); atomicMin(&b->clo[k], kc[k]); atomicMax(&b->chi[k], kc[k]); }
        }
    }
    __syncthreads();
    Bin* dst = gbins + (size_t)t * 3 * kBins;
    for (int i = threadIdx.x; i < 3 * kBins; i += blockDim.x) {
        const Bin& b = s_bins[i];
        if (b.cnt == 0u) continue;
        atomicAdd(&dst[i].cnt, b.cnt);
#pragma unroll
        for (int k = 0; k < 3; k++) { atomicMin(&dst[i].lo[k], b.lo[k]); atomicMax(&dst[i].hi[k], b.hi[k]); atomicMin(&dst[i].clo[k], b.clo[k]); atomicMax(&dst[i].chi[k], b.chi[k]); }
    }
}

// `buf`: which index buffer holds the children's segments AFTER this level's partition (recorded with the small tasks)
__global__ void __launch_bounds__(128) largeSplitKernel(LargeTask* tasks, int nTasks, DNode* nodes, DBox* cbs, int32_t* parent, const DBox* __restrict__ prims,
                                                        const int32_t* __restrict__ idx, const Bin* __restrict__ gbins, BuildGlobals* g, LargeTask* nextLarge,
                                                        int2* smallTasks, int buf) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= nTasks) return;
    const int node = tasks[t].node;
    const DNode nd = nodes[node];
    const DBox cb = cbs[node];
    Split s;
    sweepBins(gbins + (size_t)t * 3 * kBins, nd.count, cb, lane, &s);
    int l, r, nl;
    if (s.axis >= 0) {
        nl = s.nl;
        makeChildren(nodes, cbs, parent, g, node, nd.first, nd.count, s, s.llo, s.lhi, s.rlo, s.rhi, s.lclo, s.lchi, s.rclo, s.rchi, nl, lane, &l, &r);
    } else {
        // every centroid coincides: split by index, nothing moves (bvh_host.cpp: mid = first + count / 2; the host also puts the
        // segment in ascending primitive order first, which a node of more than kSmall coincident primitives does not get here)
        nl = nd.count / 2;
        double a0[3], a1[3], a2[3], a3[3], b0[3], b1[3], b2[3], b3[3];
        segmentBounds(prims, idx, nd.first, nl, lane, a0, a1, a2, a3);
        segmentBounds(prims, idx, nd.first + nl, nd.count - nl, lane, b0, b1, b2, b3);
        makeChildren(nodes, cbs, parent, g, node, nd.first, nd.count, s, a0, a1, b0, b1, a2, a3, b2, b3, nl, lane, &l, &r);
    }
    if (lane == 0) {
        tasks[t].axis = s.axis; tasks[t].bin = s.bin; tasks[t].nl = nl; tasks[t].leftCur = 0u; tasks[t].rightCur = 0u;
        const int cnt[2] = {nl, nd.count - nl}, id[2] = {l, r};
        for (int c = 0; c < 2; c++) {
            if (cnt[c] > kSmall) {
                LargeTask nt; nt.node = id[c]; nt.axis = -1; nt.bin = 0; nt.nl = 0; nt.leftCur = nt.rightCur = 0u;
                nextLarge[atomicAdd(&g->nNextLarge, 1u)] = nt;
            } else {
                smallTasks[atomicAdd(&g->nSmall, 1u)] = make_int2(id[c], buf);
            }
        }
    }
}

__global__ void __launch_bounds__(256) largePartitionKernel(LargeTask* tasks, const uint32_t* __restrict__ chunkStart, int nTasks, const DNode* __restrict__ nodes,
                                                            const DBox* __restrict__ cbs, const DBox* __restrict__ prims, const int32_t* __restrict__ in,
                                                            int32_t* __restrict__ out) {
    __shared__ int s_task;
    __shared__ uint32_t s_baseL, s_baseR, s_wL[8], s_wV[8];
    if (threadIdx.x == 0) s_task = findTask(chunkStart, nTasks, blockIdx.x);
    __syncthreads();
    const int t = s_task;
    const LargeTask tk = tasks[t];
    const DNode nd = nodes[tk.node];
    const int c0 = (int)(blockIdx.x - chunkStart[t]) * kChunk;
    const int c1 = min(nd.count, c0 + kChunk);
    if (tk.axis < 0) {
        for (int p = c0 + threadIdx.x; p < c1; p += blockDim.x) out[nd.first + p] = in[nd.first + p];
        return;
    }
    const DBox cb = cbs[tk.node];
    const double cmin = cb.lo[tk.axis], scale = kBins / (cb.hi[tk.axis] - cmin);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int kPer = kChunk / 256;
    int32_t id[kPer]; bool valid[kPer], left[kPer];
#pragma unroll
    for (int q = 0; q < kPer; q++) {
        const int p = c0 + q * 256 + threadIdx.x;
        valid[q] = p < c1; left[q] = false; id[q] = 0;
        if (valid[q]) {
            id[q] = in[nd.first + p];
            const DBox bx = prims[id[q]];
            const double c = 0.5 * (bx.lo[tk.axis] + bx.hi[tk.axis]);
            left[q] = binOf(c, cmin, scale) <= tk.bin;
        }
    }
    // ranks: per q, ballot within the warp; warps and q's are ordered (q major, then warp)
    uint32_t rankL[kPer], rankR[kPer], cntL[kPer], cntV[kPer];
#pragma unroll
    for (int q = 0; q < kPer; q++) {
        const unsigned mL = __ballot_sync(0xffffffffu, valid[q] && left[q]), mV = __ballot_sync(0xffffffffu, valid[q]);
        const unsigned below = (1u << lane) - 1u;
        rankL[q] = __popc(mL & below); rankR[q] = __popc(mV & ~mL & below);
        cntL[q] = __popc(mL); cntV[q] = __popc(mV);
    }
    // per-warp totals
    uint32_t wL = 0, wV = 0;
#pragma unroll
    for (int q = 0; q < kPer; q++) { wL += cntL[q]; wV += cntV[q]; }
    if (lane == 0) { s_wL[wid] = wL; s_wV[wid] = wV; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tl = 0, tv = 0;
        for (int w = 0; w < 8; w++) { tl += s_wL[w]; tv += s_wV[w]; }
        s_baseL = atomicAdd(&tasks[t].leftCur, tl);
        s_baseR = atomicAdd(&tasks[t].rightCur, tv - tl);
    }
    __syncthreads();
    uint32_t offL = s_baseL, offR = s_baseR;
    for (int w = 0; w < wid; w++) { offL += s_wL[w]; offR += s_wV[w] - s_wL[w]; }
#pragma unroll
    for (int q = 0; q < kPer; q++) {
        if (valid[q]) {
            if (left[q]) out[nd.first + offL + rankL[q]] = id[q];
            else out[nd.first + tk.nl + offR + rankR[q]] = id[q];
        }
        offL += cntL[q]; offR += cntV[q] - cntL[q];
    }
}

// ---- (2b) small nodes: one warp per subtree ------------------------------------------------------------------------------
constexpr int kSmallWarps = 2;
__global__ void __launch_bounds__(32 * kSmallWarps) smallKernel(const int2* __restrict__ smallTasks, int nTasks, DNode* nodes, DBox* cbs, int32_t* parent,
                                                                const DBox* __restrict__ prims, const int32_t* __restrict__ bufA, const int32_t* __restrict__ bufB,
                                                                int32_t* __restrict__ order, BuildGlobals* g) {
    __shared__ Bin s_bins[kSmallWarps][3 * kBins];
    __shared__ int32_t s_tmp[kSmallWarps][kSmall];
    __shared__ int32_t s_stack[kSmallWarps][16];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    Bin* bins = s_bins[wid];
    int32_t* tmp = s_tmp[wid];
    int32_t* stack = s_stack[wid];
    for (;;) {
        int task = 0;
        if (lane == 0) task = (int)atomicAdd(&g->taskCounter, 1u);
        task = __shfl_sync(full, task, 0);
        if (task >= nTasks) return;
        const int2 tk = smallTasks[task];
        {   // bring the segment into the final order array
            const DNode root = nodes[tk.x];
            const int32_t* src = tk.y ? bufB : bufA;
            for (int p = lane; p < root.count; p += 32) order[root.first + p] = src[root.first + p];
        }
        __syncwarp();
        int sp = 0;
        if (lane == 0) stack[0] = tk.x;
        sp = 1;
        __syncwarp();
        while (sp > 0) {
            const int node = stack[--sp];
            __syncwarp();
            const DNode nd = nodes[node];
            if (nd.count <= 1) continue;
            const DBox cb = cbs[node];
            for (int i = lane; i < 3 * kBins; i += 32) binClear(&bins[i]);
            __syncwarp();
            double scale[3]; bool use[3];
#pragma unroll
            for (int k = 0; k < 3; k++) { use[k] = cb.hi[k] > cb.lo[k]; scale[k] = use[k] ? kBins / (cb.hi[k] - cb.lo[k]) : 0.0; }
            for (int p = lane; p < nd.count; p += 32) {
                const DBox bx = prims[order[nd.first + p]];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (!use[a]) continue;
                    const double ca = 0.5 * (bx.lo[a] + bx.hi[a]);
                    Bin* b = &bins[a * kBins + binOf(ca, cb.lo[a], scale[a])];
                    atomicAdd(&b->cnt, 1u);
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const unsigned long long kc = dkey(0.5 * (bx.lo[k] + bx.hi[k]));
                        atomicMin(&b->lo[k], dkey(bx.lo[k])); atomicMax(&b->hi[k], dkey(bx.hi[k])); atomicMin(&b->clo[k], kc); atomicMax(&b->chi[k], kc);
                    }
                }
            }
            __syncwarp();
            Split s;
            sweepBins(bins, nd.count, cb, lane, &s);
            int l, r, nl;
            if (s.axis >= 0) {
                nl = s.nl;
                // partition through a copy of the segment: lefts keep their order, then the rights
                const double cmin = cb.lo[s.axis], sc = kBins / (cb.hi[s.axis] - cmin);
                int doneL = 0, doneR = 0;
                for (int p0 = 0; p0 < nd.count; p0 += 32) {
                    const int p = p0 + lane;
                    const bool valid = p < nd.count;
                    int id = 0; bool left = false;
                    if (valid) {
                        id = order[nd.first + p];
                        const DBox bx = prims[id];
                        left = binOf(0.5 * (bx.lo[s.axis] + bx.hi[s.axis]), cmin, sc) <= s.bin;
                    }
                    const unsigned mL = __ballot_sync(full, valid && left), mV = __ballot_sync(full, valid);
                    const unsigned below = (1u << lane) - 1u;
                    if (valid) {
                        if (left) tmp[doneL + __popc(mL & below)] = id;
                        else tmp[nl + doneR + __popc(mV & ~mL & below)] = id;
                    }
                    doneL += __popc(mL); doneR += __popc(mV & ~mL);
                }
                __syncwarp();
                for (int p = lane; p < nd.count; p += 32) order[nd.first + p] = tmp[p];
                __syncwarp();
                makeChildren(nodes, cbs, parent, g, node, nd.first, nd.count, s, s.llo, s.lhi, s.rlo, s.rhi, s.lclo, s.lchi, s.rclo, s.rchi, nl, lane, &l, &r);
            } else {
                // every centroid coincides: split by index, in ascending primitive order (bvh_host.cpp does the same), by a
                // rank sort through the shared-memory copy -- these segments are a handful of primitives
                for (int p = lane; p < nd.count; p += 32) {
                    const int32_t id = order[nd.first + p];
                    int rank = 0;
                    for (int q2 = 0; q2 < nd.count; q2++) rank += order[nd.first + q2] < id;
                    tmp[rank] = id;
                }
                __syncwarp();
                for (int p = lane; p < nd.count; p += 32) order[nd.first + p] = tmp[p];
                __syncwarp();
                nl = nd.count / 2;
                double a0[3], a1[3], a2[3], a3[3], b0[3], b1[3], b2[3], b3[3];
                segmentBounds(prims, order, nd.first, nl, lane, a0, a1, a2, a3);
                segmentBounds(prims, order, nd.first + nl, nd.count - nl, lane, b0, b1, b2, b3);
                makeChildren(nodes, cbs, parent, g, node, nd.first, nd.count, s, a0, a1, b0, b1, a2, a3, b2, b3, nl, lane, &l, &r);
            }
            __syncwarp();
            // the larger child is pushed first, the smaller one is taken next: the stack holds at most log2(kSmall) + 1 entries
            const int nr = nd.count - nl;
            if (lane == 0) {
                if (nl >= nr) { stack[sp] = l; stack[sp + 1] = r; } else { stack[sp] = r; stack[sp + 1] = l; }
            }
            sp += 2;
            __syncwarp();
        }
    }
}

// ---- (3) cost-optimal collapse to 8-wide -------------------------------------------------------------------------------
struct DpRow { float c[8]; uint8_t k[8]; uint8_t k8, leaf, pad0, pad1; };       // c[i-1] = C(n, i), k[i-1] = left share (0: same as i-1)

__device__ __forceinline__ double nodeAreaD(const DNode& n) { return boxArea(n.lo, n.hi); }

// a row written by another thread block (the child that arrived first): read past L1
__device__ __forceinline__ DpRow loadRowCg(const DpRow* p) {
    static_assert(sizeof(DpRow) == 44, "DpRow layout");
    union { DpRow r; uint32_t w[11]; } u;
#pragma unroll
    for (int i = 0; i < 11; i++) u.w[i] = __ldcg((const uint32_t*)p + i);
    return u.r;
}

__device__ void computeRow(const DNode* nodes, DpRow* rows, int b, double rootArea, int maxLeaf, float cTri) {
    const DNode n = nodes[b];
    DpRow R;
    const float inf = 3.0e38f, cNode = 1.0f;
    const float A = (float)(nodeAreaD(n) / rootArea);
    const bool isLeaf = n.left < 0 && n.right < 0;
    const float cLeaf = (n.count <= maxLeaf || isLeaf) ? A * (float)n.count * cTri : inf;
    if (isLeaf) {
        for (int i = 0; i < 8; i++) { R.c[i] = cLeaf; R.k[i] = 0; }
        R.k8 = 0; R.leaf = 1; R.pad0 = R.pad1 = 0;
        rows[b] = R;
        return;
    }
    const DpRow L = loadRowCg(rows + n.left), Rr = loadRowCg(rows + n.right);
    float D[9]; uint8_t Dk[9];
    for (int j = 2; j <= 8; j++) {
        float best = inf; uint8_t bk = 1;
        for (int k = 1; k < j; k++) {
            const float v = L.c[k - 1] + Rr.c[j - k - 1];
            if (v < best) { best = v; bk = (uint8_t)k; }
        }
        D[j] = best; Dk[j] = bk;
    }
    const float cInner = A * cNode + D[8];
    R.k8 = Dk[8];
    R.leaf = cLeaf <= cInner ? 1 : 0;
    R.c[0] = R.leaf ? cLeaf : cInner; R.k[0] = 0;
    for (int i = 2; i <= 7; i++) {
        if (D[i] < R.c[i - 2]) { R.c[i - 1] = D[i]; R.k[i - 1] = Dk[i]; }
        else { R.c[i - 1] = R.c[i - 2]; R.k[i - 1] = 0; }
    }
    R.c[7] = R.c[6]; R.k[7] = 0;
    R.pad0 = R.pad1 = 0;
    rows[b] = R;
}

// bottom up: a thread starts at every leaf; the second child to arrive at a node computes its row and carries on
__global__ void __launch_bounds__(256) dpRowKernel(const DNode* __restrict__ nodes, int nNodes, const int32_t* __restrict__ parent, DpRow* rows,
                                                   unsigned int* flags, int maxLeaf, float cTri) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nNodes) return;
    if (!(nodes[b].left < 0 && nodes[b].right < 0)) return;
    const double rootArea = fmax(nodeAreaD(nodes[0]), 1e-300);
    computeRow(nodes, rows, b, rootArea, maxLeaf, cTri);
    __threadfence();
    int p = parent[b];
    while (p >= 0) {
        if (atomicAdd(&flags[p], 1u) == 0u) return;
        __threadfence();
        computeRow(nodes, rows, p, rootArea, maxLeaf, cTri);
        __threadfence();
        p = parent[p];
    }
}

struct WideWork { int32_t bnode; int32_t nch; int32_t ch[8]; uint8_t leaf[8]; };       // one wide node being emitted

// children of the forest C(b, i) (bvh_host.cpp, collectDp), appended to w->ch in left-to-right order
__device__ void collectDp(const DNode* __restrict__ nodes, const DpRow* __restrict__ rows, int b0, int i0, WideWork* w) {
    int fb[12]; uint8_t fi[12]; int nf = 0;      // the parts of the entries on the stack are >= 1 each and sum to <= 8
    fb[nf] = b0; fi[nf] = (uint8_t)i0; nf++;
    while (nf > 0) {
        nf--;
        const int b = fb[nf]; int i = (int)fi[nf];
        const DNode& n = nodes[b];
        if (n.left < 0 && n.right < 0) { w->leaf[w->nch] = 1; w->ch[w->nch++] = b; continue; }
        while (i > 1 && rows[b].k[i - 1] == 0) i--;
        if (i == 1) { w->leaf[w->nch] = rows[b].leaf; w->ch[w->nch++] = b; continue; }
        const int k = rows[b].k[i - 1];
        fb[nf] = n.right; fi[nf] = (uint8_t)(i - k); nf++;        // popped after the left part
        fb[nf] = n.left; fi[nf] = (uint8_t)k; nf++;
    }
}

// pass 1 of a level: the children of every wide node of the level, and how many inner children / triangles it brings
__global__ void __launch_bounds__(128) gatherKernel(const DNode* __restrict__ nodes, const DpRow* __restrict__ rows, const int32_t* __restrict__ levelNodes,
                                                    int nLevel, int maxLeaf, bool isRoot, WideWork* work, uint32_t* nInner, uint32_t* nTris) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nLevel) return;
    WideWork w;
    w.bnode = levelNodes[i]; w.nch = 0;
    for (int k = 0; k < 8; k++) { w.ch[k] = -1; w.leaf[k] = 0; }
    const DNode& n = nodes[w.bnode];
    const bool leafLike = n.count <= maxLeaf || (n.left < 0 && n.right < 0);
    if (isRoot && leafLike) { w.ch[0] = w.bnode; w.leaf[0] = 1; w.nch = 1; }          // degenerate root: a single leaf child
    else if (n.left < 0 && n.right < 0) { w.ch[0] = w.bnode; w.leaf[0] = 1; w.nch = 1; }
    else { const int k = rows[w.bnode].k8; collectDp(nodes, rows, n.left, k, &w); collectDp(nodes, rows, n.right, 8 - k, &w); }
    uint32_t ni = 0, nt = 0;
    for (int c = 0; c < w.nch; c++) { if (w.leaf[c]) nt += (uint32_t)nodes[w.ch[c]].count; else ni++; }
    work[i] = w; nInner[i] = ni; nTris[i] = nt;
}

__device__ __forceinline__ int ceilLog2d(double v) {       // smallest e with 2^e >= v (v > 0), exactly (bvh_host.cpp uses the same rule)
    int x;
    const double m = frexp(v, &x);                         // v = m * 2^x, m in [0.5, 1)
    return m == 0.5 ? x - 1 : x;
}

// pass 2 of a level: slot assignment, quantisation, node and triangle records (bvh_host.cpp, encode_wide, per wide node)
__global__ void __launch_bounds__(128) emitKernel(const DNode* __restrict__ nodes, const int32_t* __restrict__ order, const double* __restrict__ verts,
                                                  const WideWork* __restrict__ work, const uint32_t* __restrict__ innerScan, const uint32_t* __restrict__ triScan,
                                                  int nLevel, uint32_t levelBase, uint32_t nextBase, uint32_t triBase, double inflate, int triFormat,
                                                  WideNode* wide, void* tris, int32_t* nextLevelNodes, BuildGlobals* g, double rootArea) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nLevel) return;
    const WideWork w = work[i];
    const DNode bn = nodes[w.bnode];
    const int nch = w.nch;
    // ---- slot assignment: child whose offset from the node centre aligns best with the slot diagonal; greedy on the largest dot product
    double ctr[3];
    for (int k = 0; k < 3; k++) ctr[k] = 0.5 * (bn.lo[k] + bn.hi[k]);
    int slotOf[8]; bool slotUsed[8], chDone[8];
    for (int k = 0; k < 8; k++) { slotOf[k] = 0; slotUsed[k] = false; chDone[k] = false; }
    for (int it = 0; it < nch; it++) {
        double bestV = -DBL_MAX; int bc = -1, bs = -1;
        for (int c = 0; c < nch; c++) {
            if (chDone[c]) continue;
            const DNode& cn = nodes[w.ch[c]];
            double off[3];
            for (int k = 0; k < 3; k++) off[k] = 0.5 * (cn.lo[k] + cn.hi[k]) - ctr[k];
            for (int s = 0; s < 8; s++) {
                if (slotUsed[s]) continue;
                const double v = ((s & 1) ? off[0] : -off[0]) + ((s & 2) ? off[1] : -off[1]) + ((s & 4) ? off[2] : -off[2]);
                if (v > bestV) { bestV = v; bc = c; bs = s; }
            }
        }
        chDone[bc] = true; slotUsed[bs] = true; slotOf[bc] = bs;
    }
    int slotChild[8]; bool slotLeaf[8];
    for (int s = 0; s < 8; s++) { slotChild[s] = -1; slotLeaf[s] = false; }
    for (int c = 0; c < nch; c++) { slotChild[slotOf[c]] = w.ch[c]; slotLeaf[slotOf[c]] = w.leaf[c] != 0; }

    // ---- quantisation grid
    WideNode wn;
    memset(&wn, 0, sizeof(wn));
    double plo[3], ext[3];
    for (int k = 0; k < 3; k++) {
        double mn = DBL_MAX, mx = -DBL_MAX;
        for (int c = 0; c < nch; c++) {
            mn = fmin(mn, nodes[w.ch[c]].lo[k] - inflate);
            mx = fmax(mx, nodes[w.ch[c]].hi[k] + inflate);
        }
        wn.p[k] = __double2float_rd(mn);
        plo[k] = (double)wn.p[k];
        ext[k] = mx - plo[k];
    }
    for (int k = 0; k < 3; k++) {
        int e = ceilLog2d(fmax(ext[k], 1e-300) / 255.0);
        while (ldexp(255.0, e) < ext[k]) e++;
        if (e < -100) e = -100;
        for (;;) {
            bool ok = true;
            const double step = ldexp(1.0, e);
            for (int s = 0; s < 8 && ok; s++) {
                if (slotChild[s] < 0) continue;
                const double lo = nodes[slotChild[s]].lo[k] - inflate, hi = nodes[slotChild[s]].hi[k] + inflate;
                double ql = floor((lo - plo[k]) / step);
                double qh = ceil((hi - plo[k]) / step);
                while (plo[k] + ql * step > lo) ql -= 1.0;
                while (plo[k] + qh * step < hi) qh += 1.0;
                if (ql < 0.0) ql = 0.0;
                if (qh > 255.0) { ok = false; break; }
                wn.qlo[k][s] = (uint8_t)ql;
                wn.qhi[k][s] = (uint8_t)qh;
            }
            if (ok) break;
            e++;
        }
        wn.e[k] = (uint8_t)(e + 127);
    }
    // ---- children: inner ones get contiguous wide indices in slot order; leaves get triangles
    wn.child_base = nextBase + innerScan[i];
    wn.tri_base = triBase + triScan[i];
    uint32_t innerAt = innerScan[i], triAt = triBase + triScan[i];
    int triOff = 0;
    double cost = 0.0;
    for (int s = 0; s < 8; s++) {
        const int c = slotChild[s];
        if (c < 0) { wn.meta[s] = 0; continue; }
        const DNode& cn = nodes[c];
        const double relA = nodeAreaD(cn) / rootArea;
        if (slotLeaf[s]) {
            const int n = cn.count;
            const uint8_t unary = (uint8_t)((1u << n) - 1u);
            wn.meta[s] = (uint8_t)((unary << 5) | (uint8_t)triOff);
            int32_t ids[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};       // ascending primitive index within a leaf (bvh_host.cpp, encode_wide)
            for (int q = 0; q < n && q < 3; q++) ids[q] = order[cn.first + q];
            if (ids[0] > ids[1]) { const int32_t t = ids[0]; ids[0] = ids[1]; ids[1] = t; }
            if (ids[1] > ids[2]) { const int32_t t = ids[1]; ids[1] = ids[2]; ids[2] = t; }
            if (ids[0] > ids[1]) { const int32_t t = ids[0]; ids[0] = ids[1]; ids[1] = t; }
            for (int q = 0; q < n; q++) {
                const int32_t prim = ids[q];
                const double* v = verts + (size_t)prim * 9;
                if (triFormat == 0) {
                    TriF32 t;
                    for (int k = 0; k < 3; k++) { t.v0[k] = (float)v[k]; t.v1[k] = (float)v[3 + k]; t.v2[k] = (float)v[6 + k]; }
                    t.id = prim; t.rank = prim; t.pad = 0;
                    ((TriF32*)tris)[triAt] = t;
                } else {
                    TriF64 t;
                    for (int k = 0; k < 9; k++) t.v[k] = v[k];
                    t.id = prim; t.rank = prim;
                    ((TriF64*)tris)[triAt] = t;
                }
                triAt++;
            }
            triOff += n;
            cost += relA * n;
        } else {
            wn.imask |= (uint8_t)(1u << s);
            wn.meta[s] = (uint8_t)((1u << 5) | (24 + s));
            nextLevelNodes[innerAt++] = c;
            cost += relA;
        }
    }
    wide[levelBase + i] = wn;
    atomicAdd(&g->sahCost, cost);
}

#define SB(call)                                                                                                     \
    do {                                                                                                             \
        const cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess) {                                                                                     \
            freeAll();                                                                                               \
            cudaOk(ctx, e_, #call);                                                                                  \
            return e_ == cudaErrorMemoryAllocation ? SPB_ERR_OOM : SPB_ERR_CUDA;                                     \
        }                                                                                                            \
    } while (0)

}  // namespace

// Builds the 8-wide BVH of ctx->verts on the context's GPU, leaves it in ctx->d_nodes / ctx->d_tris and fills ctx->bvh's
// statistics (its host arrays stay empty) and ctx->sp.
int buildSahDevice(spb_ctx* ctx, int maxLeaf) {
    const int64_t n64 = ctx->n_tris;
    HostBVH& hb = ctx->bvh;
    hb.nodes.clear(); hb.tris.clear();
    hb.n_tris = n64; hb.max_depth = 0; hb.sah_cost = 0.0; hb.n_binary_nodes = 0; hb.n_wide = 0; hb.tri_bytes_device = 0;
    if (maxLeaf < 1) maxLeaf = 1;
    if (maxLeaf > 3) maxLeaf = 3;
    if (n64 <= 0) return SPB_OK;
    const int n = (int)n64;
    cudaStream_t st = ctx->stream;
    // SPICA_BUILD_TIMING=1: phase times on stderr (allocation + upload, binary tree, collapse + emission, hand-over)
    const bool timing = std::getenv("SPICA_BUILD_TIMING") != nullptr;
    auto tPrev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        cudaStreamSynchronize(st);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[spb build] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - tPrev).count());
        tPrev = now;
    };

    double* d_verts = nullptr; DBox* d_prims = nullptr; int32_t *d_idx[2] = {nullptr, nullptr}, *d_order = nullptr, *d_parent = nullptr;
    DNode* d_nodes = nullptr; DBox* d_cbs = nullptr; BuildGlobals* d_g = nullptr; LargeTask* d_large[2] = {nullptr, nullptr};
    int2* d_small = nullptr; uint32_t *d_chunks = nullptr, *d_chunkStart = nullptr; Bin* d_bins = nullptr; void* d_scan = nullptr;
    DpRow* d_rows = nullptr; unsigned int* d_flags = nullptr; WideWork* d_work = nullptr; uint32_t *d_nInner = nullptr, *d_nTris = nullptr, *d_innerScan = nullptr, *d_triScan = nullptr;
    int32_t* d_level[2] = {nullptr, nullptr};
    WideNode* d_wide = nullptr; void* d_tris = nullptr; size_t trisBytes = 0;
    bool keepWide = false;
    // work arrays: bumped out of the context's arena (context.h), which keeps its memory for the next build; the triangle
    // records and the final node array, which other streams, GPUs and processes read, are plain allocations
    ScratchArena& arena = ctx->build_arena;
    arena.reset();
    const size_t arenaGrow = (size_t)n * 480 + ((size_t)64 << 20);      // one build's worth: the peak is ~ 460 B per triangle
    auto freeAll = [&]() {
        cudaStreamSynchronize(st);
        if (!keepWide) { cudaFree(d_tris); d_tris = nullptr; }
    };
#define SALLOC(ptr, bytes) SB(arena.alloc((void**)&(ptr), (bytes), arenaGrow))
    const int nNodesMax = 2 * n - 1;
    const int maxLarge = n / kSmall + 2;                     // large nodes of one level are disjoint and hold > kSmall primitives each
    const int maxSmallTasks = 2 * maxLarge + 2;              // every small task is the child of a large node (or the root)
    SALLOC(d_verts, (size_t)n * 9 * sizeof(double));
    SALLOC(d_order, (size_t)n * 4);
    SALLOC(d_parent, (size_t)nNodesMax * 4);
    SALLOC(d_nodes, (size_t)nNodesMax * sizeof(DNode));
    SALLOC(d_g, sizeof(BuildGlobals));
    SALLOC(d_large[0], (size_t)maxLarge * sizeof(LargeTask)); SALLOC(d_large[1], (size_t)maxLarge * sizeof(LargeTask));
    SALLOC(d_small, (size_t)maxSmallTasks * sizeof(int2));
    SALLOC(d_chunks, (size_t)(maxLarge + 1) * 4); SALLOC(d_chunkStart, (size_t)(maxLarge + 1) * 4);
    const ScratchArena::Mark afterTree = arena.mark();        // what follows dies with the binary tree's construction
    SALLOC(d_prims, (size_t)n * sizeof(DBox));
    SALLOC(d_idx[0], (size_t)n * 4); SALLOC(d_idx[1], (size_t)n * 4);
    SALLOC(d_cbs, (size_t)nNodesMax * sizeof(DBox));
    SALLOC(d_bins, (size_t)maxLarge * 3 * kBins * sizeof(Bin));
    size_t scanBytes = 0;
    SB(cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, d_chunks, d_chunkStart, std::max(maxLarge + 1, n / 2 + 2), st));
    SALLOC(d_scan, scanBytes);

    lap("allocations");
    SB(cudaMemcpyAsync(d_verts, ctx->geo->verts.data(), (size_t)n * 9 * sizeof(double), cudaMemcpyHostToDevice, st));
    lap("vertex upload");
    initGlobalsKernel<<<1, 1, 0, st>>>(d_g);
    primBoxKernel<<<(n + 255) / 256, 256, 0, st>>>(d_verts, n, d_prims, d_idx[0], d_g);
    rootKernel<<<1, 1, 0, st>>>(d_g, d_nodes, d_cbs, d_parent, n);
    int launches = 3;
    BuildGlobals hg;

    // ---- (2) the binary tree
    int nLarge = 0, cur = 0;
    if (n > kSmall) {
        LargeTask rootTask; rootTask.node = 0; rootTask.axis = -1; rootTask.bin = 0; rootTask.nl = 0; rootTask.leftCur = rootTask.rightCur = 0u;
        SB(cudaMemcpyAsync(d_large[0], &rootTask, sizeof(rootTask), cudaMemcpyHostToDevice, st));
        nLarge = 1;
    } else {
        const int2 t0 = make_int2(0, 0);
        SB(cudaMemcpyAsync(d_small, &t0, sizeof(t0), cudaMemcpyHostToDevice, st));
        const uint32_t one = 1u;
        SB(cudaMemcpyAsync(&d_g->nSmall, &one, 4, cudaMemcpyHostToDevice, st));
    }
    int lt = 0;                     // which task list is current
    while (nLarge > 0) {
        chunkCountKernel<<<(nLarge + 255) / 256, 256, 0, st>>>(d_large[lt], nLarge, d_nodes, d_chunks);
        SB(cub::DeviceScan::ExclusiveSum(d_scan, scanBytes, d_chunks, d_chunkStart, nLarge + 1, st));     // entry nLarge = the total (its input is never read past)
        uint32_t totalChunks = 0, lastChunks = 0, lastStart = 0;
        SB(cudaMemcpyAsync(&lastStart, d_chunkStart + (nLarge - 1), 4, cudaMemcpyDeviceToHost, st));
        SB(cudaMemcpyAsync(&lastChunks, d_chunks + (nLarge - 1), 4, cudaMemcpyDeviceToHost, st));
        SB(cudaStreamSynchronize(st));
        totalChunks = lastStart + lastChunks;
        const int nCells = nLarge * 3 * kBins;
        clearBinsKernel<<<(nCells + 255) / 256, 256, 0, st>>>(d_bins, nCells);
        largeBinKernel<<<totalChunks, 256, 0, st>>>(d_large[lt], d_chunkStart, nLarge, d_nodes, d_cbs, d_prims, d_idx[cur], d_bins);
        const uint32_t zero = 0u;
        SB(cudaMemcpyAsync(&d_g->nNextLarge, &zero, 4, cudaMemcpyHostToDevice, st));
        largeSplitKernel<<<(nLarge * 32 + 127) / 128, 128, 0, st>>>(d_large[lt], nLarge, d_nodes, d_cbs, d_parent, d_prims, d_idx[cur], d_bins, d_g, d_large[lt ^ 1], d_small,
                                                                   cur ^ 1);
        largePartitionKernel<<<totalChunks, 256, 0, st>>>(d_large[lt], d_chunkStart, nLarge, d_nodes, d_cbs, d_prims, d_idx[cur], d_idx[cur ^ 1]);
        launches += 6;
        SB(cudaMemcpyAsync(&hg, d_g, sizeof(hg), cudaMemcpyDeviceToHost, st));
        SB(cudaStreamSynchronize(st));
        SB(cudaGetLastError());
        nLarge = (int)hg.nNextLarge;
        if (nLarge > maxLarge || (int)hg.nSmall > maxSmallTasks) { freeAll(); return fail(ctx, SPB_ERR_CUDA, "internal: builder task list overflow"); }
        lt ^= 1; cur ^= 1;
    }
    SB(cudaMemcpyAsync(&hg, d_g, sizeof(hg), cudaMemcpyDeviceToHost, st));
    SB(cudaStreamSynchronize(st));
    const int nSmall = (int)hg.nSmall;
    if (nSmall > 0) {
        const int grid = std::min((nSmall + kSmallWarps - 1) / kSmallWarps, ctx->sm_count * 16);
        smallKernel<<<grid, 32 * kSmallWarps, 0, st>>>(d_small, nSmall, d_nodes, d_cbs, d_parent, d_prims, d_idx[0], d_idx[1], d_order, d_g);
        launches++;
    }
    SB(cudaMemcpyAsync(&hg, d_g, sizeof(hg), cudaMemcpyDeviceToHost, st));
    SB(cudaStreamSynchronize(st));
    SB(cudaGetLastError());
    lap("binary tree");
    const int nNodes = (int)hg.nodeCounter;
    if (nNodes != nNodesMax) { freeAll(); return fail(ctx, SPB_ERR_CUDA, "internal: the device builder produced " + std::to_string(nNodes) + " nodes for " + std::to_string(n) + " triangles"); }
    // the large-node scratch is dead
    d_bins = nullptr; d_idx[0] = d_idx[1] = nullptr; d_cbs = nullptr; d_prims = nullptr; d_scan = nullptr; scanBytes = 0;
    arena.rewind(afterTree);

    // ---- world box, triangle format, inflation (bvh_host.cpp, encode_wide)
    double wlo[3], whi[3], mag = 0.0;
    for (int k = 0; k < 3; k++) {
        unsigned long long kl = hg.wlo[k], kh = hg.whi[k];
        auto un = [](unsigned long long key) { const unsigned long long b = (key >> 63) ? (key & 0x7fffffffffffffffull) : ~key; double d; std::memcpy(&d, &b, 8); return d; };
        wlo[k] = un(kl); whi[k] = un(kh);
        hb.wlo[k] = wlo[k]; hb.whi[k] = whi[k];
        mag = std::max(mag, std::max(std::abs(wlo[k]), std::abs(whi[k])));
        mag = std::max(mag, whi[k] - wlo[k]);
    }
    if (!(mag < 1e30)) { freeAll(); return fail(ctx, SPB_ERR_UNSUPPORTED, "spb_bvh_build: scene coordinates too large for the float32 culling grid"); }
    const double inflate = std::max(mag * std::ldexp(1.0, -19), 1e-30);
    hb.inflate = inflate;
    hb.tri_format = hg.f32ok ? 0 : 1;
    hb.n_binary_nodes = nNodes;
    double rootArea;
    {
        const double dx = whi[0] - wlo[0], dy = whi[1] - wlo[1], dz = whi[2] - wlo[2];
        rootArea = std::max(dx < 0 ? 0.0 : 2.0 * (dx * dy + dy * dz + dz * dx), 1e-300);
    }

    // ---- (3) collapse
    float cTri = 0.7f;
    if (const char* e = std::getenv("SPICA_BVH_CTRI")) cTri = (float)std::atof(e);
    SALLOC(d_rows, (size_t)nNodes * sizeof(DpRow));
    SALLOC(d_flags, (size_t)nNodes * 4);
    SB(cudaMemsetAsync(d_flags, 0, (size_t)nNodes * 4, st));
    dpRowKernel<<<(nNodes + 255) / 256, 256, 0, st>>>(d_nodes, nNodes, d_parent, d_rows, d_flags, maxLeaf, cTri);
    launches++;
    lap("collapse: cost rows");
    const size_t triSize = hb.tri_format == 0 ? sizeof(TriF32) : sizeof(TriF64);
    const int wideMax = n + 16;                  // at most one wide node per ... every wide node below the root has >= 2 primitives beneath it
    const int levelMax = n + 16;
    SALLOC(d_wide, (size_t)wideMax * sizeof(WideNode));
    if (ctx->spare_tris && ctx->spare_tris_bytes >= (size_t)n * triSize && ctx->spare_tris_bytes <= 2 * (size_t)n * triSize) {
        d_tris = ctx->spare_tris; trisBytes = ctx->spare_tris_bytes; ctx->spare_tris = nullptr; ctx->spare_tris_bytes = 0;
    } else {
        SB(cudaMalloc(&d_tris, (size_t)n * triSize)); trisBytes = (size_t)n * triSize;
    }
    SALLOC(d_level[0], (size_t)levelMax * 4); SALLOC(d_level[1], (size_t)levelMax * 4);
    SALLOC(d_work, (size_t)levelMax * sizeof(WideWork));
    SALLOC(d_nInner, (size_t)(levelMax + 1) * 4); SALLOC(d_nTris, (size_t)(levelMax + 1) * 4);
    SALLOC(d_innerScan, (size_t)(levelMax + 1) * 4); SALLOC(d_triScan, (size_t)(levelMax + 1) * 4);
    {
        size_t need = 0;
        SB(cub::DeviceScan::ExclusiveSum(nullptr, need, d_nInner, d_innerScan, levelMax + 1, st));
        if (need > scanBytes) { SALLOC(d_scan, need); scanBytes = need; }
    }
    lap("emission: allocations");
    const int32_t rootNode = 0;
    SB(cudaMemcpyAsync(d_level[0], &rootNode, 4, cudaMemcpyHostToDevice, st));
    int nLevel = 1, lv = 0, depth = 0;
    uint32_t levelBase = 0, triBase = 0;
    while (nLevel > 0) {
        depth++;
        if ((int64_t)levelBase + nLevel > wideMax) { freeAll(); return fail(ctx, SPB_ERR_CUDA, "internal: wide node array overflow"); }
        gatherKernel<<<(nLevel + 127) / 128, 128, 0, st>>>(d_nodes, d_rows, d_level[lv], nLevel, maxLeaf, depth == 1, d_work, d_nInner, d_nTris);
        SB(cub::DeviceScan::ExclusiveSum(d_scan, scanBytes, d_nInner, d_innerScan, nLevel + 1, st));
        SB(cub::DeviceScan::ExclusiveSum(d_scan, scanBytes, d_nTris, d_triScan, nLevel + 1, st));
        const uint32_t nextBase = levelBase + (uint32_t)nLevel;
        emitKernel<<<(nLevel + 127) / 128, 128, 0, st>>>(d_nodes, d_order, d_verts, d_work, d_innerScan, d_triScan, nLevel, levelBase, nextBase, triBase, inflate,
                                                        hb.tri_format, d_wide, d_tris, d_level[lv ^ 1], d_g, rootArea);
        launches += 4;
        uint32_t tot[2] = {0, 0};
        SB(cudaMemcpyAsync(&tot[0], d_innerScan + nLevel, 4, cudaMemcpyDeviceToHost, st));
        SB(cudaMemcpyAsync(&tot[1], d_triScan + nLevel, 4, cudaMemcpyDeviceToHost, st));
        SB(cudaStreamSynchronize(st));
        SB(cudaGetLastError());
        levelBase = nextBase; triBase += tot[1];
        nLevel = (int)tot[0];
        lv ^= 1;
    }
    lap("emission: levels");
    if ((int64_t)triBase != n64) { freeAll(); return fail(ctx, SPB_ERR_CUDA, "internal: the device collapse emitted " + std::to_string(triBase) + " of " + std::to_string(n) + " triangles"); }
    if (depth > kStackCapacity - 2) { freeAll(); return fail(ctx, SPB_ERR_UNSUPPORTED, "spb_bvh_build: wide tree deeper than the traversal stack (" + std::to_string(depth) + ")"); }
    SB(cudaMemcpyAsync(&hg, d_g, sizeof(hg), cudaMemcpyDeviceToHost, st));
    SB(cudaStreamSynchronize(st));
    hb.max_depth = depth;
    hb.sah_cost = hg.sahCost;
    hb.n_wide = (int64_t)levelBase;
    hb.tri_bytes_device = (int64_t)((size_t)n * triSize);
    ctx->kernel_launches += launches;
    // shrink the node array to its size and hand both over
    WideNode* d_fit = nullptr;
    size_t fitBytes = (size_t)levelBase * sizeof(WideNode);
    if (ctx->spare_nodes && ctx->spare_nodes_bytes >= fitBytes && ctx->spare_nodes_bytes <= 2 * fitBytes) {
        d_fit = (WideNode*)ctx->spare_nodes; fitBytes = ctx->spare_nodes_bytes; ctx->spare_nodes = nullptr; ctx->spare_nodes_bytes = 0;
    } else {
        SB(cudaMalloc(&d_fit, fitBytes));
    }
    SB(cudaMemcpyAsync(d_fit, d_wide, (size_t)levelBase * sizeof(WideNode), cudaMemcpyDeviceToDevice, st));
    SB(cudaStreamSynchronize(st));
    ctx->d_nodes = d_fit; ctx->d_tris = d_tris;
    ctx->d_nodes_bytes = fitBytes; ctx->d_tris_bytes = trisBytes;
    keepWide = true;
    freeAll();
    lap("hand-over + frees");
    return SPB_OK;
}
#undef SALLOC
#undef SB

}  // namespace spb
