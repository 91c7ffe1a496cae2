// sm_100a ray-cast kernels over the 8-wide compressed BVH, shared by the ray-cast entry points
// (trace.cu) and the wavefront integrator (integrator.cu).
//
// K1 trace_closest / K2 trace_any replace BVHAccel::intersectBVH x2 (reference
// accelerators/bvh.cc:331-387) + Triangle::intersect x2 (core/triangle.cc:98-178) for whole ray
// batches.
//
// The result sink is a small functor passed by value (`Out::store(i, rayState)`), so the
// integrator fuses its own epilogues (add an unoccluded shadow contribution, add an MIS emission)
// into the traversal kernel instead of writing hit records to HBM and reading them back.
//
// The ray count may live in device memory (`n_dev`): the integrator's queues are filled by the
// previous kernel and never round-trip through the host.
//
// Variants (spb_set_option "trace_variant"; every one returns the same records, tests/test_trace_gpu.py):
//   0  one thread per ray: the measurement baseline.
//   1  persistent threads: the grid is sized to the machine (SMs x resident CTAs), every warp pulls rays
//      from a global cursor with one warp-aggregated atomic (ballot + popc + shuffle), and lanes whose ray
//      has terminated are refilled while the rest of the warp keeps traversing ("dynamic fetch"), so
//      incoherent rays do not leave the warp mostly idle while its longest ray finishes.
//   2  1 + warp-pooled float32 triangle pre-test (traceCoopKernel); the fallback for float64 triangles.
//   3  2 in visit -> select -> triangles order (traceCoopAheadKernel); the fallback for trees deeper
//      than the shared-memory stack.
//   4  3 with the traversal stack in shared memory.
//   5  4 with three node visits per pooled triangle phase (traceCoopPairKernel): the default.
//   10+ measurement variants, compiled with `make EXP=1` only (profiles/r01g_kernel_experiments.md).
#pragma once
#include <algorithm>
#include <type_traits>

#include "context.h"
#include "trace_core.h"

namespace spb {

// ---- ray / result records ----------------------------------------------------------------------
__device__ __forceinline__ uint4 ldStream(const void* p) {
    uint4 v;
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stStream(void* p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ double u2d(uint32_t lo, uint32_t hi) {
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

__device__ __forceinline__ bool loadRay(const SceneParams& sp, const spb_ray_f32* rays, int64_t i, RayState& r) {
    const uint4 a = ldStream(rays + i), b = ldStream((const char*)(rays + i) + 16);
    return rayBegin(sp, (double)__uint_as_float(a.x), (double)__uint_as_float(a.y), (double)__uint_as_float(a.z),
                    (double)__uint_as_float(a.w), (double)__uint_as_float(b.x), (double)__uint_as_float(b.y),
                    (double)__uint_as_float(b.w), r);
}
__device__ __forceinline__ bool loadRay(const SceneParams& sp, const spb_ray_f64* rays, int64_t i, RayState& r) {
    const char* p = (const char*)(rays + i);
    const uint4 a = ldStream(p), b = ldStream(p + 16), c = ldStream(p + 32), d = ldStream(p + 48);
    return rayBegin(sp, u2d(a.x, a.y), u2d(a.z, a.w), u2d(b.x, b.y), u2d(b.z, b.w), u2d(c.x, c.y), u2d(c.z, c.w),
                    u2d(d.z, d.w), r);
}

// ---- result sinks ---------------------------------------------------------------------------------
struct HitOut {       // spb_hit records (16 B)
    spb_hit* out;
    __device__ __forceinline__ void store(int64_t i, const RayState& r) const {
        const bool hit = r.best_prim >= 0;
        uint4 v;
        v.x = __float_as_uint(hit ? (float)r.best_t : 0.0f);
        v.y = (uint32_t)r.best_prim;
        v.z = __float_as_uint(r.best_u);
        v.w = __float_as_uint(r.best_v);
        stStream(out + i, v);
    }
};
struct HitOut64 {     // spb_hit_f64 records (32 B)
    spb_hit_f64* out;
    __device__ __forceinline__ void store(int64_t i, const RayState& r) const {
        const bool hit = r.best_prim >= 0;
        const unsigned long long t = (unsigned long long)__double_as_longlong(hit ? r.best_t : 0.0);
        const unsigned long long u = (unsigned long long)__double_as_longlong((double)r.best_u);
        const unsigned long long v = (unsigned long long)__double_as_longlong((double)r.best_v);
        stStream((char*)(out + i), make_uint4((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)u, (uint32_t)(u >> 32)));
        stStream((char*)(out + i) + 16, make_uint4((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)r.best_prim, 0u));
    }
};
struct OccOut {       // 1 = occluded
    uint8_t* out;
    __device__ __forceinline__ void store(int64_t i, const RayState& r) const { out[i] = r.best_prim >= 0 ? 1 : 0; }
};

// ---- variant 0: one thread per ray ---------------------------------------------------------------
template <int FMT, bool ANY, bool COUNT, class RayT, class Out>
__global__ void __launch_bounds__(256) traceSimpleKernel(SceneParams sp, const RayT* __restrict__ rays, int64_t n,
                                                       Out out, unsigned long long* ctr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RayState r;
    const bool valid = loadRay(sp, rays, i, r);
    TraceCounters c = {0ull, 0ull, 0ull};
    traceRay<FMT, ANY>(sp, r, valid, COUNT ? &c : nullptr);
    out.store(i, r);
    if (COUNT) { atomicAdd(ctr + 1, c.nodes); atomicAdd(ctr + 2, c.tris); }
}

// ---- variant 1: persistent threads, dynamic fetch ------------------------------------------------
// REFILL_MIN: a warp goes back to the cursor once at least this many lanes are idle.
// ctr[0] is the global ray cursor (zeroed before the launch).  n_dev, when non-NULL, overrides n.
template <int FMT, bool ANY, bool COUNT, class RayT, class Out, int REFILL_MIN, int STEPS>
__global__ void __launch_bounds__(128, 7) tracePersistentKernel(SceneParams sp, const RayT* __restrict__ rays, int64_t n,
                                                           const uint32_t* __restrict__ n_dev, Out out,
                                                           unsigned long long* ctr) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    if (n_dev) n = (int64_t)__ldg(n_dev);
    Traverser<FMT, ANY> tr;
    RayState r;
    TraceCounters c = {0ull, 0ull, 0ull};
    int64_t mine = -1;
    bool active = false, exhausted = (n <= 0);

    for (;;) {
        const unsigned idle = __ballot_sync(full, !active);
        const int nIdle = __popc(idle);
        if (!exhausted && (nIdle >= REFILL_MIN || nIdle == 32)) {
            const int leader = __ffs(idle) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(ctr, (unsigned long long)nIdle);
            base = __shfl_sync(full, base, leader);
            if ((int64_t)base + nIdle >= n) exhausted = true;
            if (!active) {
                const int64_t idx = (int64_t)base + __popc(idle & ((1u << lane) - 1u));
                if (idx < n) {
                    mine = idx;
                    const bool valid = loadRay(sp, rays, idx, r);
                    tr.begin(valid);
                    if (tr.finished) out.store(mine, r);   // trivial miss
                    else active = true;
                }
            }
        }
        if (!__any_sync(full, active)) {
            if (exhausted) break;
            continue;
        }
#pragma unroll 1
        for (int s = 0; s < STEPS; s++) {
            if (active) {
                tr.step(sp, r, COUNT ? &c : nullptr);
                if (tr.finished) { out.store(mine, r); active = false; }
            }
        }
    }
    if (COUNT) { atomicAdd(ctr + 1, c.nodes); atomicAdd(ctr + 2, c.tris); }
}


// ---- variant 2: persistent threads + warp-cooperative triangle pre-test --------------------------
// In variant 1 every lane walks ITS OWN candidate list after a node visit, so the warp runs
// max-over-lanes rounds of the float32 pre-test with 3-6 lanes alive (mean 0.5 candidates per lane
// and node visit; profiles/r01c: 37 % of all issue slots).  Here the candidates of the whole warp are
// pooled: an inclusive scan of the per-lane counts numbers them, lane j takes pooled candidate j
// (owner found by a 5-step shuffle search over the scan, culling ray fetched from the owner with
// seven shuffles), and the survivors go back to the owner as a bit mask through shared memory.
// One full-width round usually covers the warp.  Only the owner runs the exact double test.
template <int FMT, bool ANY, bool COUNT, class RayT, class Out, int REFILL_MIN>
__global__ void __launch_bounds__(128, 7) traceCoopKernel(SceneParams sp, const RayT* __restrict__ rays, int64_t n,
                                                          const uint32_t* __restrict__ n_dev, Out out,
                                                          unsigned long long* ctr) {
    __shared__ uint32_t s_filt[4][32];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    uint32_t* filt = s_filt[threadIdx.x >> 5];
    if (n_dev) n = (int64_t)__ldg(n_dev);
    Traverser<FMT, ANY> tr;
    RayState r;
    r.cox = r.coy = r.coz = r.fdx = r.fdy = r.fdz = r.ctmax = 0.f;
    TraceCounters c = {0ull, 0ull, 0ull};
    int64_t mine = -1;
    bool active = false, exhausted = (n <= 0);

    for (;;) {
        const unsigned idle = __ballot_sync(full, !active);
        const int nIdle = __popc(idle);
        if (!exhausted && (nIdle >= REFILL_MIN || nIdle == 32)) {
            const int leader = __ffs(idle) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(ctr, (unsigned long long)nIdle);
            base = __shfl_sync(full, base, leader);
            if ((int64_t)base + nIdle >= n) exhausted = true;
            if (!active) {
                const int64_t idx = (int64_t)base + __popc(idle & ((1u << lane) - 1u));
                if (idx < n) {
                    mine = idx;
                    const bool valid = loadRay(sp, rays, idx, r);
                    tr.begin(valid);
                    if (tr.finished) out.store(mine, r);   // trivial miss
                    else active = true;
                }
            }
        }
        if (!__any_sync(full, active)) {
            if (exhausted) break;
            continue;
        }
        U2 tg; tg.x = 0u; tg.y = 0u;
        if (active) tg = tr.nodePhase(sp, r, COUNT ? &c : nullptr);
        if (__any_sync(full, tg.y != 0u)) {
            {
                const int cnt = __popc(tg.y);
                if (COUNT) c.tris += (unsigned long long)cnt;
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(full, incl, d);
                    if (lane >= d) incl += v;
                }
                const int total = __shfl_sync(full, incl, 31);
                filt[lane] = 0u;
                __syncwarp();
                for (int base = 0; base < total; base += 32) {
                    const int j = base + lane;
                    int L = 0;                      // number of lanes whose candidates all come before j = the owner of j
#pragma unroll
                    for (int s = 16; s >= 1; s >>= 1) {
                        const int v = __shfl_sync(full, incl, L + s - 1);
                        if (v <= j) L += s;
                    }
                    const int inclL = __shfl_sync(full, incl, L);
                    const uint32_t maskL = __shfl_sync(full, tg.y, L);
                    const uint32_t baseL = __shfl_sync(full, tg.x, L);
                    CullRay cr;
                    cr.cox = __shfl_sync(full, r.cox, L); cr.coy = __shfl_sync(full, r.coy, L); cr.coz = __shfl_sync(full, r.coz, L);
                    cr.fdx = __shfl_sync(full, r.fdx, L); cr.fdy = __shfl_sync(full, r.fdy, L); cr.fdz = __shfl_sync(full, r.fdz, L);
                    cr.ctmax = __shfl_sync(full, r.ctmax, L);
                    if (j < total) {
                        int k = j - (inclL - __popc(maskL));
                        uint32_t m = maskL;
                        for (; k > 0; k--) m &= m - 1u;
                        const int bit = __ffs((int)m) - 1;
                        const TriF32* tp = sp.pre_tris + (baseL + (uint32_t)bit);
                        const U4 a = ldg4(&tp->v0[0]), b = ldg4(&tp->v1[0]), cc = ldg4(&tp->v2[0]);
                        if (triPretestMayHit(cr, sp.max_coord, a, b, cc, FMT ? sp.pre_round : 0.f)) atomicOr(&filt[L], 1u << bit);
                    }
                }
                __syncwarp();
                tg.y = filt[lane];
                if (tg.y) tr.template triPhase<false>(sp, r, tg, COUNT ? &c : nullptr);
            }
        }
        if (active) {
            tr.popPhase();
            if (tr.finished) { out.store(mine, r); active = false; }
        }
    }
    if (COUNT) { atomicAdd(ctr + 1, c.nodes); atomicAdd(ctr + 2, c.tris); }
}


// ---- variants 3, 4 (and the measurement variants 10-17): variant 2 in early-select order ----------
// Traverser::selectPhase picks the next node BEFORE the triangle phase, so its 80 bytes can be on
// their way while the warp runs the pooled pre-test and the exact tests (ncu on variant 2: 31 % of
// the stall samples sit on the node load and on the stack pop that feeds its address).
//   PF 0: early-select order only (measurement control)
//   PF 1: prefetch.global.L1 of the node's three 32-byte sectors
//   PF 2: cp.async of the node into a per-thread shared-memory slot (5 x 16 B, [piece][thread]
//         layout: conflict-free), consumed with cp.async.wait_all + LDS.128 at the next visit
__device__ __forceinline__ void prefetchL1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cpAsync16(uint32_t smem, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(g) : "memory");
}
__device__ __forceinline__ void cpAsyncWaitAll() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int PF>
__device__ __forceinline__ void fetchNodeAhead(const SceneParams& sp, uint32_t node, uint32_t smemSlot) {
    const char* np = (const char*)(sp.nodes + node);
    if (PF == 1) { prefetchL1(np); prefetchL1(np + 32); prefetchL1(np + 64); }
    if (PF == 2) {
#pragma unroll
        for (int k = 0; k < 5; k++) cpAsync16(smemSlot + (uint32_t)k * 128u * 16u, np + 16 * k);
    }
}

// SS > 0: the traversal stack lives in shared memory (SS entries per thread, [entry][thread] layout)
struct SharedStack {
    uint2* s;
    uint2* unused;                      // (same initialiser list as DeepStack)
    __device__ __forceinline__ U2 get(int i) const { const uint2 v = s[i * 128]; return U2{v.x, v.y}; }
    __device__ __forceinline__ void set(int i, U2 v) const { s[i * 128] = make_uint2(v.x, v.y); }
};

template <int FMT, int ANY_, bool COUNT, class RayT, class Out, int REFILL_MIN, int PF, bool DEFER_, int MINB, int SS>
__global__ void __launch_bounds__(128, MINB) traceCoopAheadKernel(SceneParams sp, const RayT* __restrict__ rays, int64_t n,
                                                               const uint32_t* __restrict__ n_dev, Out out,
                                                               unsigned long long* ctr) {
    constexpr bool ANY = ANY_ != 0;
    constexpr bool DEFER = DEFER_ && FMT == 0;   // the postponed exact test reads float32-exact records
    constexpr uint32_t NONE = 0xffffffffu;
    __shared__ uint32_t s_filt[4][32];
    __shared__ uint32_t s_cert[DEFER ? 4 : 1][32], s_cmin[DEFER ? 4 : 1][32];
    __shared__ float s_clo[DEFER ? 4 : 1][32];
    __shared__ uint4 s_node[PF == 2 ? 5 * 128 : 1];
    __shared__ uint2 s_stack[SS > 0 ? SS * 128 : 1];
    const SharedStack sstack = {s_stack + (SS > 0 ? threadIdx.x : 0)};
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    uint32_t* filt = s_filt[wid];
    uint32_t* cert = s_cert[DEFER ? wid : 0];
    uint32_t* cmin = s_cmin[DEFER ? wid : 0];
    float* clo = s_clo[DEFER ? wid : 0];
    const uint4* myNode = s_node + (PF == 2 ? threadIdx.x : 0);
    const uint32_t smemSlot = (uint32_t)__cvta_generic_to_shared(myNode);
    if (n_dev) n = (int64_t)__ldg(n_dev);
    Traverser<FMT, ANY> tr;
    tr.pend = NONE; tr.pend_lo = 0.f;
    RayState r;
    r.cox = r.coy = r.coz = r.fdx = r.fdy = r.fdz = r.ctmax = 0.f;
    TraceCounters c = {0ull, 0ull, 0ull};
    int64_t mine = -1;
    bool active = false, exhausted = (n <= 0);

    for (;;) {
        const unsigned idle = __ballot_sync(full, !active);      // lanes waiting for a ray (with DEFER: some still hold a postponed exact test)
        const int nIdle = __popc(idle);
        const bool refill = !exhausted && (nIdle >= REFILL_MIN || nIdle == 32);
        if (DEFER && !ANY && (refill || exhausted)) {
            // the postponed exact tests of all waiting lanes, together
            if (!active && tr.pend != NONE) {
                tr.resolvePending(sp, r, COUNT ? &c : nullptr);
                out.store(mine, r);
            }
        }
        if (refill) {
            const int leader = __ffs(idle) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(ctr, (unsigned long long)nIdle);
            base = __shfl_sync(full, base, leader);
            if ((int64_t)base + nIdle >= n) exhausted = true;
            if (!active) {
                const int64_t idx = (int64_t)base + __popc(idle & ((1u << lane) - 1u));
                if (idx < n) {
                    mine = idx;
                    const bool valid = loadRay(sp, rays, idx, r);
                    tr.begin2(valid);
                    if (tr.finished) out.store(mine, r);   // trivial miss
                    else {
                        active = true;
                        if (PF == 2 && ANY) cpAsyncWaitAll();   // an any-hit may have left a fetch in flight to this slot
                        fetchNodeAhead<PF>(sp, 0u, smemSlot);
                    }
                }
            }
        }
        if (!__any_sync(full, active)) {
            if (exhausted) break;
            continue;
        }
        U2 tg; tg.x = 0u; tg.y = 0u;
        if (active) {
            U2 g;
            if (PF == 2) {
                cpAsyncWaitAll();
                const uint4 a0 = myNode[0], a1 = myNode[128], a2 = myNode[256], a3 = myNode[384], a4 = myNode[512];
                tg = tr.visitLoaded(U4{a0.x, a0.y, a0.z, a0.w}, U4{a1.x, a1.y, a1.z, a1.w}, U4{a2.x, a2.y, a2.z, a2.w},
                                    U4{a3.x, a3.y, a3.z, a3.w}, U4{a4.x, a4.y, a4.z, a4.w}, r, &g, COUNT ? &c : nullptr, sp.f32_one);
            } else {
                tg = tr.visitPhase(sp, r, &g, COUNT ? &c : nullptr);
            }
            if (SS > 0) tr.selectPhaseOn(r, g, sstack); else tr.selectPhase(r, g);
            if (!tr.finished) fetchNodeAhead<PF>(sp, tr.cur, smemSlot);
        }
        if (__any_sync(full, tg.y != 0u)) {
            const int cnt = __popc(tg.y);
            if (COUNT) c.tris += (unsigned long long)cnt;
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(full, incl, d);
                if (lane >= d) incl += v;
            }
            const int total = __shfl_sync(full, incl, 31);
            filt[lane] = 0u;
            if (DEFER) { cert[lane] = 0u; cmin[lane] = 0x7f800000u; }
            __syncwarp();
            for (int base = 0; base < total; base += 32) {
                const int j = base + lane;
                int L = 0;                      // the owner of pooled candidate j
#pragma unroll
                for (int s = 16; s >= 1; s >>= 1) {
                    const int v = __shfl_sync(full, incl, L + s - 1);
                    if (v <= j) L += s;
                }
                const int inclL = __shfl_sync(full, incl, L);
                const uint32_t maskL = __shfl_sync(full, tg.y, L);
                const uint32_t baseL = __shfl_sync(full, tg.x, L);
                CullRay cr;
                cr.cox = __shfl_sync(full, r.cox, L); cr.coy = __shfl_sync(full, r.coy, L); cr.coz = __shfl_sync(full, r.coz, L);
                cr.fdx = __shfl_sync(full, r.fdx, L); cr.fdy = __shfl_sync(full, r.fdy, L); cr.fdz = __shfl_sync(full, r.fdz, L);
                cr.ctmax = __shfl_sync(full, r.ctmax, L);
                int cls = 0, bit = 0;
                float tlo = 0.f, thi = 0.f;
                if (j < total) {
                    int k = j - (inclL - __popc(maskL));
                    uint32_t m = maskL;
                    for (; k > 0; k--) m &= m - 1u;
                    bit = __ffs((int)m) - 1;
                    const TriF32* tp = sp.pre_tris + (baseL + (uint32_t)bit);
                    const U4 a = ldg4(&tp->v0[0]), b = ldg4(&tp->v1[0]), cc = ldg4(&tp->v2[0]);
                    if (DEFER) {
                        cls = triPretestClassify(cr, sp.max_coord, a, b, cc, &tlo, &thi);
                        if (cls == 1) atomicOr(&filt[L], 1u << bit);
                        if (cls == 2) atomicMin(&cmin[L], __float_as_uint(thi));
                    } else if (triPretestMayHit(cr, sp.max_coord, a, b, cc, FMT ? sp.pre_round : 0.f)) {
                        atomicOr(&filt[L], 1u << bit);
                    }
                }
                if (DEFER) {
                    // second pass: a certain hit stays unless it is certainly farther than the nearest certain one
                    __syncwarp();
                    if (cls == 2) {
                        const uint32_t mn = cmin[L];
                        if (tlo <= __uint_as_float(mn)) {
                            atomicOr(&cert[L], 1u << bit);
                            if (__float_as_uint(thi) == mn) clo[L] = tlo;
                        }
                    }
                }
            }
            __syncwarp();
            if (DEFER) {
                const uint32_t maybe = filt[lane], ce = cert[lane];
                if (maybe | ce) tr.mergePhase(sp, r, tg.x, maybe, ce, clo[lane], __uint_as_float(cmin[lane]), COUNT ? &c : nullptr);
            } else {
                tg.y = filt[lane];
                if (tg.y) tr.template triPhase<false>(sp, r, tg, COUNT ? &c : nullptr);
            }
        }
        if (active && tr.finished) {
            active = false;
            if (!DEFER || ANY || tr.pend == NONE) out.store(mine, r);     // else: waits, with its postponed test, for the next refill
        }
    }
    if (COUNT) { atomicAdd(ctr + 1, c.nodes); atomicAdd(ctr + 2, c.tris); }
}


// ---- variant 5: K node visits per pooled triangle phase -------------------------------------------
// Variant 4 with each lane visiting K nodes (visit, select, visit, select, ...) before the warp pools
// its triangle candidates.  With K = 3 the pool holds about 42 candidates for 32 lanes instead of 14,
// and the scan / owner search / survivor hand-back and the divergent exact-test phase run a third as
// often per node.  Testing a node's triangles up to two visits later costs about 1 % more node visits
// and a few % more candidates on the C2 rays (tests/emul mode 3) and never changes a result: the later
// nodes are only visited with a larger ctmax than they could have had.  Measured: K = 2 2244,
// K = 3 2306, K = 4 2310 Mrays/s on C2; K = 3 is also the best of the three on the Cornell renders.
// What the pooled pre-test needs of a lane (its candidate groups and its culling ray) is parked
// in shared memory, where any lane can read it; the owner keeps neither in registers.
// The same stack for trees of any depth (variant 5 on a tree deeper than SS - 2 wide levels): the first SS entries in shared
// memory, the rest in the thread's local memory (only a chain-like tree ever reaches them).
template <int SS>
struct DeepStack {
    uint2* s;
    uint2* deep;
    __device__ __forceinline__ U2 get(int i) const { const uint2 v = i < SS ? s[i * 128] : deep[i - SS]; return U2{v.x, v.y}; }
    __device__ __forceinline__ void set(int i, U2 v) const { if (i < SS) s[i * 128] = make_uint2(v.x, v.y); else deep[i - SS] = make_uint2(v.x, v.y); }
};

template <int FMT, int ANY_, bool COUNT, class RayT, class Out, int REFILL_MIN, int MINB, int SS, int K, bool DEEP = false>
__global__ void __launch_bounds__(128, MINB) traceCoopPairKernel(SceneParams sp, const RayT* __restrict__ rays, int64_t n,
                                                               const uint32_t* __restrict__ n_dev, Out out,
                                                               unsigned long long* ctr) {
    constexpr bool ANY = ANY_ != 0;
    __shared__ uint32_t s_filt[K][4][32];
    __shared__ uint2 s_tg[K][4][32];          // (first triangle record, candidate mask) of the K visits
    __shared__ float4 s_cull[2][4][32];       // (cox, coy, coz, ctmax), (fdx, fdy, fdz, -)
    __shared__ uint2 s_stack[SS * 128];
    uint2 deepEntries[DEEP ? kStackCapacity - SS : 1];
    const typename std::conditional<DEEP, DeepStack<SS>, SharedStack>::type sstack = {s_stack + threadIdx.x, deepEntries};
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    uint32_t* filt = &s_filt[0][wid][0];      // group q at filt + q * 128
    uint2* tgs = &s_tg[0][wid][0];            // group q at tgs + q * 128
    float4* cullA = s_cull[0][wid];
    float4* cullB = s_cull[1][wid];
    if (n_dev) n = (int64_t)__ldg(n_dev);
    Traverser<FMT, ANY> tr;
    RayState r;
    r.ctmax = 0.f;
    TraceCounters c = {0ull, 0ull, 0ull};
    int64_t mine = -1;
    bool active = false, exhausted = (n <= 0);

    for (;;) {
        const unsigned idle = __ballot_sync(full, !active);
        const int nIdle = __popc(idle);
        if (!exhausted && (nIdle >= REFILL_MIN || nIdle == 32)) {
            const int leader = __ffs(idle) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(ctr, (unsigned long long)nIdle);
            base = __shfl_sync(full, base, leader);
            if ((int64_t)base + nIdle >= n) exhausted = true;
            if (!active) {
                const int64_t idx = (int64_t)base + __popc(idle & ((1u << lane) - 1u));
                if (idx < n) {
                    mine = idx;
                    const bool valid = loadRay(sp, rays, idx, r);
                    tr.begin2(valid);
                    if (tr.finished) out.store(mine, r);   // trivial miss
                    else {
                        active = true;
                        cullA[lane] = make_float4(r.cox, r.coy, r.coz, r.ctmax);
                        cullB[lane] = make_float4(r.fdx, r.fdy, r.fdz, 0.f);
                    }
                }
            }
        }
        if (!__any_sync(full, active)) {
            if (exhausted) break;
            continue;
        }
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < K; q++) {
            U2 tg; tg.x = 0u; tg.y = 0u;
            if (active && !tr.finished) {
                U2 g;
                tg = tr.visitPhase(sp, r, &g, COUNT ? &c : nullptr);
                tr.selectPhaseOn(r, g, sstack);
            }
            tgs[q * 128 + lane] = make_uint2(tg.x, tg.y);
            cnt += __popc(tg.y);
        }
        if (active) reinterpret_cast<float*>(&cullA[lane])[3] = r.ctmax;
        if (__any_sync(full, cnt != 0)) {
            if (COUNT) c.tris += (unsigned long long)cnt;
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(full, incl, d);
                if (lane >= d) incl += v;
            }
            const int total = __shfl_sync(full, incl, 31);
#pragma unroll
            for (int q = 0; q < K; q++) filt[q * 128 + lane] = 0u;
            __syncwarp();
            for (int base = 0; base < total; base += 32) {
                const int j = base + lane;
                int L = 0;                      // the owner of pooled candidate j
#pragma unroll
                for (int s = 16; s >= 1; s >>= 1) {
                    const int v = __shfl_sync(full, incl, L + s - 1);
                    if (v <= j) L += s;
                }
                const int inclL = __shfl_sync(full, incl, L);
                const int cntL = __shfl_sync(full, cnt, L);
                if (j < total) {
                    int k = j - (inclL - cntL);            // index among L's candidates, groups in visit order
                    uint2 g = tgs[L];
                    int which = 0;
#pragma unroll
                    for (int q = 1; q < K; q++) {
                        const int nq = __popc(g.y);
                        if (which == q - 1 && k >= nq) { k -= nq; which = q; g = tgs[q * 128 + L]; }
                    }
                    uint32_t m = g.y;
                    for (; k > 0; k--) m &= m - 1u;
                    const int bit = __ffs((int)m) - 1;
                    const float4 c0 = cullA[L], c1 = cullB[L];
                    const CullRay cr = {c0.x, c0.y, c0.z, c1.x, c1.y, c1.z, c0.w};
                    const TriF32* tp = sp.pre_tris + (g.x + (uint32_t)bit);
                    const U4 ta = ldg4(&tp->v0[0]), tb = ldg4(&tp->v1[0]), tc = ldg4(&tp->v2[0]);
                    if (triPretestMayHit(cr, sp.max_coord, ta, tb, tc, FMT ? sp.pre_round : 0.f)) atomicOr(&filt[which * 128 + L], 1u << bit);
                }
            }
            __syncwarp();
            // the exact tests of all groups in ONE loop: the lanes with a survivor in any group go through
            // the double-precision code together (it is the most divergent part of the kernel)
            uint32_t sv[K];
            uint32_t any = 0u;
#pragma unroll
            for (int q = 0; q < K; q++) { sv[q] = filt[q * 128 + lane]; any |= sv[q]; }
            if (any) {
                do {
                    uint32_t index = 0u;
                    bool taken = false;
#pragma unroll
                    for (int q = 0; q < K; q++) {
                        if (!taken && sv[q]) {
                            index = tgs[q * 128 + lane].x + (uint32_t)(__ffs((int)sv[q]) - 1);
                            sv[q] &= sv[q] - 1u; taken = true;
                        }
                    }
                    tr.acceptExact(sp, r, index, COUNT ? &c : nullptr);
                    any = 0u;
#pragma unroll
                    for (int q = 0; q < K; q++) any |= sv[q];
                } while (any && !(ANY && r.best_prim >= 0));
            }
        }
        if (active && tr.finished) { out.store(mine, r); active = false; }
    }
    if (COUNT) { atomicAdd(ctr + 1, c.nodes); atomicAdd(ctr + 2, c.tris); }
}


// ---- launch ------------------------------------------------------------------------------------
// `cursor` points at 4 x u64: [0] the ray cursor (zeroed here), [1] node visits, [2] triangle tests.
template <int FMT, bool ANY, bool COUNT, class RayT, class Out>
static int launchTraceTyped(spb_ctx* ctx, const RayT* d_rays, int64_t n, const uint32_t* n_dev, Out out,
                            unsigned long long* cursor, cudaStream_t st, bool cursorIsClear) {
    if (n <= 0) return SPB_OK;     // n is the capacity bound when n_dev is given
    if (COUNT) { SPB_CUDA(ctx, cudaMemsetAsync(cursor, 0, 4 * sizeof(unsigned long long), st)); }
    else if (!cursorIsClear) { SPB_CUDA(ctx, cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st)); }
    if (ctx->opt_variant == 0 && !n_dev) {
        const int block = std::min(ctx->opt_block, 256);
        const int64_t grid = (n + block - 1) / block;
        traceSimpleKernel<FMT, ANY, COUNT, RayT, Out><<<(unsigned)grid, block, 0, st>>>(ctx->sp, d_rays, n, out, cursor);
    } else {
        int coop = ctx->opt_variant >= 2 ? 1 : 0;
        void (*kern)(SceneParams, const RayT*, int64_t, const uint32_t*, Out, unsigned long long*) =
            coop ? traceCoopKernel<FMT, ANY, COUNT, RayT, Out, 8> : tracePersistentKernel<FMT, ANY, COUNT, RayT, Out, 8, 2>;
        {
            const int v = ctx->opt_variant;
            if (v == 3) { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 8, 0, false, 7, 0>; coop = 2; }
            // 4: the stack in shared memory, when the tree is shallow enough for its 16 entries
            if (v == 4) {
                if (ctx->sp.max_depth <= 14) { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 8, 0, false, 7, 16>; coop = 11; }
                else { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 8, 0, false, 7, 0>; coop = 2; }
            }
            // 5: three node visits per pooled triangle phase, 6 CTAs per SM (80 registers); deeper trees: the stack continues in local memory
            if (v == 5) {
                if (ctx->sp.max_depth <= 14) { kern = traceCoopPairKernel<FMT, ANY, COUNT, RayT, Out, 8, 6, 16, 3>; coop = 15; }
                else { kern = traceCoopPairKernel<FMT, ANY, COUNT, RayT, Out, 8, 6, 16, 3, true>; coop = 20; }
            }
#ifdef SPB_EXPERIMENTAL_VARIANTS      // measurement variants (trace.cu only; profiles/r01g_kernel_experiments.md)
            if (v == 10) { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 8, 1, false, 7, 0>; coop = 3; }
            if (v == 11) { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 8, 2, false, 7, 0>; coop = 4; }
            if (v == 12) { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 8, 0, true, 7, 0>; coop = 5; }
            if (v == 13) { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 8, 0, false, 8, 0>; coop = 6; }
            if (v == 14) { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 8, 0, false, 6, 0>; coop = 7; }
            if (v == 15 && ctx->sp.max_depth <= 14) { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 4, 0, false, 7, 16>; coop = 8; }
            if (v == 16 && ctx->sp.max_depth <= 14) { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 6, 0, false, 7, 16>; coop = 9; }
            if (v == 17 && ctx->sp.max_depth <= 14) { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 12, 0, false, 7, 16>; coop = 10; }
            if (v == 25 && ctx->sp.max_depth <= 14) { kern = traceCoopPairKernel<FMT, ANY, COUNT, RayT, Out, 8, 5, 16, 3>; coop = 16; }
            if (v == 21 && ctx->sp.max_depth <= 14) { kern = traceCoopPairKernel<FMT, ANY, COUNT, RayT, Out, 8, 7, 16, 3>; coop = 19; }
            if (v == 22 && ctx->sp.max_depth <= 14) { kern = traceCoopPairKernel<FMT, ANY, COUNT, RayT, Out, 8, 6, 16, 2>; coop = 17; }
            if (v == 23 && ctx->sp.max_depth <= 14) { kern = traceCoopPairKernel<FMT, ANY, COUNT, RayT, Out, 8, 6, 16, 4>; coop = 18; }
            if (v == 18 && ctx->sp.max_depth <= 10) { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 8, 0, false, 7, 12>; coop = 12; }
            if (v == 19 && ctx->sp.max_depth <= 14) { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 8, 0, false, 8, 16>; coop = 13; }
            if (v == 20 && ctx->sp.max_depth <= 14) { kern = traceCoopAheadKernel<FMT, ANY, COUNT, RayT, Out, 8, 0, false, 6, 16>; coop = 14; }
#endif
        }
        const int block = 128;
        int perSm = ctx->opt_ctas_per_sm;
        if (perSm <= 0) {
            int& cached = ctx->occupancy[(const void*)kern];
            if (cached <= 0) {
                SPB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached, kern, block, 0));
                if (cached < 1) cached = 1;
            }
            perSm = cached;
        }
        int64_t grid = (int64_t)ctx->sm_count * perSm;
        const int64_t need = (n + block - 1) / block;
        if (grid > need) grid = need;
        kern<<<(unsigned)grid, block, 0, st>>>(ctx->sp, d_rays, n, n_dev, out, cursor);
    }
    SPB_CUDA(ctx, cudaGetLastError());
    ctx->kernel_launches++;
    return SPB_OK;
}

// fromLoop: a launch of the integrator's streaming loop -- its control kernel has already cleared the cursor, and the
// traversal counters (a ray-cast diagnostic) are never collected there.
template <bool ANY, class RayT, class Out>
static int launchTrace(spb_ctx* ctx, const RayT* d_rays, int64_t n, const uint32_t* n_dev, Out out,
                       unsigned long long* cursor, cudaStream_t st, bool fromLoop = false) {
    const int fmt = ctx->sp.tri_format;
    if (ctx->opt_counters && !fromLoop) {
        return fmt == 0 ? launchTraceTyped<0, ANY, true>(ctx, d_rays, n, n_dev, out, cursor, st, false)
                        : launchTraceTyped<1, ANY, true>(ctx, d_rays, n, n_dev, out, cursor, st, false);
    }
    return fmt == 0 ? launchTraceTyped<0, ANY, false>(ctx, d_rays, n, n_dev, out, cursor, st, fromLoop)
                    : launchTraceTyped<1, ANY, false>(ctx, d_rays, n, n_dev, out, cursor, st, fromLoop);
}

}  // namespace spb
