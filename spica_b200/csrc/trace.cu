// K1 trace_closest / K2 trace_any: sm_100a ray-cast kernels over the 8-wide compressed BVH.
// Replace BVHAccel::intersectBVH x2 (reference accelerators/bvh.cc:331-387) + Triangle::intersect x2
// (core/triangle.cc:98-178) for whole ray batches.
//
// Variant 1 (default) is a persistent-thread kernel: the grid is sized to the machine
// (SMs x resident CTAs), every warp pulls rays from a global cursor with one warp-aggregated
// atomic (ballot + popc + shuffle), and lanes whose ray has terminated are refilled while the
// rest of the warp keeps traversing ("dynamic fetch"), so incoherent rays do not leave the warp
// mostly idle while its longest ray finishes.  Variant 0 is the plain one-thread-per-ray kernel
// kept as the measurement baseline.
#include <algorithm>

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

__device__ __forceinline__ void storeResult(spb_hit* out, int64_t i, const RayState& r) {
    const bool hit = r.best_prim >= 0;
    uint4 v;
    v.x = __float_as_uint(hit ? (float)r.best_t : 0.0f);
    v.y = (uint32_t)r.best_prim;
    v.z = __float_as_uint(r.best_u);
    v.w = __float_as_uint(r.best_v);
    stStream(out + i, v);
}
__device__ __forceinline__ void storeResult(spb_hit_f64* out, int64_t i, const RayState& r) {
    const bool hit = r.best_prim >= 0;
    const unsigned long long t = (unsigned long long)__double_as_longlong(hit ? r.best_t : 0.0);
    const unsigned long long u = (unsigned long long)__double_as_longlong((double)r.best_u);
    const unsigned long long v = (unsigned long long)__double_as_longlong((double)r.best_v);
    stStream((char*)(out + i), make_uint4((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)u, (uint32_t)(u >> 32)));
    stStream((char*)(out + i) + 16, make_uint4((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)r.best_prim, 0u));
}
__device__ __forceinline__ void storeResult(uint8_t* out, int64_t i, const RayState& r) {
    out[i] = r.best_prim >= 0 ? 1 : 0;
}

// ---- variant 0: one thread per ray ---------------------------------------------------------------
template <int FMT, bool ANY, bool COUNT, class RayT, class OutT>
__global__ void __launch_bounds__(256) traceSimpleKernel(SceneParams sp, const RayT* __restrict__ rays, int64_t n,
                                                       OutT* __restrict__ out, unsigned long long* ctr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RayState r;
    const bool valid = loadRay(sp, rays, i, r);
    TraceCounters c = {0ull, 0ull};
    traceRay<FMT, ANY>(sp, r, valid, COUNT ? &c : nullptr);
    storeResult(out, i, r);
    if (COUNT) { atomicAdd(ctr + 1, c.nodes); atomicAdd(ctr + 2, c.tris); }
}

// ---- variant 1: persistent threads, dynamic fetch ------------------------------------------------
// REFILL_MIN: a warp goes back to the cursor once at least this many lanes are idle.
template <int FMT, bool ANY, bool COUNT, class RayT, class OutT, int REFILL_MIN, int STEPS>
__global__ void __launch_bounds__(128) tracePersistentKernel(SceneParams sp, const RayT* __restrict__ rays, int64_t n,
                                                           OutT* __restrict__ out, unsigned long long* ctr) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    Traverser<FMT, ANY> tr;
    RayState r;
    TraceCounters c = {0ull, 0ull};
    int64_t mine = -1;
    bool active = false, exhausted = false;

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
                    if (tr.finished) storeResult(out, mine, r);   // trivial miss
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
                if (tr.finished) { storeResult(out, mine, r); active = false; }
            }
        }
    }
    if (COUNT) { atomicAdd(ctr + 1, c.nodes); atomicAdd(ctr + 2, c.tris); }
}

// ---- launch ------------------------------------------------------------------------------------
template <int FMT, bool ANY, bool COUNT, class RayT, class OutT>
static int launchTyped(spb_ctx* ctx, const RayT* d_rays, int64_t n, OutT* d_out, cudaStream_t st) {
    if (n <= 0) return SPB_OK;
    SPB_CUDA(ctx, cudaMemsetAsync(ctx->d_work, 0, 4 * sizeof(unsigned long long), st));
    if (ctx->opt_variant == 0) {
        const int block = std::min(ctx->opt_block, 256);
        const int64_t grid = (n + block - 1) / block;
        traceSimpleKernel<FMT, ANY, COUNT, RayT, OutT><<<(unsigned)grid, block, 0, st>>>(ctx->sp, d_rays, n, d_out, ctx->d_work);
    } else {
        auto kern = tracePersistentKernel<FMT, ANY, COUNT, RayT, OutT, 8, 2>;
        const int block = 128;
        int perSm = ctx->opt_ctas_per_sm;
        if (perSm <= 0) {
            SPB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, block, 0));
            if (perSm < 1) perSm = 1;
        }
        int64_t grid = (int64_t)ctx->sm_count * perSm;
        const int64_t need = (n + block - 1) / block;
        if (grid > need) grid = need;
        kern<<<(unsigned)grid, block, 0, st>>>(ctx->sp, d_rays, n, d_out, ctx->d_work);
    }
    SPB_CUDA(ctx, cudaGetLastError());
    ctx->kernel_launches++;
    return SPB_OK;
}

template <bool ANY, class RayT, class OutT>
static int launch(spb_ctx* ctx, const RayT* d_rays, int64_t n, OutT* d_out, cudaStream_t st) {
    const int fmt = ctx->sp.tri_format;
    if (ctx->opt_counters) {
        return fmt == 0 ? launchTyped<0, ANY, true>(ctx, d_rays, n, d_out, st) : launchTyped<1, ANY, true>(ctx, d_rays, n, d_out, st);
    }
    return fmt == 0 ? launchTyped<0, ANY, false>(ctx, d_rays, n, d_out, st) : launchTyped<1, ANY, false>(ctx, d_rays, n, d_out, st);
}

static int readCounters(spb_ctx* ctx, int64_t n) {
    if (!ctx->opt_counters) return SPB_OK;
    unsigned long long h[4];
    SPB_CUDA(ctx, cudaMemcpy(h, ctx->d_work, sizeof(h), cudaMemcpyDeviceToHost));
    ctx->c_rays += n; ctx->c_nodes += (int64_t)h[1]; ctx->c_tris += (int64_t)h[2];
    return SPB_OK;
}

// device-resident buffers: one launch, timed with events on the launching stream
template <bool ANY, class RayT, class OutT>
static int traceDev(spb_ctx* ctx, const RayT* d_rays, int64_t n, OutT* d_out) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    if (!ctx->bvh_ready) return fail(ctx, SPB_ERR_INVALID, "trace: no acceleration structure (call spb_bvh_build or spb_bvh_import_binary)");
    if (n < 0 || (n > 0 && (!d_rays || !d_out))) return fail(ctx, SPB_ERR_INVALID, "trace: bad arguments");
    cudaSetDevice(ctx->device);
    SPB_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = launch<ANY>(ctx, d_rays, n, d_out, ctx->stream);
    if (rc) return rc;
    SPB_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    SPB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_kernel_ms = ms;
    return readCounters(ctx, n);
}

static int ensureScratch(spb_ctx* ctx, size_t inBytes, size_t outBytes) {
    if (inBytes > ctx->in_cap) {
        for (int i = 0; i < 2; i++) { if (ctx->d_in[i]) cudaFree(ctx->d_in[i]); ctx->d_in[i] = nullptr; }
        for (int i = 0; i < 2; i++) SPB_CUDA(ctx, cudaMalloc(&ctx->d_in[i], inBytes));
        ctx->in_cap = inBytes;
    }
    if (outBytes > ctx->out_cap) {
        for (int i = 0; i < 2; i++) { if (ctx->d_out[i]) cudaFree(ctx->d_out[i]); ctx->d_out[i] = nullptr; }
        for (int i = 0; i < 2; i++) SPB_CUDA(ctx, cudaMalloc(&ctx->d_out[i], outBytes));
        ctx->out_cap = outBytes;
    }
    return SPB_OK;
}

// host buffers: chunked, double buffered; H2D, kernel and D2H of neighbouring chunks overlap
template <bool ANY, class RayT, class OutT>
static int traceHost(spb_ctx* ctx, const RayT* rays, int64_t n, OutT* out) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    if (!ctx->bvh_ready) return fail(ctx, SPB_ERR_INVALID, "trace: no acceleration structure (call spb_bvh_build or spb_bvh_import_binary)");
    if (n < 0 || (n > 0 && (!rays || !out))) return fail(ctx, SPB_ERR_INVALID, "trace: bad arguments");
    if (n == 0) return SPB_OK;
    cudaSetDevice(ctx->device);
    const int64_t chunk = std::min<int64_t>(ctx->opt_chunk, n);
    int rc = ensureScratch(ctx, (size_t)chunk * sizeof(RayT), (size_t)chunk * sizeof(OutT));
    if (rc) return rc;
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t off = 0;
    for (int c = 0; off < n; c++, off += chunk) {
        const int b = c & 1;
        const int64_t m = std::min(chunk, n - off);
        SPB_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, ctx->ev_k[b], 0));        // buffer b consumed
        SPB_CUDA(ctx, cudaMemcpyAsync(ctx->d_in[b], rays + off, (size_t)m * sizeof(RayT), cudaMemcpyHostToDevice, ctx->h2d));
        SPB_CUDA(ctx, cudaEventRecord(ctx->ev_in[b], ctx->h2d));
        SPB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[b], 0));
        SPB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out[b], 0));   // previous results drained
        rc = launch<ANY>(ctx, (const RayT*)ctx->d_in[b], m, (OutT*)ctx->d_out[b], ctx->stream);
        if (rc) return rc;
        SPB_CUDA(ctx, cudaEventRecord(ctx->ev_k[b], ctx->stream));
        SPB_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h, ctx->ev_k[b], 0));
        SPB_CUDA(ctx, cudaMemcpyAsync(out + off, ctx->d_out[b], (size_t)m * sizeof(OutT), cudaMemcpyDeviceToHost, ctx->d2h));
        SPB_CUDA(ctx, cudaEventRecord(ctx->ev_out[b], ctx->d2h));
        if (ctx->opt_counters) { SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); rc = readCounters(ctx, m); if (rc) return rc; }
    }
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->d2h));
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->h2d));
    ctx->last_kernel_ms = -1.0;
    return SPB_OK;
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_trace_closest(spb_ctx* ctx, const spb_ray_f32* rays, int64_t n, spb_hit* hits) { return traceHost<false>(ctx, rays, n, hits); }
int spb_trace_closest_f64(spb_ctx* ctx, const spb_ray_f64* rays, int64_t n, spb_hit_f64* hits) { return traceHost<false>(ctx, rays, n, hits); }
int spb_trace_any(spb_ctx* ctx, const spb_ray_f32* rays, int64_t n, uint8_t* occ) { return traceHost<true>(ctx, rays, n, occ); }
int spb_trace_any_f64(spb_ctx* ctx, const spb_ray_f64* rays, int64_t n, uint8_t* occ) { return traceHost<true>(ctx, rays, n, occ); }
int spb_trace_closest_dev(spb_ctx* ctx, const spb_ray_f32* d_rays, int64_t n, spb_hit* d_hits) { return traceDev<false>(ctx, d_rays, n, d_hits); }
int spb_trace_any_dev(spb_ctx* ctx, const spb_ray_f32* d_rays, int64_t n, uint8_t* d_occ) { return traceDev<true>(ctx, d_rays, n, d_occ); }

}  // extern "C"
