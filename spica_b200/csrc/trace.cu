// K1 trace_closest / K2 trace_any: the ray-cast entry points of the C ABI over the sm_100a kernels of
// trace_kernels.cuh.  Replace BVHAccel::intersectBVH x2 (reference accelerators/bvh.cc:331-387) +
// Triangle::intersect x2 (core/triangle.cc:98-178) for whole ray batches.
//
// The default kernel is variant 5 (traceCoopPairKernel: persistent CTAs, dynamic refill, shared-memory
// stack, three node visits per warp-pooled float32 pre-test); variants 0-4 are the steps that led to it
// and stay selectable for measurement (spb_set_option "trace_variant"); the 10+ measurement variants
// are compiled only with `make EXP=1`.
#include <algorithm>

#include "context.h"
#include "trace_kernels.cuh"

namespace spb {

static int readCounters(spb_ctx* ctx, int64_t n) {
    if (!ctx->opt_counters) return SPB_OK;
    unsigned long long h[4];
    SPB_CUDA(ctx, cudaMemcpy(h, ctx->d_work, sizeof(h), cudaMemcpyDeviceToHost));
    ctx->c_rays += n; ctx->c_nodes += (int64_t)h[1]; ctx->c_tris += (int64_t)h[2];
    return SPB_OK;
}

// device-resident buffers: one launch, timed with events on the launching stream
template <bool ANY, class RayT, class Out>
static int traceDev(spb_ctx* ctx, const RayT* d_rays, int64_t n, Out d_out) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    if (!ctx->bvh_ready) return fail(ctx, SPB_ERR_INVALID, "trace: no acceleration structure (call spb_bvh_build or spb_bvh_import_binary)");
    if (n < 0 || (n > 0 && (!d_rays || !d_out.out))) return fail(ctx, SPB_ERR_INVALID, "trace: bad arguments");
    cudaSetDevice(ctx->device);
    SPB_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = launchTrace<ANY>(ctx, d_rays, n, nullptr, d_out, ctx->d_work, ctx->stream);
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
        for (int i = 0; i < spb_ctx::kPipe; i++) { if (ctx->d_in[i]) cudaFree(ctx->d_in[i]); ctx->d_in[i] = nullptr; }
        for (int i = 0; i < spb_ctx::kPipe; i++) SPB_CUDA(ctx, cudaMalloc(&ctx->d_in[i], inBytes));
        ctx->in_cap = inBytes;
    }
    if (outBytes > ctx->out_cap) {
        for (int i = 0; i < spb_ctx::kPipe; i++) { if (ctx->d_out[i]) cudaFree(ctx->d_out[i]); ctx->d_out[i] = nullptr; }
        for (int i = 0; i < spb_ctx::kPipe; i++) SPB_CUDA(ctx, cudaMalloc(&ctx->d_out[i], outBytes));
        ctx->out_cap = outBytes;
    }
    return SPB_OK;
}

// host buffers: chunked, kPipe buffers in a ring; H2D, kernel and D2H of neighbouring chunks overlap
template <bool ANY, class Out, class RayT, class OutT>
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
    const bool serial = ctx->opt_counters != 0;       // counters are read back per chunk
    // On any error the pipeline is drained before returning: copies into and out of the caller's buffers must not
    // be in flight when the caller sees the status.
    auto drain = [&]() {
        cudaStreamSynchronize(ctx->h2d);
        for (int i = 0; i < spb_ctx::kPipe; i++) cudaStreamSynchronize(ctx->kstream[i]);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->d2h);
    };
#define SPB_PIPE(call)                                                                                   \
    do {                                                                                                 \
        const cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) { drain(); cudaOk(ctx, e_, #call); return SPB_ERR_CUDA; }                 \
    } while (0)
    int64_t off = 0;
    for (int c = 0; off < n; c++, off += chunk) {
        const int b = c % spb_ctx::kPipe;
        const int64_t m = std::min(chunk, n - off);
        cudaStream_t ks = serial ? ctx->stream : ctx->kstream[b];
        unsigned long long* cursor = serial ? ctx->d_work : ctx->d_work + 4 * (b + 1);
        SPB_PIPE(cudaStreamWaitEvent(ctx->h2d, ctx->ev_k[b], 0));        // buffer b consumed
        SPB_PIPE(cudaMemcpyAsync(ctx->d_in[b], rays + off, (size_t)m * sizeof(RayT), cudaMemcpyHostToDevice, ctx->h2d));
        SPB_PIPE(cudaEventRecord(ctx->ev_in[b], ctx->h2d));
        SPB_PIPE(cudaStreamWaitEvent(ks, ctx->ev_in[b], 0));
        SPB_PIPE(cudaStreamWaitEvent(ks, ctx->ev_out[b], 0));            // previous results drained
        rc = launchTrace<ANY>(ctx, (const RayT*)ctx->d_in[b], m, nullptr, Out{(OutT*)ctx->d_out[b]}, cursor, ks);
        if (rc) { drain(); return rc; }
        SPB_PIPE(cudaEventRecord(ctx->ev_k[b], ks));
        SPB_PIPE(cudaStreamWaitEvent(ctx->d2h, ctx->ev_k[b], 0));
        SPB_PIPE(cudaMemcpyAsync(out + off, ctx->d_out[b], (size_t)m * sizeof(OutT), cudaMemcpyDeviceToHost, ctx->d2h));
        SPB_PIPE(cudaEventRecord(ctx->ev_out[b], ctx->d2h));
        if (serial) { SPB_PIPE(cudaStreamSynchronize(ctx->stream)); rc = readCounters(ctx, m); if (rc) { drain(); return rc; } }
    }
#undef SPB_PIPE
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->d2h));
    for (int i = 0; i < spb_ctx::kPipe; i++) SPB_CUDA(ctx, cudaStreamSynchronize(ctx->kstream[i]));
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->h2d));
    ctx->last_kernel_ms = -1.0;
    return SPB_OK;
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_trace_closest(spb_ctx* ctx, const spb_ray_f32* rays, int64_t n, spb_hit* hits) { return traceHost<false, HitOut>(ctx, rays, n, hits); }
int spb_trace_closest_f64(spb_ctx* ctx, const spb_ray_f64* rays, int64_t n, spb_hit_f64* hits) { return traceHost<false, HitOut64>(ctx, rays, n, hits); }
int spb_trace_any(spb_ctx* ctx, const spb_ray_f32* rays, int64_t n, uint8_t* occ) { return traceHost<true, OccOut>(ctx, rays, n, occ); }
int spb_trace_any_f64(spb_ctx* ctx, const spb_ray_f64* rays, int64_t n, uint8_t* occ) { return traceHost<true, OccOut>(ctx, rays, n, occ); }
int spb_trace_closest_dev(spb_ctx* ctx, const spb_ray_f32* d_rays, int64_t n, spb_hit* d_hits) { return traceDev<false>(ctx, d_rays, n, HitOut{d_hits}); }
int spb_trace_any_dev(spb_ctx* ctx, const spb_ray_f32* d_rays, int64_t n, uint8_t* d_occ) { return traceDev<true>(ctx, d_rays, n, OccOut{d_occ}); }

}  // extern "C"
