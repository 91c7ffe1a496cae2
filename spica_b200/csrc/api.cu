// C ABI: context, geometry upload, acceleration-structure build / import, raw device memory.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "context.h"

namespace spb {

static std::mutex g_errMutex;
static std::string g_lastError;

void setGlobalError(const std::string& msg) {
    std::lock_guard<std::mutex> lk(g_errMutex);
    g_lastError = msg;
}

int fail(spb_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    setGlobalError(msg);
    return code;
}

bool cudaOk(spb_ctx* ctx, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    fail(ctx, SPB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return false;
}

static std::mutex g_poolMutex;
static cudaMemPool_t g_pools[64] = {};
static cudaMemPool_t scratchPool(int device) {
    std::lock_guard<std::mutex> lk(g_poolMutex);
    if (device < 0 || device >= 64) return nullptr;
    if (!g_pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        g_pools[device] = pool;
    }
    return g_pools[device];
}
cudaError_t scratchAlloc(spb_ctx* ctx, void** p, size_t bytes, cudaStream_t st) {
    *p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaMemPool_t pool = scratchPool(ctx->device);
    if (!pool) return cudaMalloc(p, bytes);                       // (no pool support: plain allocations; scratchFree copes)
    return cudaMallocFromPoolAsync(p, bytes, pool, st);
}
void scratchFree(void* p, cudaStream_t st) {
    if (!p) return;
    if (cudaFreeAsync(p, st) != cudaSuccess) { cudaGetLastError(); cudaFree(p); }
}
void scratchRelease(int device) {
    std::lock_guard<std::mutex> lk(g_poolMutex);
    if (device >= 0 && device < 64 && g_pools[device]) { cudaDeviceSynchronize(); cudaMemPoolTrimTo(g_pools[device], 0); }
}

static void freeBvhDevice(spb_ctx* ctx) {
    if (ctx->d_nodes) cudaFree(ctx->d_nodes);
    if (ctx->d_tris) cudaFree(ctx->d_tris);
    if (ctx->d_pre_tris) cudaFree(ctx->d_pre_tris);
    ctx->d_nodes = ctx->d_tris = ctx->d_pre_tris = nullptr;
    ctx->d_nodes_bytes = ctx->d_tris_bytes = 0;
    ctx->bvh_ready = false;
    // nothing may keep pointing at the freed tree: the kernels' parameter block goes back to "no geometry" and a
    // render that was begun on the old tree needs a new spb_render_begin
    std::memset(&ctx->sp, 0, sizeof(ctx->sp));
    ctx->sp.empty = 1; ctx->sp.f32_one = 0x3f800000u;
    renderSceneChanged(ctx);
}

// TriF64 records -> the float32 records the pre-test reads (vertices rounded to nearest, same id and tie rank)
__global__ void roundTrisKernel(const TriF64* __restrict__ in, TriF32* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TriF64 t = in[i];
    TriF32 o;
    for (int k = 0; k < 3; k++) { o.v0[k] = (float)t.v[k]; o.v1[k] = (float)t.v[3 + k]; o.v2[k] = (float)t.v[6 + k]; }
    o.id = t.id; o.rank = t.rank; o.pad = 0;
    out[i] = o;
}

// fills the kernels' parameter block from ctx->bvh's statistics and the device arrays
static int setSceneParams(spb_ctx* ctx) {
    const HostBVH& b = ctx->bvh;
    SceneParams& sp = ctx->sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.f32_one = 0x3f800000u;
    sp.empty = (b.n_tris == 0 || b.n_wide == 0) ? 1 : 0;
    sp.n_tris = (int32_t)b.n_tris;
    sp.tri_format = b.tri_format;
    sp.max_depth = b.max_depth;
    sp.inflate = (float)b.inflate;
    for (int k = 0; k < 3; k++) {
        sp.wlo[k] = b.wlo[k] - 2.0 * b.inflate;
        sp.whi[k] = b.whi[k] + 2.0 * b.inflate;
    }
    double m = 0.0;
    for (int k = 0; k < 3; k++) m = std::max(m, std::max(std::abs(sp.wlo[k]), std::abs(sp.whi[k])));
    sp.max_coord = (float)(m * 1.0000002);
    sp.nodes = (const WideNode*)ctx->d_nodes;
    sp.tris = ctx->d_tris;
    sp.pre_tris = (const TriF32*)ctx->d_tris;
    sp.pre_round = 0.f;
    if (!sp.empty && b.tri_format == 1) {
        if (ctx->d_pre_tris) { cudaFree(ctx->d_pre_tris); ctx->d_pre_tris = nullptr; }
        SPB_CUDA(ctx, cudaMalloc(&ctx->d_pre_tris, (size_t)b.n_tris * sizeof(TriF32)));
        roundTrisKernel<<<(unsigned)((b.n_tris + 255) / 256), 256, 0, ctx->stream>>>((const TriF64*)ctx->d_tris, (TriF32*)ctx->d_pre_tris, (int64_t)b.n_tris);
        SPB_CUDA(ctx, cudaGetLastError());
        SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        sp.pre_tris = (const TriF32*)ctx->d_pre_tris;
        sp.pre_round = sp.max_coord * 1.1920929e-07f;     // 2^-23 max_coord
    }
    ctx->bvh_ready = true;
    return SPB_OK;
}

// host-built tree -> device
static int uploadBvh(spb_ctx* ctx) {
    freeBvhDevice(ctx);
    const HostBVH& b = ctx->bvh;
    if (b.n_tris > 0 && !b.nodes.empty()) {
        SPB_CUDA(ctx, cudaMalloc(&ctx->d_nodes, b.nodes.size() * sizeof(WideNode)));
        SPB_CUDA(ctx, cudaMalloc(&ctx->d_tris, b.tris.size()));
        ctx->d_nodes_bytes = b.nodes.size() * sizeof(WideNode); ctx->d_tris_bytes = b.tris.size();
        SPB_CUDA(ctx, cudaMemcpyAsync(ctx->d_nodes, b.nodes.data(), b.nodes.size() * sizeof(WideNode), cudaMemcpyHostToDevice, ctx->stream));
        SPB_CUDA(ctx, cudaMemcpyAsync(ctx->d_tris, b.tris.data(), b.tris.size(), cudaMemcpyHostToDevice, ctx->stream));
        SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return setSceneParams(ctx);
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_version(void) { return SPB_VERSION; }

const char* spb_last_error(const spb_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> lk(g_errMutex);
    static thread_local std::string copy;
    copy = g_lastError;
    return copy.c_str();
}

int spb_ctx_create(int device, spb_ctx** out) {
    if (!out) return fail(nullptr, SPB_ERR_INVALID, "spb_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, SPB_ERR_NO_DEVICE, std::string("no CUDA device (spica_b200 has no CPU path): ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, SPB_ERR_INVALID, "spb_ctx_create: device index out of range");
    spb_ctx* ctx = new spb_ctx();
    ctx->device = device;
    auto bail = [&](cudaError_t ce, const char* what) {
        fail(nullptr, SPB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(ce));
        delete ctx;
        return SPB_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->h2d, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    for (int i = 0; i < spb_ctx::kPipe; i++)
        if ((e = cudaStreamCreateWithFlags(&ctx->kstream[i], cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1);
    for (int i = 0; i < spb_ctx::kPipe; i++) {
        cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming);
    }
    if ((e = cudaMalloc(&ctx->d_work, 4 * (spb_ctx::kPipe + 1) * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc");
    cudaMemset(ctx->d_work, 0, 4 * (spb_ctx::kPipe + 1) * sizeof(unsigned long long));
    std::memset(&ctx->sp, 0, sizeof(ctx->sp));
    ctx->sp.empty = 1; ctx->sp.f32_one = 0x3f800000u;
    renderStateEnsure(ctx);
    *out = ctx;
    return SPB_OK;
}

void spb_ctx_destroy(spb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    renderStateDestroy(ctx);
    freeBvhDevice(ctx);
    ctx->build_arena.release();
    for (int i = 0; i < spb_ctx::kPipe; i++) {
        if (ctx->d_in[i]) cudaFree(ctx->d_in[i]);
        if (ctx->d_out[i]) cudaFree(ctx->d_out[i]);
        cudaEventDestroy(ctx->ev_in[i]); cudaEventDestroy(ctx->ev_k[i]); cudaEventDestroy(ctx->ev_out[i]);
    }
    if (ctx->d_work) cudaFree(ctx->d_work);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream); cudaStreamDestroy(ctx->h2d); cudaStreamDestroy(ctx->d2h);
    for (int i = 0; i < spb_ctx::kPipe; i++) cudaStreamDestroy(ctx->kstream[i]);
    delete ctx;
}

int spb_scene_set_triangles(spb_ctx* ctx, const double* verts, const float* normals, const float* uvs,
                            const int32_t* material_id, const int32_t* light_id, int64_t n) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    if (n < 0 || (n > 0 && !verts)) return fail(ctx, SPB_ERR_INVALID, "spb_scene_set_triangles: bad arguments");
    if (n > 0x7fffffff / 2) return fail(ctx, SPB_ERR_UNSUPPORTED, "more than 2^30 triangles");
    cudaSetDevice(ctx->device);
    ctx->n_tris = n;
    auto geo = std::make_shared<spb_ctx::Geometry>();
    geo->verts.assign(verts, verts + n * 9);
    if (normals) geo->normals.assign(normals, normals + n * 9);
    if (uvs) geo->uvs.assign(uvs, uvs + n * 6);
    ctx->geo = geo;
    if (material_id) ctx->material_id.assign(material_id, material_id + n); else ctx->material_id.assign((size_t)n, 0);
    if (light_id) ctx->light_id.assign(light_id, light_id + n); else ctx->light_id.assign((size_t)n, -1);
    freeBvhDevice(ctx);
    ctx->bin = BinaryBVH();
    renderSceneChanged(ctx);
    return SPB_OK;
}

int spb_scene_set_triangle_attributes(spb_ctx* ctx, const int32_t* material_id, const int32_t* light_id, int64_t n) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    if (n != ctx->n_tris) return fail(ctx, SPB_ERR_INVALID, "spb_scene_set_triangle_attributes: triangle count differs from spb_scene_set_triangles");
    if (material_id) ctx->material_id.assign(material_id, material_id + n);
    if (light_id) ctx->light_id.assign(light_id, light_id + n);
    renderSceneChanged(ctx);
    return SPB_OK;
}

int spb_bvh_build(spb_ctx* ctx, const spb_build_opts* opts) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    cudaSetDevice(ctx->device);
    spb_build_opts o = {SPB_BUILDER_DEVICE_SAH, 0, 32, 0};
    if (opts) o = *opts;
    if (o.max_leaf_tris <= 0) o.max_leaf_tris = 3;     // measured best with the cost-optimal collapse (profiles/r01g_kernel_experiments.md)
    if (o.sah_bins <= 0) o.sah_bins = 32;
    if (o.builder != SPB_BUILDER_DEVICE_SAH && o.builder != SPB_BUILDER_LBVH && o.builder != SPB_BUILDER_HOST_SAH)
        return fail(ctx, SPB_ERR_UNSUPPORTED, "spb_bvh_build: builder must be 0 (device binned SAH), 1 (device LBVH) or 2 (host binned SAH)");
    // the device builder sweeps 32 bins (one per lane) and always collapses cost-optimally; other settings are the host builder's
    bool greedy = false;
    if (const char* e = std::getenv("SPICA_BVH_COLLAPSE")) greedy = std::atoi(e) == 0;
    if (o.builder == SPB_BUILDER_DEVICE_SAH && (o.sah_bins != 32 || greedy)) o.builder = SPB_BUILDER_HOST_SAH;
    const auto t0 = std::chrono::steady_clock::now();
    if (o.builder == SPB_BUILDER_DEVICE_SAH) {
        // a rebuild (an animated mesh, say) keeps the previous tree's two arrays for the builder to fill again when they are
        // large enough: with the work arrays coming from the scratch pool it then makes no driver allocation at all
        ctx->spare_nodes = ctx->d_nodes; ctx->spare_nodes_bytes = ctx->d_nodes_bytes;
        ctx->spare_tris = ctx->d_tris; ctx->spare_tris_bytes = ctx->d_tris_bytes;
        ctx->d_nodes = ctx->d_tris = nullptr;
        freeBvhDevice(ctx);
        ctx->bin = BinaryBVH();
        const int rc = buildSahDevice(ctx, o.max_leaf_tris);
        if (ctx->spare_nodes) cudaFree(ctx->spare_nodes);
        if (ctx->spare_tris) cudaFree(ctx->spare_tris);
        ctx->spare_nodes = ctx->spare_tris = nullptr; ctx->spare_nodes_bytes = ctx->spare_tris_bytes = 0;
        if (rc) { freeBvhDevice(ctx); return rc; }
        ctx->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        ctx->builder_used = SPB_BUILDER_DEVICE_SAH;
        return setSceneParams(ctx);
    }
    if (o.builder == SPB_BUILDER_LBVH) {
        const int rc = buildLbvhDevice(ctx, &ctx->bin);
        if (rc) return rc;
    } else {
        build_binary_sah(ctx->geo->verts.data(), ctx->n_tris, o.sah_bins, &ctx->bin);
    }
    std::string err;
    if (!encode_wide(ctx->bin, ctx->geo->verts.data(), ctx->n_tris, o.max_leaf_tris, &ctx->bvh, &err))
        return fail(ctx, SPB_ERR_UNSUPPORTED, "spb_bvh_build: " + err);
    ctx->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    ctx->builder_used = o.builder;
    return uploadBvh(ctx);
}

int spb_bvh_import_binary(spb_ctx* ctx, const spb_import_node* nodes, int64_t n_nodes, int32_t root) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    if (!nodes) return fail(ctx, SPB_ERR_INVALID, "spb_bvh_import_binary: nodes is NULL");
    cudaSetDevice(ctx->device);
    const auto t0 = std::chrono::steady_clock::now();
    std::string err;
    if (!import_binary(nodes, n_nodes, root, ctx->geo->verts.data(), ctx->n_tris, &ctx->bin, &err))
        return fail(ctx, SPB_ERR_INVALID, "spb_bvh_import_binary: " + err);
    if (!encode_wide(ctx->bin, ctx->geo->verts.data(), ctx->n_tris, 3, &ctx->bvh, &err))
        return fail(ctx, SPB_ERR_UNSUPPORTED, "spb_bvh_import_binary: " + err);
    ctx->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    ctx->builder_used = SPB_BUILDER_HOST_SAH;
    return uploadBvh(ctx);
}

// ---- sharing one built tree: export / adopt (any process), clone (same process, device to device) ------------------------
int spb_bvh_export(spb_ctx* ctx, void* nodes, size_t node_capacity, void* tris, size_t tri_capacity) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    if (!ctx->bvh_ready) return fail(ctx, SPB_ERR_INVALID, "spb_bvh_export: no acceleration structure built");
    const size_t nb = (size_t)ctx->bvh.n_wide * sizeof(WideNode), tb = (size_t)ctx->bvh.tri_bytes_device;
    if ((nb && !nodes) || (tb && !tris) || node_capacity < nb || tri_capacity < tb)
        return fail(ctx, SPB_ERR_INVALID, "spb_bvh_export: buffers smaller than spb_bvh_get_stats' node_bytes / tri_bytes");
    cudaSetDevice(ctx->device);
    if (nb) SPB_CUDA(ctx, cudaMemcpyAsync(nodes, ctx->d_nodes, nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (tb) SPB_CUDA(ctx, cudaMemcpyAsync(tris, ctx->d_tris, tb, cudaMemcpyDeviceToHost, ctx->stream));
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPB_OK;
}

static void adoptStats(spb_ctx* ctx, const spb_bvh_stats& st) {
    HostBVH& b = ctx->bvh;
    b.nodes.clear(); b.tris.clear();
    b.n_tris = st.n_tris; b.n_wide = st.n_wide_nodes; b.tri_bytes_device = st.tri_bytes; b.n_binary_nodes = st.n_binary_nodes;
    b.tri_format = st.tri_format; b.max_depth = st.max_depth; b.sah_cost = st.sah_cost; b.inflate = st.inflate;
    for (int k = 0; k < 3; k++) { b.wlo[k] = st.world_lo[k]; b.whi[k] = st.world_hi[k]; }
}

int spb_bvh_import_wide(spb_ctx* ctx, const spb_bvh_stats* st, const void* nodes, const void* tris) {
    if (!ctx || !st) return fail(ctx, SPB_ERR_INVALID, "spb_bvh_import_wide: NULL argument");
    if (st->n_tris != ctx->n_tris) return fail(ctx, SPB_ERR_INVALID, "spb_bvh_import_wide: the tree was built for a different number of triangles than spb_scene_set_triangles gave this context");
    const size_t triSize = st->tri_format == 0 ? sizeof(TriF32) : sizeof(TriF64);
    if (st->n_wide_nodes < 0 || st->node_bytes != st->n_wide_nodes * (int64_t)sizeof(WideNode) || st->tri_bytes != st->n_tris * (int64_t)triSize ||
        (st->tri_format != 0 && st->tri_format != 1) || st->max_depth < 0 || st->max_depth > kStackCapacity - 2 || !(st->inflate > 0.0) ||
        (st->n_tris > 0 && (!nodes || !tris || st->n_wide_nodes < 1)))
        return fail(ctx, SPB_ERR_INVALID, "spb_bvh_import_wide: inconsistent description (pass the exporter's spb_bvh_get_stats unchanged)");
    cudaSetDevice(ctx->device);
    const auto t0 = std::chrono::steady_clock::now();
    freeBvhDevice(ctx);
    ctx->bin = BinaryBVH();
    adoptStats(ctx, *st);
    if (st->n_tris > 0) {
        SPB_CUDA(ctx, cudaMalloc(&ctx->d_nodes, (size_t)st->node_bytes));
        SPB_CUDA(ctx, cudaMalloc(&ctx->d_tris, (size_t)st->tri_bytes));
        ctx->d_nodes_bytes = (size_t)st->node_bytes; ctx->d_tris_bytes = (size_t)st->tri_bytes;
        SPB_CUDA(ctx, cudaMemcpyAsync(ctx->d_nodes, nodes, (size_t)st->node_bytes, cudaMemcpyHostToDevice, ctx->stream));
        SPB_CUDA(ctx, cudaMemcpyAsync(ctx->d_tris, tris, (size_t)st->tri_bytes, cudaMemcpyHostToDevice, ctx->stream));
        SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    ctx->builder_used = -1;
    return setSceneParams(ctx);
}

int spb_ctx_clone_scene(spb_ctx* dst, spb_ctx* src) {
    if (!dst || !src || dst == src) return fail(dst, SPB_ERR_INVALID, "spb_ctx_clone_scene: two different contexts are needed");
    if (!src->bvh_ready) return fail(dst, SPB_ERR_INVALID, "spb_ctx_clone_scene: the source context has no acceleration structure");
    const auto t0 = std::chrono::steady_clock::now();
    // the two GPUs into each other's address space when they are peers (NVLink / NVSwitch): the copies below then go device
    // to device instead of through the host, and spb_film_reduce_peers finds the mapping in place (enabling it costs
    // milliseconds, which do not belong into a frame)
    if (dst->device != src->device) {
        const int pair[2][2] = {{src->device, dst->device}, {dst->device, src->device}};
        for (const auto& pr : pair) {
            int can = 0;
            cudaSetDevice(pr[0]);
            if (cudaDeviceCanAccessPeer(&can, pr[0], pr[1]) == cudaSuccess && can) cudaDeviceEnablePeerAccess(pr[1], 0);
            cudaGetLastError();          // "already enabled" is fine
        }
    }
    cudaSetDevice(dst->device);
    dst->n_tris = src->n_tris;
    dst->geo = src->geo;                 // shared, not copied: 72 B per triangle stay where they are
    dst->material_id = src->material_id; dst->light_id = src->light_id;
    freeBvhDevice(dst);
    dst->bin = BinaryBVH();
    spb_bvh_stats st;
    int rc = spb_bvh_get_stats(src, &st);
    if (rc) return fail(dst, rc, "spb_ctx_clone_scene: cannot describe the source tree");
    adoptStats(dst, st);
    if (st.n_tris > 0) {
        SPB_CUDA(dst, cudaMalloc(&dst->d_nodes, (size_t)st.node_bytes));
        SPB_CUDA(dst, cudaMalloc(&dst->d_tris, (size_t)st.tri_bytes));
        dst->d_nodes_bytes = (size_t)st.node_bytes; dst->d_tris_bytes = (size_t)st.tri_bytes;
        // device to device: over NVLink when the GPUs are peers (mapped above), through the host otherwise (the runtime picks)
        SPB_CUDA(dst, cudaMemcpyPeerAsync(dst->d_nodes, dst->device, src->d_nodes, src->device, (size_t)st.node_bytes, dst->stream));
        SPB_CUDA(dst, cudaMemcpyPeerAsync(dst->d_tris, dst->device, src->d_tris, src->device, (size_t)st.tri_bytes, dst->stream));
        SPB_CUDA(dst, cudaStreamSynchronize(dst->stream));
    }
    dst->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    dst->builder_used = -1;
    { const int rc = setSceneParams(dst); if (rc) return rc; }
    renderSceneClone(dst, src);
    return SPB_OK;
}

int spb_bvh_get_stats(const spb_ctx* ctx, spb_bvh_stats* out) {
    if (!ctx || !out) return fail(nullptr, SPB_ERR_INVALID, "spb_bvh_get_stats: NULL argument");
    if (!ctx->bvh_ready) return fail(const_cast<spb_ctx*>(ctx), SPB_ERR_INVALID, "spb_bvh_get_stats: no acceleration structure built");
    const HostBVH& b = ctx->bvh;
    std::memset(out, 0, sizeof(*out));
    out->n_tris = b.n_tris;
    out->n_wide_nodes = b.n_wide;
    out->n_binary_nodes = b.n_binary_nodes;
    out->node_bytes = b.n_wide * (int64_t)sizeof(WideNode);
    out->tri_bytes = b.tri_bytes_device;
    out->sah_cost = b.sah_cost;
    out->build_seconds = ctx->build_seconds;
    out->tri_format = b.tri_format;
    out->max_depth = b.max_depth;
    for (int k = 0; k < 3; k++) { out->world_lo[k] = b.wlo[k]; out->world_hi[k] = b.whi[k]; }
    out->inflate = b.inflate;
    out->builder = ctx->builder_used;
    return SPB_OK;
}

int spb_set_option(spb_ctx* ctx, const char* name, int64_t value) {
    if (!ctx || !name) return fail(ctx, SPB_ERR_INVALID, "spb_set_option: NULL argument");
    const std::string n(name);
    if (n == "counters") ctx->opt_counters = value ? 1 : 0;
    else if (n == "trace_block") {
        if (value < 32 || value > 1024 || (value % 32)) return fail(ctx, SPB_ERR_INVALID, "trace_block must be a multiple of 32 in [32,1024]");
        ctx->opt_block = (int)value;
    } else if (n == "trace_ctas_per_sm") {
        if (value < 0 || value > 32) return fail(ctx, SPB_ERR_INVALID, "trace_ctas_per_sm must be in [0,32] (0 = occupancy query)");
        ctx->opt_ctas_per_sm = (int)value;
    } else if (n == "shade_minb") {
        if (value != 0 && (value < 4 || value > 6)) return fail(ctx, SPB_ERR_INVALID, "shade_minb must be 0 (default), 4, 5 or 6");
        ctx->opt_shade_minb = (int)value;
    } else if (n == "trace_variant") {
#ifdef SPB_EXPERIMENTAL_VARIANTS
        const bool ok = value >= 0 && value <= 25;
#else
        const bool ok = value >= 0 && value <= 5;
#endif
        if (!ok) return fail(ctx, SPB_ERR_INVALID, "trace_variant must be in [0,5] (measurement variants need a `make EXP=1` build)");
        ctx->opt_variant = (int)value;
    }
    else if (n == "render_graph") ctx->opt_render_graph = value ? 1 : 0;
    else if (n == "wave_slots") { if (value < 1024) return fail(ctx, SPB_ERR_INVALID, "wave_slots too small"); ctx->opt_wave_slots = value; }
    else if (n == "release_scratch") { cudaSetDevice(ctx->device); cudaDeviceSynchronize(); ctx->build_arena.release(); scratchRelease(ctx->device); }   // the builder's arena and the cached queue memory of this GPU go back to the driver
    else if (n == "chunk_rays") { if (value < 1024) return fail(ctx, SPB_ERR_INVALID, "chunk_rays too small"); ctx->opt_chunk = value; }
    else return fail(ctx, SPB_ERR_INVALID, "spb_set_option: unknown option " + n);
    return SPB_OK;
}

int spb_get_counters(spb_ctx* ctx, spb_counters* out) {
    if (!ctx || !out) return fail(ctx, SPB_ERR_INVALID, "spb_get_counters: NULL argument");
    out->last_kernel_ms = ctx->last_kernel_ms;
    out->kernel_launches = ctx->kernel_launches;
    out->rays = ctx->c_rays; out->node_visits = ctx->c_nodes; out->tri_tests = ctx->c_tris;
    return SPB_OK;
}

int spb_dev_alloc(spb_ctx* ctx, size_t bytes, void** d_ptr) {
    if (!ctx || !d_ptr) return fail(ctx, SPB_ERR_INVALID, "spb_dev_alloc: NULL argument");
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 16);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(ctx, SPB_ERR_OOM, "spb_dev_alloc: out of device memory"); }
    SPB_CUDA(ctx, e);
    return SPB_OK;
}
int spb_dev_free(spb_ctx* ctx, void* d_ptr) {
    if (!ctx) return fail(ctx, SPB_ERR_INVALID, "spb_dev_free: NULL ctx");
    cudaSetDevice(ctx->device);
    SPB_CUDA(ctx, cudaFree(d_ptr));
    return SPB_OK;
}
int spb_dev_upload(spb_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
    if (!ctx) return fail(ctx, SPB_ERR_INVALID, "spb_dev_upload: NULL ctx");
    cudaSetDevice(ctx->device);
    SPB_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPB_OK;
}
int spb_dev_download(spb_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
    if (!ctx) return fail(ctx, SPB_ERR_INVALID, "spb_dev_download: NULL ctx");
    cudaSetDevice(ctx->device);
    SPB_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPB_OK;
}
int spb_dev_sync(spb_ctx* ctx) {
    if (!ctx) return fail(ctx, SPB_ERR_INVALID, "spb_dev_sync: NULL ctx");
    cudaSetDevice(ctx->device);
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPB_OK;
}
void* spb_ctx_stream(spb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

}  // extern "C"
