// The unidirectional path tracer as a wavefront pipeline of sm_100a kernels.
//
// Replaces SamplerIntegrator::render (reference core/integrator.cc:46-110) and PathIntegrator::Li
// (integrators/path/path.cc:42-125) with uniformSampleOneLight / estimateDirectLight
// (core/mis.cc:21-136) for queues of paths:
//
//   control       one thread: statistics, queue counters, this iteration's regeneration plan
//   K3 generate   camera rays for the next (pixel, sample) pairs, appended   integrator.cc:80-86, perspective.cc:53-74
//                 to the extend queue behind the surviving paths
//   K1 extend     closest hit for the ray queue (trace_kernels.cuh)          scene.cc:37 -> bvh.cc:331-360
//   K4 classify   (scenes with several BSDF types) queue positions sorted into one index list per material bucket
//      shade      one kernel instance per bucket: emission, NEE light        path.cc:58-94, mis.cc:35-136
//                 sample + MIS BSDF sample, BSDF sample for the next segment, Russian roulette; survivors are pushed
//                 (ray + state) to the other extend queue, shadow and MIS rays to theirs
//   K2 connect    any-hit for the shadow queue; the unoccluded               visibility_tester.cc:21-24 -> bvh.cc:362-387
//                 contribution is added to the film inside the traversal kernel
//   K1' mis       closest hit for the (rare) MIS rays; the emission is       mis.cc:113-130
//                 added to the film inside the traversal kernel when the expected light is hit
//   K5 film       every radiance term (emission, unoccluded light sample,    film.cc:65-74, integrator.cc:88-90
//                 MIS hit) and each path's filter weight go straight into the RGBW film: one 128-bit reduction each
//
// STREAMING, not waves: the extend queue is topped up with new camera paths at the start of every
// iteration (path regeneration), so every launch works on a full queue until the samples of the call
// run out; a path's whole state (beta, sampler key, flags, filter weight) travels WITH its ray record,
// indexed by queue position, so the shade kernel reads and writes it coalesced and nothing is indexed
// by a per-path slot.  Queue sizes never visit the host: every kernel reads its count from device
// memory, the loop's bookkeeping runs in a one-thread control kernel, and two iterations (queue A,
// queue B) are captured once into a CUDA graph that the host re-launches until the control kernel
// reports an empty pipeline through mapped host memory (polled one launch behind, so the stream never drains).
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <dlfcn.h>

#include "context.h"
#include "shading.cuh"
#include "trace_kernels.cuh"

namespace spb {

// ---- device state -----------------------------------------------------------------------------------
struct Queues {
    float4* ray[2];       // extend queue (double buffered), 2 x float4 per entry = one spb_ray_f32 whose tmin carries the film index
    float4* state[2];     // parallel to ray[]: {beta.rgb, key0}, {flags, filter weight, key1, -}
                          //   flags: [11:0] bounces, [12] specularBounce, [31:16] path vertex counter (sampler dimension base)
    spb_hit* hit;         // parallel to ray[cur]
    float4*  shadow;      // connect queue (2 x float4 per entry), tmin = film index
    float4*  shadowC;     // contribution rgb, already multiplied by the path's filter weight
    float4*  mis;         // MIS closest-hit queue
    float4*  misC;        // contribution rgb (weighted), w = expected primitive id (bits) or -1 for "escapes"
    uint32_t* count;      // [0],[1] extend queue sizes, [2] shadow, [3] mis
};

// Control block of the streaming loop (device memory; written by controlKernel only).
struct LoopCtl {
    unsigned long long cursor, total;   // work items (pixel, sample) handed out so far / in this call
    unsigned long long stats[4];        // [0] paths [1] closest-hit rays [2] shadow rays [3] MIS rays, since spb_render_begin
    unsigned long long iterations;
    unsigned long long gen_item0;       // this iteration's regeneration: items [gen_item0, gen_item0 + gen_n) -> queue positions [gen_dst, ...)
    uint32_t gen_n, gen_dst;
    uint32_t capacity;
    int32_t  first, stride;             // sample index of item k: first + (k / pixels) * stride
};

struct DeviceScene {
    const ShadeTri* tris;
    const float*    vnormals;   // 9 per triangle or NULL
    const spb_material* mats;
    const spb_light* lights;
    int n_mats, n_lights;
    // textures: NULL mat_tex = no material is textured (the common case costs one uniform branch)
    const int2*        mat_tex;   // per material {kr texture, kt texture} or -1
    const spb_texture* texs;
    const float4*      texels;    // all bitmaps, rgb_
    const float*       uvs;       // 6 per triangle
    EnvMap env;
};

struct CameraPOD {
    double r2c[16], c2w[16];
    double lens_radius, focal;
};

struct RenderParamsPOD {
    int width, height, max_depth, filter, rr_start;
    int integrator;                          // SPB_INTEGRATOR_*
    float frx, fry, fbeta, fexpx, fexpy;     // filter parameters
    uint64_t seed;
};

// radiance into the film: one vector reduction per term (sm_90+: RED.E.ADD.F32x4)
__device__ __forceinline__ void filmAdd(float4* film, uint32_t index, float r, float g, float b, float w) {
    atomicAdd(film + index, make_float4(r, g, b, w));
}

struct SinkShadow {      // connect: add the contribution when NOTHING was hit
    const float4* rays; const float4* contrib; float4* film;
    __device__ __forceinline__ void store(int64_t i, const RayState& r) const {
        if (r.best_prim >= 0) return;
        const float4 c = contrib[i];
        filmAdd(film, __float_as_uint(rays[2 * i + 1].z), c.x, c.y, c.z, 0.f);
    }
};
struct SinkMis {         // MIS: add the contribution when exactly the expected light (or nothing) was hit
    const float4* rays; const float4* contrib; float4* film;
    __device__ __forceinline__ void store(int64_t i, const RayState& r) const {
        const float4 c = contrib[i];
        if (r.best_prim != (int32_t)__float_as_uint(c.w)) return;
        filmAdd(film, __float_as_uint(rays[2 * i + 1].z), c.x, c.y, c.z, 0.f);
    }
};

// One context's render worker: spb_render_samples_async / spb_film_reduce_async hand their host-side loops to it.
struct RenderWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv, cvDone;
    std::deque<std::function<int()>> tasks;
    int pending = 0;
    bool stop = false;
    int firstError = SPB_OK;
    std::string firstMessage;
};

struct RenderState {
    bool begun = false;
    spb_render_desc desc{};
    RenderParamsPOD rp{};
    CameraPOD cam{};
    // scene
    ShadeTri* d_tris = nullptr; float* d_vnormals = nullptr; int64_t n_tris = 0;
    spb_material* d_mats = nullptr; spb_light* d_lights = nullptr;
    std::vector<spb_material> mats; std::vector<spb_light> lights;
    std::vector<spb_texture> texs; std::vector<float> texels; std::vector<int32_t> mat_tex;   // textures + per-material bindings
    int2* d_mat_tex = nullptr; spb_texture* d_texs = nullptr; float4* d_texels = nullptr; float* d_uvs = nullptr;
    bool scene_dirty = true;
    bool sort_materials = false;        // more than one BSDF type in the scene: classify, then one shade instance per bucket
    uint32_t bucket_mask = 0u;          // buckets (kBucket*) the scene can produce
    uint8_t* d_prim_bucket = nullptr;   // per triangle: its bucket
    uint32_t* d_lists = nullptr; uint32_t* d_list_count = nullptr; int64_t list_stride = 0;
    uint32_t type_mask = 0x7fu;         // lobe types the scene's materials can produce: picks the shade kernel instance
    // envmap
    std::vector<float> env_rgb; int env_w = 0, env_h = 0; double env_l2w[16]; double env_scale = 1.0, env_center[3] = {0, 0, 0}, env_radius = 2.0;
    bool env_present = false, env_dirty = false;
    bool mis_any_hit = false;           // no area light in the scene: every MIS ray only asks whether it escapes (set by spb_render_begin)
    float4* d_env_texels = nullptr; float* d_env_floats = nullptr;
    DeviceScene ds{};
    // film + queues
    float4* d_film = nullptr; int64_t film_pixels = 0;
    int64_t slots = 0;                       // capacity of each queue
    void* d_pool = nullptr;
    Queues q{};
    LoopCtl* d_ctl = nullptr;
    unsigned long long* d_cursor = nullptr;  // 3 x 4 u64 ray cursors for the three trace launches of an iteration
    uint32_t* h_status = nullptr;            // mapped host memory: [0] pipeline empty, [1] iterations run
    uint32_t* d_status = nullptr;            // its device address
    cudaGraphExec_t graph_exec = nullptr;    // two iterations (queue 0, queue 1)
    cudaEvent_t poll[4] = {};                // the host polls h_status one graph launch behind
    cudaEvent_t ev_r0 = nullptr, ev_r1 = nullptr;
    cudaStream_t side = nullptr;             // the MIS launch of an iteration runs here, beside the connect launch
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int64_t launches = 0; int launches_per_iteration = 6; double render_ms = 0.0, reduce_ms = 0.0;
    void* d_scratch = nullptr; size_t scratch_bytes = 0;    // grow-only staging of spb_film_resolve* / spb_film_add
    RenderWorker* worker = nullptr;
    // nccl
    void* nccl_lib = nullptr; void* comm = nullptr;
    std::vector<void*> ipc_films;            // other processes' films mapped by spb_film_import_handles (cudaIpcOpenMemHandle)
};

static void workerStop(RenderState* R);
static int workerDrain(spb_ctx* ctx, RenderState* R);

static RenderState* rs(spb_ctx* ctx) {
    if (!ctx->render) ctx->render = new RenderState();
    return ctx->render;
}

// ---- K3: camera rays -------------------------------------------------------------------------------
// Transform::apply(Point3d) with its w-divide (core/transform.cc:60-76)
__device__ __forceinline__ void applyPoint(const double* m, double x, double y, double z, double* o) {
    double r[4];
#pragma unroll
    for (int i = 0; i < 4; i++) r[i] = m[i * 4 + 0] * x + m[i * 4 + 1] * y + m[i * 4 + 2] * z + m[i * 4 + 3];
    if (r[3] != 1.0) { const double w = r[3] + 1.0e-12; r[0] /= w; r[1] /= w; r[2] /= w; }
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
}
__device__ __forceinline__ void applyVector(const double* m, double x, double y, double z, double* o) {
#pragma unroll
    for (int i = 0; i < 3; i++) o[i] = m[i * 4 + 0] * x + m[i * 4 + 1] * y + m[i * 4 + 2] * z;
}

// ---- loop control: one thread, once per iteration, BEFORE the iteration's kernels -------------------------------
// Folds the sizes of the queues the previous iteration consumed into the statistics, clears them and the three
// ray cursors, and plans this iteration's regeneration: the extend queue `cur` (which holds the paths that
// survived the previous shade) is topped up to capacity with the next work items of the call.
__global__ void controlKernel(LoopCtl* ctl, Queues q, int cur, unsigned long long* cursors, uint32_t* status, uint32_t* bucketCount) {
    if (bucketCount && blockIdx.x == 0 && threadIdx.x < 16) bucketCount[threadIdx.x] = 0u;      // classifyKernel's lists (kNumBuckets <= 16 counters)
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int prev = cur ^ 1;
    ctl->stats[1] += q.count[prev]; ctl->stats[2] += q.count[2]; ctl->stats[3] += q.count[3];
    q.count[prev] = 0u; q.count[2] = 0u; q.count[3] = 0u;
    cursors[0] = 0ull; cursors[4] = 0ull; cursors[8] = 0ull;
    const uint32_t n0 = q.count[cur];
    const unsigned long long room = (unsigned long long)(ctl->capacity - n0), left = ctl->total - ctl->cursor;
    const uint32_t gen = (uint32_t)(room < left ? room : left);
    ctl->gen_n = gen; ctl->gen_dst = n0; ctl->gen_item0 = ctl->cursor;
    ctl->cursor += gen; ctl->stats[0] += gen;
    q.count[cur] = n0 + gen;
    ctl->iterations++;
    status[1] = (uint32_t)ctl->iterations;
    status[0] = (n0 + gen == 0u) ? 1u : 0u;       // nothing in flight and nothing left to start: the call is finished
    __threadfence_system();
}
// start of a spb_render_samples call
__global__ void loopBeginKernel(LoopCtl* ctl, Queues q, int first, int stride, unsigned long long total, uint32_t capacity, uint32_t* status) {
    ctl->cursor = 0ull; ctl->total = total; ctl->first = first; ctl->stride = stride; ctl->capacity = capacity;
    ctl->gen_n = 0u; ctl->gen_dst = 0u; ctl->gen_item0 = 0ull;
    q.count[0] = 0u; q.count[1] = 0u; q.count[2] = 0u; q.count[3] = 0u;
    status[0] = 0u;
    __threadfence_system();
}

__device__ __forceinline__ void writeRay(float4* q, uint32_t idx, V3 o, V3 d, uint32_t tag, float tmax) {
    q[2 * idx] = make_float4(o.x, o.y, o.z, d.x);
    q[2 * idx + 1] = make_float4(d.y, d.z, __uint_as_float(tag), tmax);
}
__device__ __forceinline__ void writeState(float4* q, uint32_t idx, V3 beta, uint2 key, uint32_t flags, float fw) {
    q[2 * idx] = make_float4(beta.x, beta.y, beta.z, __uint_as_float(key.x));
    q[2 * idx + 1] = make_float4(__uint_as_float(flags), fw, __uint_as_float(key.y), 0.f);
}

__device__ __forceinline__ float filterWeight(const RenderParamsPOD& rp, float dx, float dy) {
    switch (rp.filter) {
    case SPB_FILTER_TENT: return fmaxf(0.f, rp.frx - fabsf(dx)) * fmaxf(0.f, rp.fry - fabsf(dy));                 // filters/tent.cc:27-30
    case SPB_FILTER_GAUSSIAN: return fmaxf(0.f, expf(-rp.fbeta * dx * dx) - rp.fexpx) * fmaxf(0.f, expf(-rp.fbeta * dy * dy) - rp.fexpy);  // gaussian.cc:34-40
    default: return 1.f;                                                                                             // filters/box.cc:22
    }
}

// K3: the camera rays of this iteration's regeneration (core/integrator.cc:80-86, cameras/perspective.cc:53-74),
// appended to the extend queue `cur` behind the surviving paths.  Film::addPixel's weight w = filter(randFilm - 0.5)
// (core/film.cc:65-74) is fixed here and travels with the path; the pixel is stored as the FILM index, i.e. with
// the horizontal flip of core/integrator.cc:88 applied.
__global__ void __launch_bounds__(256) generateKernel(RenderParamsPOD rp, CameraPOD cam, Queues q, const LoopCtl* __restrict__ ctl, int cur) {
    const uint32_t n = ctl->gen_n, dst = ctl->gen_dst;
    const unsigned long long item0 = ctl->gen_item0;
    const int first = ctl->first, stride = ctl->stride;
    const unsigned long long npix = (unsigned long long)rp.width * rp.height;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long item = item0 + i;
        const uint32_t pixel = (uint32_t)(item % npix);
        const uint32_t sample = (uint32_t)(first + (int)(item / npix) * stride);
        const uint2 key = samplerKey(rp.seed, pixel, sample);
        const int x = pixel % rp.width, y = pixel / rp.width;
        // core/integrator.cc:84-86, cameras/perspective.cc:53-74
        const float f0 = sample1D(key, kDimFilm), f1 = sample1D(key, kDimFilm + 1);
        double pc[3];
        applyPoint(cam.r2c, (double)x + (double)f0, (double)y + (double)f1, 0.0, pc);
        double nrm = sqrt(pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2]);
        double dir[3] = {pc[0] / nrm, pc[1] / nrm, pc[2] / nrm};
        double org[3] = {0.0, 0.0, 0.0};
        if (cam.lens_radius > 0.0) {
            float lx, ly;
            concentricDisk(sample1D(key, kDimLens), sample1D(key, kDimLens + 1), &lx, &ly);
            const double ft = cam.focal / dir[2];
            const double pf[3] = {dir[0] * ft, dir[1] * ft, dir[2] * ft};
            org[0] = cam.lens_radius * lx; org[1] = cam.lens_radius * ly;
            double d2[3] = {pf[0] - org[0], pf[1] - org[1], pf[2] - org[2]};
            nrm = sqrt(d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2]);
            dir[0] = d2[0] / nrm; dir[1] = d2[1] / nrm; dir[2] = d2[2] / nrm;
        }
        double ow[3], dw[3];
        applyPoint(cam.c2w, org[0], org[1], org[2], ow);
        applyVector(cam.c2w, dir[0], dir[1], dir[2], dw);
        const double s = 1.0 / sqrt(dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2]);   // Ray ctor, core/ray.cc:15
        const uint32_t filmIndex = (uint32_t)y * (uint32_t)rp.width + (uint32_t)(rp.width - 1 - x);
        const float fw = filterWeight(rp, f0 - 0.5f, f1 - 0.5f);
        writeRay(q.ray[cur], dst + i, v3((float)ow[0], (float)ow[1], (float)ow[2]),
                 v3((float)(dw[0] * s), (float)(dw[1] * s), (float)(dw[2] * s)), filmIndex, kRayInf);
        writeState(q.state[cur], dst + i, v3(1.f), key, 0u, fw);
    }
}

// ---- K4: shade ---------------------------------------------------------------------------------------
__device__ __forceinline__ TriGeom loadTri(const ShadeTri* tris, int prim) {
    const float4* p = (const float4*)(tris + prim);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4), f = __ldg(p + 5), g = __ldg(p + 6);
    TriGeom t;
    t.p0 = v3(a.x, a.y, a.z); t.e1 = v3(a.w, b.x, b.y); t.e2 = v3(b.z, b.w, c.x);
    t.ng = v3(c.y, c.z, c.w); t.fn = v3(d.x, d.y, d.z); t.area = d.w;
    t.ss = v3(e.x, e.y, e.z); t.material = __float_as_int(e.w);
    t.ts = v3(f.x, f.y, f.z); t.light = __float_as_int(f.w);
    t.has_normals = __float_as_int(g.x);
    return t;
}

// Texture<Spectrum>::evaluate for the two texture plugins (textures/bitmap.cc:22-26, checkerboard.cc:31-40)
__device__ __forceinline__ V3 texTexel(const float4* texels, const spb_texture& t, int s, int tt) {
    s = (s % t.width + t.width) % t.width; tt = (tt % t.height + t.height) % t.height;            // ImageWrap::Repeat (mipmap.cc:101-104)
    const float4 v = __ldg(texels + t.texel_offset + (size_t)tt * t.width + s);
    return v3(v.x, v.y, v.z);
}
__device__ __forceinline__ V3 evalTexture(const DeviceScene& sc, int id, float u, float v) {
    const spb_texture t = sc.texs[id];
    if (t.type == SPB_TEX_CHECKERBOARD) {
        const int iu = (int)((u * t.uscale + t.uoffset) * 2.f), iv = (int)((v * t.vscale + t.voffset) * 2.f);
        return ((iu + iv) % 2 != 0) ? v3(t.color0[0], t.color0[1], t.color0[2]) : v3(t.color1[0], t.color1[1], t.color1[2]);
    }
    const float s = u * t.width - 0.5f, tt = (1.f - v) * t.height - 0.5f;                           // MipMap::lookup -> bilinear(0, st)
    const int si = (int)s, ti = (int)tt;
    const float ds = s - si, dt = tt - ti;
    return (1.f - ds) * (1.f - dt) * texTexel(sc.texels, t, si, ti) + ds * (1.f - dt) * texTexel(sc.texels, t, si + 1, ti) +
           (1.f - ds) * dt * texTexel(sc.texels, t, si, ti + 1) + ds * dt * texTexel(sc.texels, t, si + 1, ti + 1);
}

// ---- material-sorted shading -------------------------------------------------------------------------------------------
// A scene with more than one BSDF type is shaded by ONE KERNEL INSTANCE PER BUCKET: classifyKernel sorts the queue
// positions of an iteration into index lists by the BSDF type of the surface that was hit, and each list is shaded by
// the instance that contains the code of that lobe only.  ncu on the r01 kernel (one instance carrying every lobe of
// the glossy Cornell box, sorted per 512-entry tile inside the kernel): 15 % issue-slot utilisation with 12.8 of every
// 16 stall cycles waiting for INSTRUCTIONS -- warps on different materials thrash the instruction cache
// (profiles/r02a_shade_glossy_ncu.txt).  Records are 32-byte aligned, so gathering them through an index list costs
// no extra sectors.  A scene with a single BSDF type skips the classification and shades in queue order.
enum { kBucketMiss = 0, kBucketNone = 1, kBucketMat0 = 2, kNumBuckets = 9 };      // 2 + SPB_MAT_*

struct ShadeLists {
    uint32_t* index;       // kNumBuckets lists of `stride` entries each (only the buckets a scene can produce are touched)
    uint32_t* count;       // kNumBuckets counters
    uint32_t  stride;
};

__global__ void __launch_bounds__(256) classifyKernel(Queues q, int cur, const uint8_t* __restrict__ primBucket, ShadeLists L) {
    // list positions are reserved per CTA and tile of 256 entries: one global atomic per bucket (see the queue pushes of shadeKernel)
    __shared__ uint32_t s_cnt[16], s_base[16];
    const uint32_t n = q.count[cur];
    const int lane = threadIdx.x & 31;
    for (uint32_t base = blockIdx.x * 256u; base < n; base += gridDim.x * 256u) {
        if (threadIdx.x < 16) s_cnt[threadIdx.x] = 0u;
        __syncthreads();
        const uint32_t i = base + threadIdx.x;
        uint32_t key = 15u;                       // past the end of the queue: counted nowhere
        if (i < n) {
            const int prim = __float_as_int(__ldcs((const float*)(q.hit + i) + 1));
            key = prim < 0 ? (uint32_t)kBucketMiss : (uint32_t)__ldg(primBucket + prim);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(peers) - 1;
        uint32_t at = 0u;
        if (lane == leader && key != 15u) at = atomicAdd(&s_cnt[key], (uint32_t)__popc(peers));
        at = __shfl_sync(0xffffffffu, at, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        __syncthreads();
        if (threadIdx.x < kNumBuckets && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(L.count + threadIdx.x, s_cnt[threadIdx.x]);
        __syncthreads();
        if (key != 15u) L.index[(size_t)key * L.stride + s_base[key] + at] = i;
    }
}

// TYPES: the lobe types (bit t = SPB_MAT_t) this instance contains; every other lobe is compiled out (shading.cuh,
// lobeLive).  MINB: resident CTAs per SM it is compiled for.  LIST: shade the queue positions of an index list
// (a bucket of classifyKernel) instead of the whole queue in order.
constexpr int kShadeMinbDiffuse = 4;
constexpr uint32_t kTypesAll = 0x7fu;
constexpr uint32_t kTypesDiffuse = 1u << SPB_MAT_DIFFUSE;
constexpr uint32_t typeClosure(int t) {     // what a material of type t can turn into (makeBsdf: alpha 0 -> the specular lobe, a black coating -> Lambertian)
    return (1u << t) | (t == SPB_MAT_ROUGHCONDUCTOR ? 1u << SPB_MAT_CONDUCTOR : 0u) | (t == SPB_MAT_ROUGHDIELECTRIC ? 1u << SPB_MAT_DIELECTRIC : 0u) |
           (t == SPB_MAT_ROUGHPLASTIC ? 1u << SPB_MAT_DIFFUSE : 0u);
}

template <uint32_t TYPES, int MINB, bool LIST>
__global__ void __launch_bounds__(128, MINB) shadeKernel(RenderParamsPOD rp, DeviceScene sc, Queues q, float4* film, int cur,
                                                       const uint32_t* __restrict__ list, const uint32_t* __restrict__ listCount) {
    const uint32_t n = LIST ? *listCount : q.count[cur];
    const float4* rays = q.ray[cur];
    const float4* states = q.state[cur];
    float4* nextQ = q.ray[cur ^ 1];
    float4* nextS = q.state[cur ^ 1];
    uint32_t* nextCount = q.count + (cur ^ 1);
    // whole CTAs iterate together: the queue pushes of a tile are aggregated over the CTA.
    // (Staging the five 16-byte pieces of an entry through shared memory with cp.async, one tile ahead, was measured in r02i:
    // no gain -- 2134 -> 2115 Msamples/s on the diffuse Cornell box; the other resident warps already cover that latency.)
    __shared__ uint32_t s_cnt[2][4], s_base[4];
    if (threadIdx.x < 8) s_cnt[threadIdx.x >> 2][threadIdx.x & 3] = 0u;
    __syncthreads();
    uint32_t tile = 0;
    for (uint32_t base = blockIdx.x * 128u; base < n; base += gridDim.x * 128u) {
      {
        const uint32_t j = base + threadIdx.x;
        const bool valid = j < n;
        const uint32_t i = (LIST && valid) ? __ldg(list + j) : j;
        bool pushNext = false, pushShadow = false, pushMis = false;
        V3 nO = v3(0.f), nD = v3(0.f), sO = v3(0.f), sD = v3(0.f), mO = v3(0.f), mD = v3(0.f);
        float sT = 0.f;
        V3 sC = v3(0.f), mC = v3(0.f);
        int mExpect = -1;
        uint32_t pix = 0, flags = 0;
        uint2 key = make_uint2(0u, 0u);
        float fw = 0.f;
        V3 beta = v3(0.f);
        if (valid) {
            const float4 r0 = __ldcs(rays + 2 * i), r1 = __ldcs(rays + 2 * i + 1);
            const float4 s0 = __ldcs(states + 2 * i), s1 = __ldcs(states + 2 * i + 1);
            const float4 hv = __ldcs((const float4*)(q.hit + i));
            pix = __float_as_uint(r1.z);
            const V3 d = v3(r0.w, r1.x, r1.y);
            const int prim = __float_as_int(hv.y);
            flags = __float_as_uint(s1.x);
            fw = s1.y;
            key = make_uint2(__float_as_uint(s0.w), __float_as_uint(s1.z));
            const int bounces = (int)(flags & 0xfffu);
            const bool specularBounce = (flags >> 12) & 1u;
            const uint32_t vertex = flags >> 16;
            beta = v3(s0.x, s0.y, s0.z);
            V3 Ladd = v3(0.f);
            const uint32_t dim = kDimBounce0 + vertex * kDimsPerBounce;

            // directlighting (integrators/directlighting/directlighting.cc:21-57): emitted light only at depth 0, one
            // light sample at every vertex, continuation only through a specular REFLECTION lobe
            // (SamplerIntegrator::specularReflect, core/integrator.cc:112-130; specularTransmit never finds a
            // matching lobe among the in-scope materials), depth growing by 2 per reflection (:53-54 + :125).
            const bool direct = rp.integrator == SPB_INTEGRATOR_DIRECT;
            if (prim < 0) {
                // escaped: path.cc:62-65 / directlighting.cc:30-35 (only environment lights answer Le(ray))
                if ((direct || bounces == 0 || specularBounce) && sc.env.present) {
                    int nEnv = 0;
                    for (int l = 0; l < sc.n_lights; l++) nEnv += (sc.lights[l].type == SPB_LIGHT_ENVMAP);
                    Ladd = beta * envLe(sc.env, d) * (float)nEnv;
                }
            } else {
                const TriGeom tri = loadTri(sc.tris, prim);
                // path.cc:58-61 + SurfaceInteraction::Le -> AreaLight::L (interaction.cc:160-163, area.cc:29-31)
                const bool emitHere = direct ? (bounces == 0 && tri.material >= 0 && tri.material < sc.n_mats)   // after the bsdf check (:38-46)
                                             : (bounces == 0 || specularBounce);
                if (emitHere && tri.light >= 0) {
                    const spb_light lt = sc.lights[tri.light];
                    if (dot(tri.ng, -d) > 0.f) Ladd = beta * v3(lt.radiance[0], lt.radiance[1], lt.radiance[2]);
                }
                if (direct || bounces < rp.max_depth) {             // path.cc:68
                    // ---- surface point (core/triangle.cc:119-156)
                    SurfacePoint sp;
                    const float u = hv.z, v = hv.w;
                    sp.p = tri.p0 + u * tri.e1 + v * tri.e2;
                    sp.ng = tri.ng; sp.ss = tri.ss; sp.ts = tri.ts; sp.ns = tri.ng;
                    if (tri.has_normals) {
                        const float* vn = sc.vnormals + (size_t)prim * 9;
                        const float w0 = 1.f - u - v;
                        const V3 ns = normalize(v3(w0 * vn[0] + u * vn[3] + v * vn[6], w0 * vn[1] + u * vn[4] + v * vn[7],
                                                   w0 * vn[2] + u * vn[5] + v * vn[8]));
                        if (fabsf(dot(ns, tri.fn)) < 1.0f - 1e-6f) {
                            sp.ss = normalize(cross(ns, tri.fn));
                            sp.ts = normalize(cross(ns, sp.ss));
                            sp.ns = normalize(cross(sp.ss, sp.ts));
                        }
                    }
                    const V3 wo = -d;
                    const bool hasMat = tri.material >= 0 && tri.material < sc.n_mats;
                    spb_material mat = sc.mats[hasMat ? tri.material : 0];
                    if (sc.mat_tex && hasMat) {
                        const int2 tx = sc.mat_tex[tri.material];
                        if ((tx.x & tx.y) != -1) {
                            const float* t = sc.uvs + (size_t)prim * 6;
                            const float w0 = 1.f - u - v;                                               // core/triangle.cc:120
                            const float tu = w0 * t[0] + u * t[2] + v * t[4], tv = w0 * t[1] + u * t[3] + v * t[5];
                            if (tx.x >= 0) { const V3 c = evalTexture(sc, tx.x, tu, tv); mat.kr[0] = c.x; mat.kr[1] = c.y; mat.kr[2] = c.z; }
                            if (tx.y >= 0) { const V3 c = evalTexture(sc, tx.y, tu, tv); mat.kt[0] = c.x; mat.kt[1] = c.y; mat.kt[2] = c.z; }
                        }
                    }
                    const Bsdf bsdf = makeBsdf(mat, TYPES);
                    if (!hasMat || bsdf.type == SPB_MAT_NONE && false) {
                        // no material: pass through without counting a bounce (path.cc:71-75)
                        nO = offsetRayOrigin(sp.p, sp.ng, d); nD = d; pushNext = true;
                        flags = (flags & 0xffffu) | ((vertex + 1u) << 16);
                    } else {
                        // ---- direct lighting: uniformSampleOneLight / estimateDirectLight (mis.cc:21-136)
                        if (bsdfNumComponents(bsdf, kBxNonSpecular) > 0 && sc.n_lights > 0) {
                            const float nl = (float)sc.n_lights;
                            int lid = (int)(sample1D(key, dim + kDLightPick) * nl);
                            if (lid > sc.n_lights - 1) lid = sc.n_lights - 1;
                            const spb_light lt = sc.lights[lid];
                            const float ul0 = sample1D(key, dim + kDLight), ul1 = sample1D(key, dim + kDLight + 1);
                            const float us0 = sample1D(key, dim + kDShade), us1 = sample1D(key, dim + kDShade + 1);
                            const float us2 = sample1D(key, dim + kDLobeMis);
                            if (lt.type == SPB_LIGHT_AREA) {
                                const TriGeom lg = loadTri(sc.tris, lt.prim);
                                const V3 Le = v3(lt.radiance[0], lt.radiance[1], lt.radiance[2]);
                                // light sampling: AreaLight::sampleLi (area.cc:33-41), Triangle::sample (triangle.cc:180-197)
                                float a = ul0, b = ul1;
                                if (a + b >= 1.f) { a = 1.f - a; b = 1.f - b; }
                                const V3 pl = lg.p0 + a * lg.e1 + b * lg.e2;
                                V3 nlgt = lg.fn;
                                if (lg.has_normals) {
                                    const float* vn = sc.vnormals + (size_t)lt.prim * 9;
                                    const float w0 = 1.f - a - b;
                                    nlgt = v3(w0 * vn[0] + a * vn[3] + b * vn[6], w0 * vn[1] + a * vn[4] + b * vn[7],
                                              w0 * vn[2] + a * vn[5] + b * vn[8]);
                                }
                                const V3 toL = pl - sp.p;
                                const V3 wi = normalize(toL);
                                const float lightPdf = trianglePdfSolidAngle(lg, sp, wi);
                                if (lightPdf > 0.f && dot(nlgt, -wi) > 0.f && !isBlack(Le)) {
                                    const V3 f = bsdfF(bsdf, sp, wo, wi, kBxNonSpecular) * absDot(wi, sp.ns);
                                    if (!isBlack(f)) {
                                        const float bp = bsdfPdf(bsdf, sp, wo, wi, kBxNonSpecular);
                                        const float wgt = powerHeuristic(lightPdf, bp);
                                        // VisibilityTester: Interaction::spawnRayTo(Interaction) (interaction.cc:88-93)
                                        sO = offsetRayOrigin(sp.p, sp.ng, toL);
                                        const V3 tgt = offsetRayOrigin(pl, nlgt, sO - pl);
                                        sD = tgt - sO; sT = length(sD);
                                        sC = beta * f * Le * (nl * wgt / lightPdf);
                                        pushShadow = !isBlack(sC) && sT > 0.f;
                                    }
                                }
                                // BSDF sampling (mis.cc:86-133)
                                V3 wi2; float bp2; int st2;
                                V3 f2 = bsdfSample(bsdf, sp, wo, us0, us1, us2, kBxNonSpecular, &wi2, &bp2, &st2);
                                f2 = f2 * absDot(wi2, sp.ns);
                                if (!isBlack(f2) && bp2 > 0.f) {
                                    const float lp2 = trianglePdfSolidAngle(lg, sp, wi2);        // AreaLight::pdfLi (area.cc:43-45)
                                    if (lp2 > 0.f && dot(lg.ng, -wi2) > 0.f) {
                                        const float wgt = powerHeuristic(bp2, lp2);
                                        mO = offsetRayOrigin(sp.p, sp.ng, wi2); mD = wi2;
                                        mC = beta * f2 * Le * (nl * wgt / bp2);
                                        mExpect = lt.prim;
                                        pushMis = !isBlack(mC);
                                    }
                                }
                            } else if (sc.env.present) {
                                // Envmap::sampleLi (envmap.cc:60-79)
                                V3 wi; float lightPdf;
                                const V3 Li = envSample(sc.env, ul0, ul1, &wi, &lightPdf);
                                if (lightPdf > 0.f && !isBlack(Li)) {
                                    const V3 f = bsdfF(bsdf, sp, wo, wi, kBxNonSpecular) * absDot(wi, sp.ns);
                                    if (!isBlack(f)) {
                                        const float bp = bsdfPdf(bsdf, sp, wo, wi, kBxNonSpecular);
                                        const float wgt = powerHeuristic(lightPdf, bp);
                                        const V3 pl = sp.p + wi * (2.f * sc.env.radius);
                                        sO = offsetRayOrigin(sp.p, sp.ng, pl - sp.p);
                                        sD = pl - sO; sT = length(sD);
                                        sC = beta * f * Li * (nl * wgt / lightPdf);
                                        pushShadow = !isBlack(sC) && sT > 0.f;
                                    }
                                }
                                V3 wi2; float bp2; int st2;
                                V3 f2 = bsdfSample(bsdf, sp, wo, us0, us1, us2, kBxNonSpecular, &wi2, &bp2, &st2);
                                f2 = f2 * absDot(wi2, sp.ns);
                                if (!isBlack(f2) && bp2 > 0.f) {
                                    const float lp2 = envPdf(sc.env, wi2);
                                    if (lp2 > 0.f) {
                                        const float wgt = powerHeuristic(bp2, lp2);
                                        mO = offsetRayOrigin(sp.p, sp.ng, wi2); mD = wi2;
                                        mC = beta * f2 * envLe(sc.env, wi2) * (nl * wgt / bp2);
                                        mExpect = -1;
                                        pushMis = !isBlack(mC);
                                    }
                                }
                            }
                        }
                        // ---- next segment (path.cc:82-94; directlighting.cc:52-55)
                        V3 wi; float pdf = 0.f; int sampled = 0;
                        V3 f = v3(0.f);
                        if (!direct) {
                            f = bsdfSample(bsdf, sp, wo, sample1D(key, dim + kDBsdf), sample1D(key, dim + kDBsdf + 1),
                                           sample1D(key, dim + kDLobePath), kBxAll, &wi, &pdf, &sampled);
                        } else if (2 * bounces + 1 < rp.max_depth) {
                            f = bsdfSample(bsdf, sp, wo, sample1D(key, dim + kDBsdf), sample1D(key, dim + kDBsdf + 1),
                                           sample1D(key, dim + kDLobePath), kBxReflection | kBxSpecular, &wi, &pdf, &sampled);
                            if (absDot(wi, sp.ns) == 0.f) pdf = 0.f;
                        }
                        if (direct) {
                            if (!isBlack(f) && pdf > 0.f) {
                                beta = beta * f * (absDot(wi, sp.ns) / pdf);
                                nO = offsetRayOrigin(sp.p, sp.ng, wi); nD = wi; pushNext = true;
                                flags = (uint32_t)(bounces + 1) | 0x1000u | ((vertex + 1u) << 16);
                            }
                        } else if (!isBlack(f) && pdf != 0.f) {
                            beta = beta * f * (absDot(wi, sp.ns) / pdf);
                            bool alive = true;
                            if (bounces > rp.rr_start) {                       // path.cc:117-121
                                const float qc = fminf(0.95f, gray(beta));
                                if (!(sample1D(key, dim + kDRoulette) <= qc) || !(qc > 0.f)) alive = false;
                                else beta = beta / qc;
                            }
                            if (alive) {
                                nO = offsetRayOrigin(sp.p, sp.ng, wi); nD = wi; pushNext = true;
                                flags = (uint32_t)(bounces + 1) | ((sampled & kBxSpecular) ? 0x1000u : 0u) | ((vertex + 1u) << 16);
                            }
                        }
                    }
                }
            }
            // K5: this vertex's emitted light, and -- when the path ends here -- the sample's filter weight
            // (Film::addPixel, core/film.cc:65-74: img += w L, wsum += w), in one reduction
            if (!pushNext || !isBlack(Ladd)) filmAdd(film, pix, fw * Ladd.x, fw * Ladd.y, fw * Ladd.z, pushNext ? 0.f : fw);
            sC = sC * fw; mC = mC * fw;
        }
        // ---- queue pushes, aggregated over the CTA: one global atomic per queue and tile of 128 entries.  With one atomic per
        // warp (r02b) 45 % of this kernel's stall samples sat on the return of ATOM.ADD to the three queue counters: ~1.5 M
        // same-address atomics per iteration is all the L2 does for one address (profiles/r02d_shade_diffuse_ncu.txt).
        __syncwarp();
        const int lane = threadIdx.x & 31;
        const unsigned below = (1u << lane) - 1u;
        const unsigned mN = __ballot_sync(0xffffffffu, pushNext), mS = __ballot_sync(0xffffffffu, pushShadow), mM = __ballot_sync(0xffffffffu, pushMis);
        uint32_t* cnt = s_cnt[tile & 1];
        uint32_t wN = 0, wS = 0, wM = 0;             // this warp's offsets inside the CTA's reservations
        if (lane == 0) {
            if (mN) wN = atomicAdd(&cnt[0], (uint32_t)__popc(mN));
            if (mS) wS = atomicAdd(&cnt[1], (uint32_t)__popc(mS));
            if (mM) wM = atomicAdd(&cnt[2], (uint32_t)__popc(mM));
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            const uint32_t c = cnt[threadIdx.x];
            uint32_t* counter = threadIdx.x == 0 ? nextCount : q.count + 1 + threadIdx.x;
            s_base[threadIdx.x] = c ? atomicAdd(counter, c) : 0u;
            s_cnt[(tile & 1) ^ 1][threadIdx.x] = 0u;                 // the other buffer is the next tile's
        }
        __syncthreads();
        wN = __shfl_sync(0xffffffffu, wN, 0) + s_base[0]; wS = __shfl_sync(0xffffffffu, wS, 0) + s_base[1]; wM = __shfl_sync(0xffffffffu, wM, 0) + s_base[2];
        if (pushNext) { const uint32_t in = wN + __popc(mN & below); writeRay(nextQ, in, nO, nD, pix, kRayInf); writeState(nextS, in, beta, key, flags, fw); }
        if (pushShadow) { const uint32_t is = wS + __popc(mS & below); writeRay(q.shadow, is, sO, sD, pix, sT); q.shadowC[is] = make_float4(sC.x, sC.y, sC.z, 0.f); }
        if (pushMis) { const uint32_t im = wM + __popc(mM & below); writeRay(q.mis, im, mO, mD, pix, kRayInf); q.misC[im] = make_float4(mC.x, mC.y, mC.z, __uint_as_float((uint32_t)mExpect)); }
        tile++;
      }
    }
}

// escaped rays of a classified iteration: only the environment answers (path.cc:62-65, directlighting.cc:30-35), and the path ends
__global__ void __launch_bounds__(256) shadeMissKernel(RenderParamsPOD rp, DeviceScene sc, Queues q, float4* film, int cur,
                                                     const uint32_t* __restrict__ list, const uint32_t* __restrict__ listCount) {
    const uint32_t n = *listCount;
    int nEnv = 0;
    if (sc.env.present) for (int l = 0; l < sc.n_lights; l++) nEnv += (sc.lights[l].type == SPB_LIGHT_ENVMAP);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t i = __ldg(list + j);
        const float4 r0 = __ldcs(q.ray[cur] + 2 * i), r1 = __ldcs(q.ray[cur] + 2 * i + 1);
        const float4 s0 = __ldcs(q.state[cur] + 2 * i), s1 = __ldcs(q.state[cur] + 2 * i + 1);
        const uint32_t flags = __float_as_uint(s1.x);
        const float fw = s1.y;
        V3 L = v3(0.f);
        if (nEnv > 0 && (rp.integrator == SPB_INTEGRATOR_DIRECT || (flags & 0xfffu) == 0u || ((flags >> 12) & 1u)))
            L = v3(s0.x, s0.y, s0.z) * envLe(sc.env, v3(r0.w, r1.x, r1.y)) * (float)nEnv;
        filmAdd(film, __float_as_uint(r1.z), fw * L.x, fw * L.y, fw * L.z, fw);
    }
}

// ---- film read-out ------------------------------------------------------------------------------------
__global__ void resolveKernel(const float4* film, float* rgb, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = film[i];
    const float inv = 1.0f / (v.w + 1.0e-12f);                               // core/film.cc:27
    rgb[i * 3 + 0] = v.x * inv; rgb[i * 3 + 1] = v.y * inv; rgb[i * 3 + 2] = v.z * inv;
}
// the film plugins' pixel encodings (core/image.cc:60-88, core/tmo.cc:53-71 + core/image.cc:484-487), on the resolved float pixel
__global__ void encodeRgbeKernel(const float4* film, uchar4* out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = film[i];
    const float inv = 1.0f / (v.w + 1.0e-12f);
    const double r = (double)(v.x * inv), g = (double)(v.y * inv), b = (double)(v.z * inv);
    double d = fmax(r, fmax(g, b));
    if (!(d > 1.0e-32)) { out[i] = make_uchar4(0, 0, 0, 0); return; }
    int ie;
    const double m = frexp(d, &ie);
    d = m * 256.0 / d;
    out[i] = make_uchar4((unsigned char)(r * d), (unsigned char)(g * d), (unsigned char)(b * d), (unsigned char)(ie + 128));
}
__global__ void encodeLdrKernel(const float4* film, unsigned char* out, double invGamma, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = film[i];
    const float inv = 1.0f / (v.w + 1.0e-12f);
    const double c[3] = {(double)(v.x * inv), (double)(v.y * inv), (double)(v.z * inv)};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double t = fmin(1.0, fmax(0.0, pow(c[k], invGamma)));
        out[i * 3 + k] = (unsigned char)(255.0 * t);
    }
}
__global__ void filmAddKernel(float4* film, const float4* add, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = film[i], b = add[i];
    film[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// ---- host side ---------------------------------------------------------------------------------------
static void freeScene(RenderState* R) {
    if (R->d_tris) cudaFree(R->d_tris);
    if (R->d_vnormals) cudaFree(R->d_vnormals);
    if (R->d_mats) cudaFree(R->d_mats);
    if (R->d_lights) cudaFree(R->d_lights);
    if (R->d_mat_tex) cudaFree(R->d_mat_tex);
    if (R->d_texs) cudaFree(R->d_texs);
    if (R->d_texels) cudaFree(R->d_texels);
    if (R->d_uvs) cudaFree(R->d_uvs);
    if (R->d_prim_bucket) cudaFree(R->d_prim_bucket);
    R->d_prim_bucket = nullptr;
    R->d_tris = nullptr; R->d_vnormals = nullptr; R->d_mats = nullptr; R->d_lights = nullptr;
    R->d_mat_tex = nullptr; R->d_texs = nullptr; R->d_texels = nullptr; R->d_uvs = nullptr;
}
static void freeEnv(RenderState* R) {
    if (R->d_env_texels) cudaFree(R->d_env_texels);
    if (R->d_env_floats) cudaFree(R->d_env_floats);
    R->d_env_texels = nullptr; R->d_env_floats = nullptr;
}

void renderStateDestroy(spb_ctx* ctx) {
    RenderState* R = ctx->render;
    if (!R) return;
    workerDrain(ctx, R);
    workerStop(R);
    freeScene(R); freeEnv(R);
    if (R->graph_exec) cudaGraphExecDestroy(R->graph_exec);
    if (R->d_film) cudaFree(R->d_film);
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    scratchFree(R->d_pool, ctx->stream);
    if (R->d_ctl) cudaFree(R->d_ctl);
    if (R->d_cursor) cudaFree(R->d_cursor);
    if (R->d_scratch) cudaFree(R->d_scratch);
    scratchFree(R->d_lists, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    if (R->d_list_count) cudaFree(R->d_list_count);
    if (R->h_status) cudaFreeHost(R->h_status);
    for (cudaEvent_t e : R->poll) if (e) cudaEventDestroy(e);
    if (R->ev_r0) cudaEventDestroy(R->ev_r0);
    if (R->ev_r1) cudaEventDestroy(R->ev_r1);
    if (R->ev_fork) cudaEventDestroy(R->ev_fork);
    if (R->ev_join) cudaEventDestroy(R->ev_join);
    if (R->side) cudaStreamDestroy(R->side);
    for (void* p : R->ipc_films) cudaIpcCloseMemHandle(p);
    if (R->comm && R->nccl_lib) {
        typedef int (*destroy_t)(void*);
        destroy_t f = (destroy_t)dlsym(R->nccl_lib, "ncclCommDestroy");
        if (f) f(R->comm);
    }
    delete R;
    ctx->render = nullptr;
}

// Geometry, attributes, materials, lights, textures or the environment changed: what spb_render_begin uploaded is stale
// (or, after spb_scene_set_triangles, freed).  Rendering needs a new spb_render_begin (SPB_ERR_INVALID otherwise).
void renderSceneChanged(spb_ctx* ctx) {
    RenderState* R = ctx->render;
    if (!R) return;
    workerDrain(ctx, R);
    R->scene_dirty = true;
    R->begun = false;
}

void renderStateEnsure(spb_ctx* ctx) { rs(ctx); }

void renderSceneClone(spb_ctx* dst, spb_ctx* src) {
    RenderState* S = src->render;
    renderSceneChanged(dst);
    if (!S) return;
    workerDrain(src, S);
    RenderState* D = rs(dst);
    workerDrain(dst, D);
    D->mats = S->mats; D->lights = S->lights; D->texs = S->texs; D->texels = S->texels; D->mat_tex = S->mat_tex;
    D->env_rgb = S->env_rgb; D->env_w = S->env_w; D->env_h = S->env_h; D->env_scale = S->env_scale; D->env_radius = S->env_radius;
    std::memcpy(D->env_l2w, S->env_l2w, sizeof(D->env_l2w)); std::memcpy(D->env_center, S->env_center, sizeof(D->env_center));
    D->env_present = S->env_present; D->env_dirty = true;
    D->scene_dirty = true; D->begun = false;
}

// The per-triangle shading records, the way the reference's Triangle constructors and Triangle::intersect build them
// (core/triangle.cc:17-67, 119-137): in double, narrowed to float32 at the end; one thread per triangle.  Also the shading
// bucket of the triangle (classifyKernel).
struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3sub(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 d3cross(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double d3norm(D3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ D3 d3normalized(D3 a) { const double n = d3norm(a); return n > 0 ? D3{a.x / n, a.y / n, a.z / n} : a; }
__device__ __forceinline__ D3 d3scaled(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ D3 d3add(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ void d3coordSys(D3 w, D3* u, D3* v) {          // core/vect_math.h:115-123
    if (fabs(w.x) > fabs(w.y)) *u = d3normalized({-w.z, 0, w.x}); else *u = d3normalized({0, w.z, -w.y});
    *v = d3normalized(d3cross(w, *u));
}

__global__ void __launch_bounds__(256) shadeTriKernel(const double* __restrict__ verts, const float* __restrict__ normals, const float* __restrict__ uvs,
                                                      const int32_t* __restrict__ material_id, const int32_t* __restrict__ light_id, int64_t n,
                                                      const spb_material* __restrict__ mats, int n_mats, ShadeTri* __restrict__ out, uint8_t* __restrict__ bucket) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* v = verts + i * 9;
    const D3 p0{v[0], v[1], v[2]}, p1{v[3], v[4], v[5]}, p2{v[6], v[7], v[8]};
    const D3 e1 = d3sub(p1, p0), e2 = d3sub(p2, p0);
    D3 fn = d3cross(e1, e2);
    bool triHasN = false;
    if (normals) {
        const float* nn = normals + i * 9;
        triHasN = !(nn[0] == 0.f && nn[1] == 0.f && nn[2] == 0.f && nn[3] == 0.f && nn[4] == 0.f && nn[5] == 0.f && nn[6] == 0.f && nn[7] == 0.f && nn[8] == 0.f);
        if (triHasN && d3norm(fn) < 1e-12) fn = d3scaled(D3{(double)nn[0] + nn[3] + nn[6], (double)nn[1] + nn[4] + nn[7], (double)nn[2] + nn[5] + nn[8]}, 1.0 / 3.0);
    }
    fn = d3normalized(fn);
    D3 dpdu, dpdv;
    double detUV = 0.0, duv01[2] = {0, 0}, duv02[2] = {0, 0};
    if (uvs) {
        const float* t = uvs + i * 6;
        duv01[0] = (double)t[2] - t[0]; duv01[1] = (double)t[3] - t[1];
        duv02[0] = (double)t[4] - t[0]; duv02[1] = (double)t[5] - t[1];
        detUV = duv01[0] * duv02[1] - duv01[1] * duv02[0];
    }
    if (detUV == 0.0) d3coordSys(fn, &dpdu, &dpdv);
    else {
        const double inv = 1.0 / detUV;
        dpdu = d3add(d3scaled(e1, duv02[1] * inv), d3scaled(e2, -duv01[1] * inv));
        dpdv = d3add(d3scaled(e1, -duv02[0] * inv), d3scaled(e2, duv01[0] * inv));
    }
    const D3 ng = d3normalized(d3cross(dpdu, dpdv));
    const D3 ss = d3normalized(dpdu), ts = d3normalized(dpdv);
    ShadeTri s;
    s.p0[0] = (float)p0.x; s.p0[1] = (float)p0.y; s.p0[2] = (float)p0.z;
    s.e1x = (float)e1.x; s.e1y = (float)e1.y; s.e1z = (float)e1.z;
    s.e2[0] = (float)e2.x; s.e2[1] = (float)e2.y; s.e2z = (float)e2.z;
    s.ng[0] = (float)ng.x; s.ng[1] = (float)ng.y; s.ng[2] = (float)ng.z;
    s.fn[0] = (float)fn.x; s.fn[1] = (float)fn.y; s.fn[2] = (float)fn.z;
    s.area = (float)(0.5 * d3norm(d3cross(e1, e2)));
    s.ss[0] = (float)ss.x; s.ss[1] = (float)ss.y; s.ss[2] = (float)ss.z;
    s.ts[0] = (float)ts.x; s.ts[1] = (float)ts.y; s.ts[2] = (float)ts.z;
    const int32_t m = material_id[i];
    s.material = m; s.light = light_id[i];
    s.has_normals = triHasN ? 1 : 0; s.pad0 = s.pad1 = s.pad2 = 0;
    out[i] = s;
    const int t = (m >= 0 && m < n_mats) ? mats[m].type : -1;
    bucket[i] = (uint8_t)((t >= 0 && t <= 6) ? kBucketMat0 + t : kBucketNone);
}

__global__ void bucketMaskKernel(const uint8_t* __restrict__ bucket, int64_t n, uint32_t* mask) {
    uint32_t m = 0u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m |= 1u << bucket[i];
    for (int d = 16; d >= 1; d >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, d);
    if ((threadIdx.x & 31) == 0 && m) atomicOr(mask, m);
}

static int uploadScene(spb_ctx* ctx, RenderState* R) {
    freeScene(R);
    const int64_t n = ctx->n_tris;
    R->n_tris = n;
    const bool hasN = !ctx->geo->normals.empty(), hasUV = !ctx->geo->uvs.empty();
    cudaStream_t st = ctx->stream;
    const size_t nm0 = std::max<size_t>(R->mats.size(), 1);
    SPB_CUDA(ctx, cudaMalloc(&R->d_mats, nm0 * sizeof(spb_material)));
    if (!R->mats.empty()) SPB_CUDA(ctx, cudaMemcpyAsync(R->d_mats, R->mats.data(), R->mats.size() * sizeof(spb_material), cudaMemcpyHostToDevice, st));
    R->bucket_mask = 1u << kBucketMiss;
    if (n > 0) {
        double* d_v = nullptr; int32_t *d_m = nullptr, *d_l = nullptr; uint32_t* d_mask = nullptr; float* d_uv = nullptr;
        auto freeTmp = [&]() { cudaFree(d_v); cudaFree(d_m); cudaFree(d_l); cudaFree(d_mask); cudaFree(d_uv); };
#define US(call) do { const cudaError_t e_ = (call); if (e_ != cudaSuccess) { freeTmp(); cudaOk(ctx, e_, #call); return e_ == cudaErrorMemoryAllocation ? SPB_ERR_OOM : SPB_ERR_CUDA; } } while (0)
        US(cudaMalloc(&R->d_tris, (size_t)n * sizeof(ShadeTri)));
        US(cudaMalloc(&R->d_prim_bucket, (size_t)n));
        US(cudaMalloc(&d_v, (size_t)n * 9 * sizeof(double)));
        US(cudaMalloc(&d_m, (size_t)n * 4)); US(cudaMalloc(&d_l, (size_t)n * 4)); US(cudaMalloc(&d_mask, 4));
        US(cudaMemcpyAsync(d_v, ctx->geo->verts.data(), (size_t)n * 9 * sizeof(double), cudaMemcpyHostToDevice, st));
        US(cudaMemcpyAsync(d_m, ctx->material_id.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        US(cudaMemcpyAsync(d_l, ctx->light_id.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        US(cudaMemsetAsync(d_mask, 0, 4, st));
        if (hasN) {
            US(cudaMalloc(&R->d_vnormals, (size_t)n * 9 * sizeof(float)));
            US(cudaMemcpyAsync(R->d_vnormals, ctx->geo->normals.data(), (size_t)n * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
        }
        if (hasUV) {
            US(cudaMalloc(&d_uv, (size_t)n * 6 * sizeof(float)));
            US(cudaMemcpyAsync(d_uv, ctx->geo->uvs.data(), (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice, st));
        }
        shadeTriKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_v, R->d_vnormals, d_uv, d_m, d_l, n, R->d_mats, (int)R->mats.size(), R->d_tris, R->d_prim_bucket);
        bucketMaskKernel<<<ctx->sm_count * 4, 256, 0, st>>>(R->d_prim_bucket, n, d_mask);
        uint32_t mask = 0u;
        US(cudaMemcpyAsync(&mask, d_mask, 4, cudaMemcpyDeviceToHost, st));
        US(cudaStreamSynchronize(st));
        US(cudaGetLastError());
#undef US
        freeTmp();
        R->bucket_mask |= mask;
        R->launches += 2;
    }
    const size_t nl = std::max<size_t>(R->lights.size(), 1);
    SPB_CUDA(ctx, cudaMalloc(&R->d_lights, nl * sizeof(spb_light)));
    if (!R->lights.empty()) SPB_CUDA(ctx, cudaMemcpy(R->d_lights, R->lights.data(), R->lights.size() * sizeof(spb_light), cudaMemcpyHostToDevice));
    R->ds.tris = R->d_tris; R->ds.vnormals = R->d_vnormals; R->ds.mats = R->d_mats; R->ds.lights = R->d_lights;
    R->ds.n_mats = (int)R->mats.size(); R->ds.n_lights = (int)R->lights.size();
    R->ds.mat_tex = nullptr; R->ds.texs = nullptr; R->ds.texels = nullptr; R->ds.uvs = nullptr;
    bool anyTex = false;
    for (int32_t t : R->mat_tex) anyTex |= (t >= 0);
    if (anyTex && n > 0) {
        if (R->mat_tex.size() != R->mats.size() * 2) return fail(ctx, SPB_ERR_INVALID, "texture bindings do not match the material list");
        for (int32_t t : R->mat_tex) if (t >= (int32_t)R->texs.size()) return fail(ctx, SPB_ERR_INVALID, "a material refers to a texture that was not set");
        for (const spb_texture& t : R->texs)
            if (t.type == SPB_TEX_BITMAP && (t.width <= 0 || t.height <= 0 || t.texel_offset < 0 ||
                                             (size_t)(t.texel_offset + (int64_t)t.width * t.height) * 3 > R->texels.size()))
                return fail(ctx, SPB_ERR_INVALID, "bitmap texture outside the texel array");
        std::vector<float> uv((size_t)n * 6, 0.f);
        if (!ctx->geo->uvs.empty()) uv = ctx->geo->uvs;
        std::vector<float4> tx(std::max<size_t>(R->texels.size() / 3, 1));
        for (size_t i = 0; i < R->texels.size() / 3; i++) tx[i] = make_float4(R->texels[i * 3], R->texels[i * 3 + 1], R->texels[i * 3 + 2], 0.f);
        SPB_CUDA(ctx, cudaMalloc(&R->d_mat_tex, R->mat_tex.size() * sizeof(int32_t)));
        SPB_CUDA(ctx, cudaMalloc(&R->d_texs, R->texs.size() * sizeof(spb_texture)));
        SPB_CUDA(ctx, cudaMalloc(&R->d_texels, tx.size() * sizeof(float4)));
        SPB_CUDA(ctx, cudaMalloc(&R->d_uvs, uv.size() * sizeof(float)));
        SPB_CUDA(ctx, cudaMemcpy(R->d_mat_tex, R->mat_tex.data(), R->mat_tex.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        SPB_CUDA(ctx, cudaMemcpy(R->d_texs, R->texs.data(), R->texs.size() * sizeof(spb_texture), cudaMemcpyHostToDevice));
        SPB_CUDA(ctx, cudaMemcpy(R->d_texels, tx.data(), tx.size() * sizeof(float4), cudaMemcpyHostToDevice));
        SPB_CUDA(ctx, cudaMemcpy(R->d_uvs, uv.data(), uv.size() * sizeof(float), cudaMemcpyHostToDevice));
        R->ds.mat_tex = R->d_mat_tex; R->ds.texs = R->d_texs; R->ds.texels = R->d_texels; R->ds.uvs = R->d_uvs;
    }
    // every lobe type a material can turn into, and the shading bucket of every triangle
    R->type_mask = 0u;
    for (const spb_material& m : R->mats) if (m.type >= 0 && m.type <= 6) R->type_mask |= typeClosure(m.type);
    // a scene of Lambertian surfaces only (the diffuse Cornell boxes) is shaded in queue order by one instance
    R->sort_materials = (R->bucket_mask & ~((1u << kBucketMiss) | (1u << (kBucketMat0 + SPB_MAT_DIFFUSE)))) != 0u;
    R->scene_dirty = false;
    return SPB_OK;
}

// 3x3 inverse (row-major) for the environment's worldToLight (core/light.cc:10)
static bool inv3(const double* m, double* o) {
    const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (det == 0.0) return false;
    const double id = 1.0 / det;
    o[0] = (m[4] * m[8] - m[5] * m[7]) * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = (m[5] * m[6] - m[3] * m[8]) * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = (m[3] * m[7] - m[4] * m[6]) * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    return true;
}

// Envmap constructor (lights/envmap.cc:17-47): scaled texels, MipMap level 0/1, sin-weighted
// luminance, Distribution2D (core/sampling.cc:25-47,100-118); all in double, narrowed to float32.
static int uploadEnv(spb_ctx* ctx, RenderState* R) {
    freeEnv(R);
    std::memset(&R->ds.env, 0, sizeof(R->ds.env));
    R->env_dirty = false;
    if (!R->env_present) return SPB_OK;
    const int w = R->env_w, h = R->env_h;
    std::vector<double> img((size_t)w * h * 3);
    for (size_t i = 0; i < img.size(); i++) img[i] = (double)R->env_rgb[i] * R->env_scale;
    auto texel = [&](const std::vector<double>& im, int iw, int ih, int s, int t, int c) {
        s = (s % iw + iw) % iw; t = (t % ih + ih) % ih;
        return im[((size_t)t * iw + s) * 3 + c];
    };
    auto bilinear = [&](const std::vector<double>& im, int iw, int ih, double s0, double t0, double* out) {
        const double s = s0 * iw - 0.5, t = t0 * ih - 0.5;
        const int si = (int)s, ti = (int)t;
        const double ds = s - si, dt = t - ti;
        for (int c = 0; c < 3; c++)
            out[c] = (1 - ds) * (1 - dt) * texel(im, iw, ih, si, ti, c) + ds * (1 - dt) * texel(im, iw, ih, si + 1, ti, c) +
                     (1 - ds) * dt * texel(im, iw, ih, si, ti + 1, c) + ds * dt * texel(im, iw, ih, si + 1, ti + 1, c);
    };
    // mip level 1 (core/mipmap.cc:52-64)
    const int w1 = std::max(1, w / 2), h1 = std::max(1, h / 2);
    std::vector<double> img1((size_t)w1 * h1 * 3);
    for (int y = 0; y < h1; y++) for (int x = 0; x < w1; x++) for (int c = 0; c < 3; c++)
        img1[((size_t)y * w1 + x) * 3 + c] = 0.25 * (texel(img, w, h, x * 2, y * 2, c) + texel(img, w, h, x * 2 + 1, y * 2, c) +
                                                     texel(img, w, h, x * 2, y * 2 + 1, c) + texel(img, w, h, x * 2 + 1, y * 2 + 1, c));
    auto roundUpPow2 = [](int v) { int p = 1; while (p < v) p <<= 1; return p; };
    auto isPow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    int rx = w, ry = h;
    if (!isPow2(w) || !isPow2(h)) { rx = roundUpPow2(w); ry = roundUpPow2(h); }
    int nLevels = 1; { int m = std::max(rx, ry); while (m > 1) { m >>= 1; nLevels++; } }
    const double filt = 1.0 / std::max(w, h);
    const double level = nLevels - 1 + std::log2(std::max(filt, 1.0e-8));
    std::vector<double> gray((size_t)w * h);
    for (int v = 0; v < h; v++) {
        const double vp = (v + 0.5) / h, sinT = std::sin(3.14159265358979323846 * (v + 0.5) / h);
        for (int u = 0; u < w; u++) {
            const double up = (u + 0.5) / w;
            double c[3];
            if (level < 0 || nLevels == 1) bilinear(img, w, h, up, vp, c);
            else if (level >= nLevels - 1) { for (int k = 0; k < 3; k++) c[k] = img1[k]; }
            else {
                const int l = (int)level; const double delta = level - l;
                double c0[3], c1[3];
                if (l == 0) { bilinear(img, w, h, up, vp, c0); bilinear(img1, w1, h1, up, vp, c1); }
                else { bilinear(img1, w1, h1, up, vp, c0); bilinear(img1, w1, h1, up, vp, c1); }
                for (int k = 0; k < 3; k++) c[k] = (1 - delta) * c0[k] + delta * c1[k];
            }
            gray[(size_t)v * w + u] = (0.2126 * c[0] + 0.7152 * c[1] + 0.0722 * c[2]) * sinT;
        }
    }
    // distributions
    std::vector<float> fl;
    const size_t oFunc = 0, oCdf = oFunc + (size_t)w * h, oInt = oCdf + (size_t)h * (w + 1), oMF = oInt + h, oMC = oMF + h, total = oMC + h + 1;
    fl.resize(total);
    std::vector<double> cdf((size_t)w + 1), marg((size_t)h);
    for (int v = 0; v < h; v++) {
        cdf[0] = 0.0;
        for (int i = 1; i <= w; i++) cdf[i] = cdf[i - 1] + gray[(size_t)v * w + i - 1] / w;
        const double integral = cdf[w];
        for (int i = 1; i <= w; i++) cdf[i] = integral == 0.0 ? (double)i / w : cdf[i] / integral;
        for (int i = 0; i < w; i++) fl[oFunc + (size_t)v * w + i] = (float)gray[(size_t)v * w + i];
        for (int i = 0; i <= w; i++) fl[oCdf + (size_t)v * (w + 1) + i] = (float)cdf[i];
        fl[oInt + v] = (float)integral; marg[v] = integral;
    }
    std::vector<double> mcdf((size_t)h + 1);
    mcdf[0] = 0.0;
    for (int i = 1; i <= h; i++) mcdf[i] = mcdf[i - 1] + marg[i - 1] / h;
    const double mint = mcdf[h];
    for (int i = 1; i <= h; i++) mcdf[i] = mint == 0.0 ? (double)i / h : mcdf[i] / mint;
    for (int i = 0; i < h; i++) fl[oMF + i] = (float)marg[i];
    for (int i = 0; i <= h; i++) fl[oMC + i] = (float)mcdf[i];
    std::vector<float4> tex((size_t)w * h);
    for (size_t i = 0; i < tex.size(); i++) tex[i] = make_float4((float)img[i * 3], (float)img[i * 3 + 1], (float)img[i * 3 + 2], 0.f);
    SPB_CUDA(ctx, cudaMalloc(&R->d_env_texels, tex.size() * sizeof(float4)));
    SPB_CUDA(ctx, cudaMalloc(&R->d_env_floats, fl.size() * sizeof(float)));
    SPB_CUDA(ctx, cudaMemcpy(R->d_env_texels, tex.data(), tex.size() * sizeof(float4), cudaMemcpyHostToDevice));
    SPB_CUDA(ctx, cudaMemcpy(R->d_env_floats, fl.data(), fl.size() * sizeof(float), cudaMemcpyHostToDevice));
    EnvMap& e = R->ds.env;
    e.texels = R->d_env_texels;
    e.condFunc = R->d_env_floats + oFunc; e.condCdf = R->d_env_floats + oCdf; e.condInt = R->d_env_floats + oInt;
    e.margFunc = R->d_env_floats + oMF; e.margCdf = R->d_env_floats + oMC; e.margInt = (float)mint;
    e.w = w; e.h = h; e.radius = (float)R->env_radius; e.present = 1;
    // lightToWorld_ = transpose(XML matrix) (envmap.cc:19); worldToLight_ = its inverse (light.cc:10)
    double l2w[9], w2l[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) l2w[i * 3 + j] = R->env_l2w[j * 4 + i];
    if (!inv3(l2w, w2l)) return fail(ctx, SPB_ERR_INVALID, "envmap toWorld matrix is singular");
    for (int i = 0; i < 9; i++) { e.l2w[i] = (float)l2w[i]; e.w2l[i] = (float)w2l[i]; }
    return SPB_OK;
}

static int allocQueues(spb_ctx* ctx, RenderState* R, int64_t slots) {
    if (R->d_pool && R->slots == slots) return SPB_OK;
    if (R->graph_exec) { cudaGraphExecDestroy(R->graph_exec); R->graph_exec = nullptr; }
    if (R->d_pool) { cudaStreamSynchronize(R->side); scratchFree(R->d_pool, ctx->stream); }
    R->d_pool = nullptr; R->slots = 0;
    // per entry: 2 extend queues x (32 B ray + 32 B state) + 16 B hit + 2 x (32 B ray + 16 B contribution) = 240 B
    const size_t per = 2 * 64 + 16 + 2 * 48;
    const size_t bytes = (size_t)slots * per + 4096;
    // from the process's scratch pool (context.h): a context created after another one went away re-uses its 8 GB
    cudaError_t e = scratchAlloc(ctx, &R->d_pool, bytes, ctx->stream);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(ctx, SPB_ERR_OOM, "out of device memory for the wavefront queues"); }
    SPB_CUDA(ctx, e);
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // (the side stream and the graph use it too)
    char* p = (char*)R->d_pool;
    auto take = [&](size_t b) { void* r = p; p += (b + 255) & ~(size_t)255; return r; };
    for (int k = 0; k < 2; k++) { R->q.ray[k] = (float4*)take((size_t)slots * 32); R->q.state[k] = (float4*)take((size_t)slots * 32); }
    R->q.shadow = (float4*)take((size_t)slots * 32); R->q.mis = (float4*)take((size_t)slots * 32);
    R->q.hit = (spb_hit*)take((size_t)slots * 16);
    R->q.shadowC = (float4*)take((size_t)slots * 16); R->q.misC = (float4*)take((size_t)slots * 16);
    R->q.count = (uint32_t*)take(64);
    R->slots = slots;
    return SPB_OK;
}

// the index lists of classifyKernel: one per bucket, each able to hold a whole queue
static int allocLists(spb_ctx* ctx, RenderState* R) {
    if (!R->sort_materials) return SPB_OK;
    if (R->d_lists && R->list_stride == R->slots) return SPB_OK;
    if (R->graph_exec) { cudaGraphExecDestroy(R->graph_exec); R->graph_exec = nullptr; }
    if (R->d_lists) { cudaStreamSynchronize(R->side); scratchFree(R->d_lists, ctx->stream); }
    R->d_lists = nullptr; R->list_stride = 0;
    cudaError_t e = scratchAlloc(ctx, (void**)&R->d_lists, (size_t)kNumBuckets * (size_t)R->slots * sizeof(uint32_t), ctx->stream);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(ctx, SPB_ERR_OOM, "out of device memory for the shading lists"); }
    SPB_CUDA(ctx, e);
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!R->d_list_count) { SPB_CUDA(ctx, cudaMalloc(&R->d_list_count, 16 * sizeof(uint32_t))); }
    SPB_CUDA(ctx, cudaMemset(R->d_list_count, 0, 16 * sizeof(uint32_t)));
    R->list_stride = R->slots;
    return SPB_OK;
}

static int scratch(spb_ctx* ctx, RenderState* R, size_t bytes, void** out) {
    if (R->scratch_bytes < bytes) {
        if (R->d_scratch) cudaFree(R->d_scratch);
        R->d_scratch = nullptr; R->scratch_bytes = 0;
        SPB_CUDA(ctx, cudaMalloc(&R->d_scratch, bytes));
        R->scratch_bytes = bytes;
    }
    *out = R->d_scratch;
    return SPB_OK;
}

// ---- one iteration of the streaming loop on queue `cur` (enqueue only) -------------------------------------------------
template <uint32_t TYPES, int MINB>
static void launchShadeList(spb_ctx* ctx, RenderState* R, int cur, int bucket, cudaStream_t st) {
    shadeKernel<TYPES, MINB, true><<<ctx->sm_count * 8, 128, 0, st>>>(R->rp, R->ds, R->q, R->d_film, cur, R->d_lists + (size_t)bucket * (size_t)R->list_stride,
                                                                   R->d_list_count + bucket);
}
static int launchShade(spb_ctx* ctx, RenderState* R, int cur, cudaStream_t st, int* launches) {
    const int grid = ctx->sm_count * 8;
    *launches = 0;
    if (!R->sort_materials) {
        const int mb = ctx->opt_shade_minb > 0 ? ctx->opt_shade_minb : kShadeMinbDiffuse;
        if (mb == 4) shadeKernel<kTypesDiffuse, 4, false><<<grid, 128, 0, st>>>(R->rp, R->ds, R->q, R->d_film, cur, nullptr, nullptr);
        else if (mb == 6) shadeKernel<kTypesDiffuse, 6, false><<<grid, 128, 0, st>>>(R->rp, R->ds, R->q, R->d_film, cur, nullptr, nullptr);
        else shadeKernel<kTypesDiffuse, 5, false><<<grid, 128, 0, st>>>(R->rp, R->ds, R->q, R->d_film, cur, nullptr, nullptr);
        *launches = 1;
    } else {
        const ShadeLists L = {R->d_lists, R->d_list_count, (uint32_t)R->list_stride};
        classifyKernel<<<grid, 256, 0, st>>>(R->q, cur, R->d_prim_bucket, L);
        shadeMissKernel<<<grid, 256, 0, st>>>(R->rp, R->ds, R->q, R->d_film, cur, R->d_lists + (size_t)kBucketMiss * (size_t)R->list_stride, R->d_list_count + kBucketMiss);
        *launches = 2;
        const uint32_t bm = R->bucket_mask;
        auto has = [&](int b) { return (bm >> b) & 1u; };
        // surfaces without a material pass the path straight through: any instance does that (the Lambertian one is the smallest)
        if (has(kBucketNone)) { launchShadeList<kTypesDiffuse, 4>(ctx, R, cur, kBucketNone, st); ++*launches; }
        if (has(kBucketMat0 + SPB_MAT_DIFFUSE)) { launchShadeList<typeClosure(SPB_MAT_DIFFUSE), 4>(ctx, R, cur, kBucketMat0 + SPB_MAT_DIFFUSE, st); ++*launches; }
        if (has(kBucketMat0 + SPB_MAT_DIELECTRIC)) { launchShadeList<typeClosure(SPB_MAT_DIELECTRIC), 5>(ctx, R, cur, kBucketMat0 + SPB_MAT_DIELECTRIC, st); ++*launches; }
        if (has(kBucketMat0 + SPB_MAT_ROUGHCONDUCTOR)) { launchShadeList<typeClosure(SPB_MAT_ROUGHCONDUCTOR), 4>(ctx, R, cur, kBucketMat0 + SPB_MAT_ROUGHCONDUCTOR, st); ++*launches; }
        if (has(kBucketMat0 + SPB_MAT_ROUGHDIELECTRIC)) { launchShadeList<typeClosure(SPB_MAT_ROUGHDIELECTRIC), 4>(ctx, R, cur, kBucketMat0 + SPB_MAT_ROUGHDIELECTRIC, st); ++*launches; }
        if (has(kBucketMat0 + SPB_MAT_CONDUCTOR)) { launchShadeList<typeClosure(SPB_MAT_CONDUCTOR), 5>(ctx, R, cur, kBucketMat0 + SPB_MAT_CONDUCTOR, st); ++*launches; }
        if (has(kBucketMat0 + SPB_MAT_PLASTIC)) { launchShadeList<typeClosure(SPB_MAT_PLASTIC), 5>(ctx, R, cur, kBucketMat0 + SPB_MAT_PLASTIC, st); ++*launches; }
        if (has(kBucketMat0 + SPB_MAT_ROUGHPLASTIC)) { launchShadeList<typeClosure(SPB_MAT_ROUGHPLASTIC), 4>(ctx, R, cur, kBucketMat0 + SPB_MAT_ROUGHPLASTIC, st); ++*launches; }
    }
    SPB_CUDA(ctx, cudaGetLastError());
    return SPB_OK;
}

static int enqueueIteration(spb_ctx* ctx, RenderState* R, int cur, cudaStream_t st) {
    const int64_t cap = R->slots;
    int rc;
    controlKernel<<<1, 32, 0, st>>>(R->d_ctl, R->q, cur, R->d_cursor, R->d_status, R->sort_materials ? R->d_list_count : nullptr);
    generateKernel<<<ctx->sm_count * 8, 256, 0, st>>>(R->rp, R->cam, R->q, R->d_ctl, cur);
    SPB_CUDA(ctx, cudaGetLastError());
    // extend
    if ((rc = launchTrace<false>(ctx, (const spb_ray_f32*)R->q.ray[cur], cap, R->q.count + cur, HitOut{R->q.hit}, R->d_cursor, st, true))) return rc;
    int shadeLaunches = 0;
    if ((rc = launchShade(ctx, R, cur, st, &shadeLaunches))) return rc;
    R->launches_per_iteration = 5 + shadeLaunches;
    // connect + MIS (skipped by their own zero counts when empty).  They are independent of each other: the MIS launch (a
    // handful of rays, ~20 us of launch and tail) runs on a side stream next to the connect launch -- a fork and a join in
    // the captured graph.
    SPB_CUDA(ctx, cudaEventRecord(R->ev_fork, st));
    SPB_CUDA(ctx, cudaStreamWaitEvent(R->side, R->ev_fork, 0));
    // a BSDF-sampled MIS ray towards an area light needs the closest hit (is it that emitter's triangle? core/mis.cc:122-125);
    // towards the environment only whether anything is in the way (mis.cc:126-128): with no area light in the scene the whole
    // queue takes the any-hit kernel, and SinkMis's "expected primitive -1" test is the same test
    if (R->mis_any_hit) rc = launchTrace<true>(ctx, (const spb_ray_f32*)R->q.mis, cap, R->q.count + 3, SinkMis{R->q.mis, R->q.misC, R->d_film}, R->d_cursor + 8, R->side, true);
    else rc = launchTrace<false>(ctx, (const spb_ray_f32*)R->q.mis, cap, R->q.count + 3, SinkMis{R->q.mis, R->q.misC, R->d_film}, R->d_cursor + 8, R->side, true);
    if (rc) return rc;
    SPB_CUDA(ctx, cudaEventRecord(R->ev_join, R->side));
    if ((rc = launchTrace<true>(ctx, (const spb_ray_f32*)R->q.shadow, cap, R->q.count + 2, SinkShadow{R->q.shadow, R->q.shadowC, R->d_film}, R->d_cursor + 4, st, true))) return rc;
    SPB_CUDA(ctx, cudaStreamWaitEvent(st, R->ev_join, 0));
    return SPB_OK;
}

// Runs the loop for sample indices first, first + stride, ... (count of them); host thread of the caller or of the worker.
static int renderLoop(spb_ctx* ctx, RenderState* R, int32_t first, int32_t count, int32_t stride) {
    cudaSetDevice(ctx->device);
    if (!R->begun) return fail(ctx, SPB_ERR_INVALID, "spb_render_samples: the scene changed since spb_render_begin (call it again)");
    const int64_t npix = (int64_t)R->rp.width * R->rp.height;
    const unsigned long long total = (unsigned long long)npix * (unsigned long long)count;
    if (total == 0) return SPB_OK;
    cudaStream_t st = ctx->stream;
    SPB_CUDA(ctx, cudaEventRecord(R->ev_r0, st));
    loopBeginKernel<<<1, 1, 0, st>>>(R->d_ctl, R->q, first, stride, total, (uint32_t)R->slots, R->d_status);
    R->launches++;
    // every iteration before this one certainly has work (regeneration alone needs that many): no polling until then
    const int64_t surePairs = (int64_t)((total + (unsigned long long)R->slots - 1) / (unsigned long long)R->slots + 1) / 2;
    int rc;
    for (int64_t pair = 0;; pair++) {
        if (R->graph_exec) { SPB_CUDA(ctx, cudaGraphLaunch(R->graph_exec, st)); }
        else {
            if ((rc = enqueueIteration(ctx, R, 0, st))) return rc;
            if ((rc = enqueueIteration(ctx, R, 1, st))) return rc;
        }
        R->launches += 2 * R->launches_per_iteration;
        SPB_CUDA(ctx, cudaEventRecord(R->poll[pair & 3], st));
        if (pair >= surePairs && pair >= 1) {
            // the PREVIOUS pair has finished (this one keeps the GPU busy meanwhile); the flag only ever goes 0 -> 1 within a call
            SPB_CUDA(ctx, cudaEventSynchronize(R->poll[(pair - 1) & 3]));
            if (*(volatile uint32_t*)R->h_status) break;
        }
    }
    SPB_CUDA(ctx, cudaEventRecord(R->ev_r1, st));
    SPB_CUDA(ctx, cudaStreamSynchronize(st));
    SPB_CUDA(ctx, cudaGetLastError());
    if (!*(volatile uint32_t*)R->h_status) return fail(ctx, SPB_ERR_CUDA, "internal: the streaming loop ended with paths in flight");
    float ms = 0.f;
    SPB_CUDA(ctx, cudaEventElapsedTime(&ms, R->ev_r0, R->ev_r1));
    R->render_ms += ms;
    return SPB_OK;
}

// ---- the render worker --------------------------------------------------------------------------------------------------
static void workerMain(RenderWorker* w) {
    for (;;) {
        std::function<int()> task;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->stop || !w->tasks.empty(); });
            if (w->tasks.empty()) return;
            task = std::move(w->tasks.front());
            w->tasks.pop_front();
        }
        const int rc = task();
        {
            std::lock_guard<std::mutex> lk(w->mu);
            if (rc != SPB_OK && w->firstError == SPB_OK) w->firstError = rc;
            w->pending--;
        }
        w->cvDone.notify_all();
    }
}
static void workerSubmit(spb_ctx* ctx, RenderState* R, std::function<int()> fn) {
    if (!R->worker) { R->worker = new RenderWorker(); R->worker->th = std::thread(workerMain, R->worker); }
    RenderWorker* w = R->worker;
    { std::lock_guard<std::mutex> lk(w->mu); w->tasks.push_back(std::move(fn)); w->pending++; }
    w->cv.notify_one();
}
// Waits for everything handed to the worker; returns (and clears) the first error a task reported.
static int workerDrain(spb_ctx* ctx, RenderState* R) {
    if (!R || !R->worker) return SPB_OK;
    RenderWorker* w = R->worker;
    std::unique_lock<std::mutex> lk(w->mu);
    w->cvDone.wait(lk, [&] { return w->pending == 0; });
    const int rc = w->firstError;
    w->firstError = SPB_OK;
    return rc;
}
static void workerStop(RenderState* R) {
    if (!R->worker) return;
    RenderWorker* w = R->worker;
    { std::lock_guard<std::mutex> lk(w->mu); w->stop = true; }
    w->cv.notify_all();
    if (w->th.joinable()) w->th.join();
    delete w;
    R->worker = nullptr;
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_scene_set_materials(spb_ctx* ctx, const spb_material* mats, int32_t n) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    if (n < 0 || (n > 0 && !mats)) return fail(ctx, SPB_ERR_INVALID, "spb_scene_set_materials: bad arguments");
    RenderState* R = rs(ctx);
    workerDrain(ctx, R);
    R->mats.assign(mats, mats + n);
    R->mat_tex.clear();
    R->scene_dirty = true; R->begun = false;
    return SPB_OK;
}

int spb_scene_set_textures(spb_ctx* ctx, const spb_texture* texs, int32_t n, const float* texels_rgb, int64_t n_texels) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    if (n < 0 || (n > 0 && !texs) || n_texels < 0 || (n_texels > 0 && !texels_rgb)) return fail(ctx, SPB_ERR_INVALID, "spb_scene_set_textures: bad arguments");
    for (int32_t i = 0; i < n; i++)
        if (texs[i].type != SPB_TEX_BITMAP && texs[i].type != SPB_TEX_CHECKERBOARD) return fail(ctx, SPB_ERR_INVALID, "spb_scene_set_textures: unknown texture type");
    RenderState* R = rs(ctx);
    workerDrain(ctx, R);
    R->texs.assign(texs, texs + n);
    R->texels.assign(texels_rgb, texels_rgb + n_texels * 3);
    R->scene_dirty = true; R->begun = false;
    return SPB_OK;
}

int spb_scene_set_material_textures(spb_ctx* ctx, const int32_t* tex_ids, int32_t n_mats) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    RenderState* R = rs(ctx);
    workerDrain(ctx, R);
    if (!tex_ids) { R->mat_tex.clear(); R->scene_dirty = true; R->begun = false; return SPB_OK; }
    if (n_mats != (int32_t)R->mats.size()) return fail(ctx, SPB_ERR_INVALID, "spb_scene_set_material_textures: call spb_scene_set_materials first (counts differ)");
    R->mat_tex.assign(tex_ids, tex_ids + (size_t)n_mats * 2);
    R->scene_dirty = true; R->begun = false;
    return SPB_OK;
}

int spb_scene_set_lights(spb_ctx* ctx, const spb_light* lights, int32_t n) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    if (n < 0 || (n > 0 && !lights)) return fail(ctx, SPB_ERR_INVALID, "spb_scene_set_lights: bad arguments");
    RenderState* R = rs(ctx);
    workerDrain(ctx, R);
    R->lights.assign(lights, lights + n);
    R->scene_dirty = true; R->begun = false;
    return SPB_OK;
}

int spb_scene_set_envmap(spb_ctx* ctx, const float* rgb, int32_t w, int32_t h, const double l2w[16], double scale,
                         const double center[3], double radius) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    RenderState* R = rs(ctx);
    workerDrain(ctx, R);
    R->begun = false;
    if (!rgb) { R->env_present = false; R->env_dirty = true; return SPB_OK; }
    if (w <= 0 || h <= 0) return fail(ctx, SPB_ERR_INVALID, "spb_scene_set_envmap: bad size");
    R->env_rgb.assign(rgb, rgb + (size_t)w * h * 3);
    R->env_w = w; R->env_h = h; R->env_scale = scale; R->env_radius = radius;
    for (int i = 0; i < 16; i++) R->env_l2w[i] = l2w ? l2w[i] : (i % 5 == 0 ? 1.0 : 0.0);
    for (int i = 0; i < 3; i++) R->env_center[i] = center ? center[i] : 0.0;
    R->env_present = true; R->env_dirty = true;
    return SPB_OK;
}

int spb_render_begin(spb_ctx* ctx, const spb_render_desc* desc) {
    if (!ctx || !desc) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: NULL argument");
    if (desc->width <= 0 || desc->height <= 0 || desc->max_depth < 0) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: bad film size or depth");
    if (desc->max_depth > 4095) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: max_depth above 4095 (the bounce counter of a path has 12 bits)");
    if ((int64_t)desc->width * desc->height >= ((int64_t)1 << 31)) return fail(ctx, SPB_ERR_UNSUPPORTED, "spb_render_begin: more than 2^31 pixels");
    if (desc->integrator != SPB_INTEGRATOR_PATH && desc->integrator != SPB_INTEGRATOR_DIRECT) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: unknown integrator");
    if (!ctx->bvh_ready) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: no acceleration structure (call spb_bvh_build first)");
    cudaSetDevice(ctx->device);
    RenderState* R = rs(ctx);
    int rc = workerDrain(ctx, R);
    if (rc) return rc;
    R->begun = false;
    for (int32_t m : ctx->material_id)
        if (m >= (int32_t)R->mats.size()) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: a triangle refers to a material that was not set");
    for (size_t i = 0; i < R->lights.size(); i++) {
        const spb_light& l = R->lights[i];
        if (l.type == SPB_LIGHT_AREA && (l.prim < 0 || l.prim >= ctx->n_tris)) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: area light refers to a missing triangle");
        if (l.type == SPB_LIGHT_ENVMAP && !R->env_present) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: envmap light without spb_scene_set_envmap");
    }
    for (int32_t l : ctx->light_id)
        if (l >= (int32_t)R->lights.size()) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: a triangle refers to a light that was not set");
    R->mis_any_hit = !R->lights.empty();
    for (const spb_light& l : R->lights) if (l.type != SPB_LIGHT_ENVMAP) R->mis_any_hit = false;
    if (R->scene_dirty && (rc = uploadScene(ctx, R))) return rc;
    if (R->env_dirty || (R->env_present && !R->ds.env.present)) { if ((rc = uploadEnv(ctx, R))) return rc; }
    R->desc = *desc;
    RenderParamsPOD& rp = R->rp;
    rp.width = desc->width; rp.height = desc->height; rp.max_depth = desc->max_depth; rp.filter = desc->filter;
    rp.rr_start = desc->rr_start_bounce;
    rp.integrator = desc->integrator;
    rp.frx = (float)desc->filter_radius[0]; rp.fry = (float)desc->filter_radius[1];
    const double sigma = desc->filter_sigma != 0.0 ? desc->filter_sigma : 0.5;
    rp.fbeta = (float)(1.0 / sigma);
    rp.fexpx = (float)std::exp(-desc->filter_radius[0] * desc->filter_radius[0] / sigma);
    rp.fexpy = (float)std::exp(-desc->filter_radius[1] * desc->filter_radius[1] / sigma);
    rp.seed = desc->seed;
    std::memcpy(R->cam.r2c, desc->raster_to_camera, sizeof(double) * 16);
    std::memcpy(R->cam.c2w, desc->camera_to_world, sizeof(double) * 16);
    R->cam.lens_radius = desc->lens_radius; R->cam.focal = desc->focal_distance;
    const int64_t npix = (int64_t)desc->width * desc->height;
    cudaStream_t st = ctx->stream;
    if (R->film_pixels != npix) {
        if (R->d_film) cudaFree(R->d_film);
        R->d_film = nullptr; R->film_pixels = 0;
        if (R->graph_exec) { cudaGraphExecDestroy(R->graph_exec); R->graph_exec = nullptr; }
        SPB_CUDA(ctx, cudaMalloc(&R->d_film, (size_t)npix * sizeof(float4)));
        R->film_pixels = npix;
    }
    SPB_CUDA(ctx, cudaMemsetAsync(R->d_film, 0, (size_t)npix * sizeof(float4), st));
    if (!R->d_ctl) { SPB_CUDA(ctx, cudaMalloc(&R->d_ctl, sizeof(LoopCtl))); }
    if (!R->d_cursor) { SPB_CUDA(ctx, cudaMalloc(&R->d_cursor, 16 * sizeof(unsigned long long))); }
    if (!R->h_status) {
        SPB_CUDA(ctx, cudaHostAlloc(&R->h_status, 64, cudaHostAllocMapped));
        SPB_CUDA(ctx, cudaHostGetDevicePointer(&R->d_status, R->h_status, 0));
        for (cudaEvent_t& e : R->poll) SPB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        SPB_CUDA(ctx, cudaEventCreate(&R->ev_r0)); SPB_CUDA(ctx, cudaEventCreate(&R->ev_r1));
        SPB_CUDA(ctx, cudaStreamCreateWithFlags(&R->side, cudaStreamNonBlocking));
        SPB_CUDA(ctx, cudaEventCreateWithFlags(&R->ev_fork, cudaEventDisableTiming)); SPB_CUDA(ctx, cudaEventCreateWithFlags(&R->ev_join, cudaEventDisableTiming));
    }
    R->h_status[0] = 0u; R->h_status[1] = 0u;
    SPB_CUDA(ctx, cudaMemsetAsync(R->d_ctl, 0, sizeof(LoopCtl), st));
    SPB_CUDA(ctx, cudaMemsetAsync(R->d_cursor, 0, 16 * sizeof(unsigned long long), st));
    R->launches = 0; R->render_ms = 0.0; R->reduce_ms = 0.0;
    // The queues: one allocation sized by "wave_slots" (but never more than 64 samples of the whole image).
    const int64_t slots = std::max<int64_t>(1024, std::min<int64_t>(ctx->opt_wave_slots, npix * 64));
    if ((rc = allocQueues(ctx, R, slots))) return rc;
    if ((rc = allocLists(ctx, R))) return rc;
    SPB_CUDA(ctx, cudaMemsetAsync(R->q.count, 0, 64, st));
    // One empty pair of iterations outside the timed region: loads every kernel of the loop (with several contexts
    // driven from one process the driver serialises module loading across the GPUs) and fills the occupancy cache.
    R->begun = true;
    loopBeginKernel<<<1, 1, 0, st>>>(R->d_ctl, R->q, 0, 1, 0ull, (uint32_t)R->slots, R->d_status);
    if ((rc = enqueueIteration(ctx, R, 0, st)) || (rc = enqueueIteration(ctx, R, 1, st))) { R->begun = false; return rc; }
    SPB_CUDA(ctx, cudaStreamSynchronize(st));
    SPB_CUDA(ctx, cudaMemsetAsync(R->d_ctl, 0, sizeof(LoopCtl), st));
    // ... and the same pair captured into a graph: the host re-launches it until the pipeline is empty.
    // (every parameter of the kernels is fixed until the next spb_render_begin; the sample range of a call lives in LoopCtl)
    if (R->graph_exec) { cudaGraphExecDestroy(R->graph_exec); R->graph_exec = nullptr; }
    if (ctx->opt_render_graph) {
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            const int r0 = enqueueIteration(ctx, R, 0, st), r1 = r0 ? r0 : enqueueIteration(ctx, R, 1, st);
            e = cudaStreamEndCapture(st, &graph);
            if (e == cudaSuccess && r1 == SPB_OK) e = cudaGraphInstantiate(&R->graph_exec, graph, 0);
            else if (e == cudaSuccess) e = cudaErrorUnknown;
            if (graph) cudaGraphDestroy(graph);
        }
        if (e != cudaSuccess) { cudaGetLastError(); R->graph_exec = nullptr; }      // plain launches instead
    }
    SPB_CUDA(ctx, cudaStreamSynchronize(st));
    return SPB_OK;
}

static int renderArgs(spb_ctx* ctx, int32_t first, int32_t count, int32_t stride, const char* who) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    RenderState* R = ctx->render;
    if (!R || !R->begun) return fail(ctx, SPB_ERR_INVALID, std::string(who) + ": call spb_render_begin first (and again after any scene change)");
    if (count < 0 || stride <= 0 || first < 0) return fail(ctx, SPB_ERR_INVALID, std::string(who) + ": bad sample range");
    return SPB_OK;
}

int spb_render_samples(spb_ctx* ctx, int32_t first, int32_t count, int32_t stride) {
    int rc = renderArgs(ctx, first, count, stride, "spb_render_samples");
    if (rc) return rc;
    RenderState* R = ctx->render;
    if ((rc = workerDrain(ctx, R))) return rc;
    return renderLoop(ctx, R, first, count, stride);
}

int spb_render_samples_async(spb_ctx* ctx, int32_t first, int32_t count, int32_t stride) {
    int rc = renderArgs(ctx, first, count, stride, "spb_render_samples_async");
    if (rc) return rc;
    RenderState* R = ctx->render;
    workerSubmit(ctx, R, [ctx, R, first, count, stride]() { return renderLoop(ctx, R, first, count, stride); });
    return SPB_OK;
}

int spb_render_wait(spb_ctx* ctx) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    return workerDrain(ctx, ctx->render);
}

static int filmArgs(spb_ctx* ctx, const char* who) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, std::string(who) + ": ctx is NULL");
    RenderState* R = ctx->render;
    if (!R || !R->d_film || R->film_pixels <= 0) return fail(ctx, SPB_ERR_INVALID, std::string(who) + ": no film (call spb_render_begin)");
    return SPB_OK;
}

static int filmReady(spb_ctx* ctx, const char* who, RenderState** out) {
    RenderState* R = ctx->render;
    if (!R || !R->d_film || R->film_pixels <= 0) return fail(ctx, SPB_ERR_INVALID, std::string(who) + ": no film (call spb_render_begin)");
    const int rc = workerDrain(ctx, R);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    *out = R;
    return SPB_OK;
}

int spb_film_read(spb_ctx* ctx, float* rgbw) {
    if (!ctx || !rgbw) return fail(ctx, SPB_ERR_INVALID, "spb_film_read: NULL argument");
    RenderState* R; int rc;
    if ((rc = filmReady(ctx, "spb_film_read", &R))) return rc;
    SPB_CUDA(ctx, cudaMemcpyAsync(rgbw, R->d_film, (size_t)R->film_pixels * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPB_OK;
}

int spb_film_resolve(spb_ctx* ctx, float* rgb) {
    if (!ctx || !rgb) return fail(ctx, SPB_ERR_INVALID, "spb_film_resolve: NULL argument");
    RenderState* R; int rc;
    if ((rc = filmReady(ctx, "spb_film_resolve", &R))) return rc;
    void* d = nullptr;
    if ((rc = scratch(ctx, R, (size_t)R->film_pixels * 3 * sizeof(float), &d))) return rc;
    resolveKernel<<<(unsigned)((R->film_pixels + 255) / 256), 256, 0, ctx->stream>>>(R->d_film, (float*)d, R->film_pixels);
    R->launches++;
    SPB_CUDA(ctx, cudaMemcpyAsync(rgb, d, (size_t)R->film_pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPB_OK;
}

static int filmEncode(spb_ctx* ctx, void* host, size_t bytesPerPixel, double invGamma, const char* what) {
    if (!ctx || !host) return fail(ctx, SPB_ERR_INVALID, std::string(what) + ": NULL argument");
    RenderState* R; int rc;
    if ((rc = filmReady(ctx, what, &R))) return rc;
    void* d = nullptr;
    if ((rc = scratch(ctx, R, (size_t)R->film_pixels * bytesPerPixel, &d))) return rc;
    const unsigned grid = (unsigned)((R->film_pixels + 255) / 256);
    if (bytesPerPixel == 4) encodeRgbeKernel<<<grid, 256, 0, ctx->stream>>>(R->d_film, (uchar4*)d, R->film_pixels);
    else encodeLdrKernel<<<grid, 256, 0, ctx->stream>>>(R->d_film, (unsigned char*)d, invGamma, R->film_pixels);
    R->launches++;
    SPB_CUDA(ctx, cudaMemcpyAsync(host, d, (size_t)R->film_pixels * bytesPerPixel, cudaMemcpyDeviceToHost, ctx->stream));
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPB_OK;
}
int spb_film_resolve_rgbe(spb_ctx* ctx, uint8_t* rgbe) { return filmEncode(ctx, rgbe, 4, 0.0, "spb_film_resolve_rgbe"); }
int spb_film_resolve_ldr(spb_ctx* ctx, double gamma, uint8_t* rgb8) {
    if (!(gamma >= 1.0e-12)) return fail(ctx, SPB_ERR_INVALID, "spb_film_resolve_ldr: too small gamma (core/tmo.cc:54)");
    return filmEncode(ctx, rgb8, 3, 1.0 / gamma, "spb_film_resolve_ldr");
}

int spb_film_add(spb_ctx* ctx, const float* rgbw) {
    if (!ctx || !rgbw) return fail(ctx, SPB_ERR_INVALID, "spb_film_add: NULL argument");
    RenderState* R; int rc;
    if ((rc = filmReady(ctx, "spb_film_add", &R))) return rc;
    void* d = nullptr;
    if ((rc = scratch(ctx, R, (size_t)R->film_pixels * sizeof(float4), &d))) return rc;
    SPB_CUDA(ctx, cudaMemcpyAsync(d, rgbw, (size_t)R->film_pixels * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    filmAddKernel<<<(unsigned)((R->film_pixels + 255) / 256), 256, 0, ctx->stream>>>(R->d_film, (const float4*)d, R->film_pixels);
    R->launches++;
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPB_OK;
}

int spb_render_get_stats(spb_ctx* ctx, spb_render_stats* out) {
    if (!ctx || !out) return fail(ctx, SPB_ERR_INVALID, "spb_render_get_stats: NULL argument");
    RenderState* R = ctx->render;
    std::memset(out, 0, sizeof(*out));
    if (!R || !R->d_ctl) return SPB_OK;
    const int rc = workerDrain(ctx, R);
    if (rc) return rc;
    cudaSetDevice(ctx->device);
    LoopCtl h;
    SPB_CUDA(ctx, cudaMemcpy(&h, R->d_ctl, sizeof(h), cudaMemcpyDeviceToHost));
    out->paths = (int64_t)h.stats[0];
    out->rays_closest = (int64_t)h.stats[1]; out->rays_shadow = (int64_t)h.stats[2]; out->rays_mis = (int64_t)h.stats[3];
    out->kernel_launches = R->launches; out->render_ms = R->render_ms;
    out->reduce_ms = R->reduce_ms; out->iterations = (int64_t)h.iterations;
    return SPB_OK;
}

// ---- NCCL (loaded lazily: libnccl.so.2 is only needed by multi-GPU jobs) -----------------------------
typedef struct { char internal[128]; } spbNcclUniqueId;
typedef int (*ncclGetUniqueId_t)(spbNcclUniqueId*);
typedef int (*ncclCommInitRank_t)(void**, int, spbNcclUniqueId, int);
typedef int (*ncclAllReduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*ncclReduce_t)(const void*, void*, size_t, int, int, int, void*, cudaStream_t);
typedef int (*ncclCommGetAsyncError_t)(void*, int*);
typedef int (*ncclCommAbort_t)(void*);
typedef int (*ncclCommDestroy_t)(void*);
typedef const char* (*ncclGetErrorString_t)(int);

static void* ncclLib() {
    static void* lib = nullptr;
    if (lib) return lib;
    // A communicator that carries one 33-133 MB reduce per frame needs neither NVLS multicast groups nor 32 channels, and
    // setting them up is most of its bring-up time: 8 GPUs in one process, 2.3-2.5 s with NCCL's defaults, 1.3-1.6 s with
    // these (profiles/r02q_cli_phases_g8.txt; the reduce itself takes 0.2 ms either way).  Only defaults: variables the user
    // exported win, and a process whose NCCL is already initialised (bench.py under torchrun) is not affected.
    setenv("NCCL_NVLS_ENABLE", "0", 0);
    setenv("NCCL_MAX_NCHANNELS", "8", 0);
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nme : names) { lib = dlopen(nme, RTLD_NOW | RTLD_LOCAL); if (lib) break; }
    return lib;
}

int spb_comm_get_unique_id(char id[SPB_COMM_ID_BYTES]) {
    void* lib = ncclLib();
    if (!lib) return fail(nullptr, SPB_ERR_UNSUPPORTED, "libnccl.so.2 not found");
    ncclGetUniqueId_t f = (ncclGetUniqueId_t)dlsym(lib, "ncclGetUniqueId");
    if (!f) return fail(nullptr, SPB_ERR_UNSUPPORTED, "ncclGetUniqueId not found");
    spbNcclUniqueId u;
    const int rc = f(&u);
    if (rc != 0) return fail(nullptr, SPB_ERR_CUDA, "ncclGetUniqueId failed");
    std::memcpy(id, u.internal, SPB_COMM_ID_BYTES);
    return SPB_OK;
}

int spb_comm_init(spb_ctx* ctx, const char id[SPB_COMM_ID_BYTES], int32_t n_ranks, int32_t rank) {
    if (!ctx || !id) return fail(ctx, SPB_ERR_INVALID, "spb_comm_init: NULL argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(ctx, SPB_ERR_INVALID, "spb_comm_init: bad rank");
    void* lib = ncclLib();
    if (!lib) return fail(ctx, SPB_ERR_UNSUPPORTED, "libnccl.so.2 not found");
    RenderState* R = rs(ctx);
    R->nccl_lib = lib;
    cudaSetDevice(ctx->device);
    ncclCommInitRank_t f = (ncclCommInitRank_t)dlsym(lib, "ncclCommInitRank");
    if (!f) return fail(ctx, SPB_ERR_UNSUPPORTED, "ncclCommInitRank not found");
    spbNcclUniqueId u;
    std::memcpy(u.internal, id, SPB_COMM_ID_BYTES);
    const int rc = f(&R->comm, n_ranks, u, rank);
    if (rc != 0) {
        ncclGetErrorString_t es = (ncclGetErrorString_t)dlsym(lib, "ncclGetErrorString");
        return fail(ctx, SPB_ERR_CUDA, std::string("ncclCommInitRank: ") + (es ? es(rc) : "error"));
    }
    // NCCL connects its channels lazily inside the first collective (tens of ms over 8 GPUs): do that here, as
    // part of communicator set-up, so that the one all-reduce of a frame costs what the transfer costs
    // (on a stream of its own: a host may bring the communicator up on one thread while another thread of the same context is
    // inside spb_render_begin, whose stream is being captured into the loop's graph)
    ncclAllReduce_t ar = (ncclAllReduce_t)dlsym(lib, "ncclAllReduce");
    if (ar) {
        cudaStream_t ws = nullptr;
        float* d_warm = nullptr;
        SPB_CUDA(ctx, cudaStreamCreateWithFlags(&ws, cudaStreamNonBlocking));
        SPB_CUDA(ctx, cudaMalloc(&d_warm, 1024 * sizeof(float)));
        SPB_CUDA(ctx, cudaMemsetAsync(d_warm, 0, 1024 * sizeof(float), ws));
        const int wrc = ar(d_warm, d_warm, 1024, 7, 0, R->comm, ws);
        SPB_CUDA(ctx, cudaStreamSynchronize(ws));
        cudaFree(d_warm);
        cudaStreamDestroy(ws);
        if (wrc != 0) return fail(ctx, SPB_ERR_CUDA, "ncclAllReduce (communicator warm-up) failed");
    }
    return SPB_OK;
}

// K7: one sum over the RGBW film per frame (ncclFloat32 = 7, ncclSum = 0), stream-ordered behind the last kernel of
// the frame.  root < 0: ncclAllReduce (every rank ends up with the frame); root >= 0: ncclReduce (only `root` does --
// half the traffic, and all a job that writes one image needs).  After the enqueue and again after the stream has
// drained the communicator is asked for asynchronous errors (a peer that died, a transport fault): a film that was
// not fully reduced must not be mistaken for a frame.
static int ncclAsyncCheck(spb_ctx* ctx, RenderState* R, const char* where) {
    ncclCommGetAsyncError_t ge = (ncclCommGetAsyncError_t)dlsym(R->nccl_lib, "ncclCommGetAsyncError");
    if (!ge) return SPB_OK;
    int async = 0;
    const int rc = ge(R->comm, &async);
    if (rc == 0 && (async == 0 || async == 7)) return SPB_OK;          // ncclSuccess, ncclInProgress
    ncclGetErrorString_t es = (ncclGetErrorString_t)dlsym(R->nccl_lib, "ncclGetErrorString");
    ncclCommAbort_t ab = (ncclCommAbort_t)dlsym(R->nccl_lib, "ncclCommAbort");
    const int code = rc != 0 ? rc : async;
    const std::string msg = std::string("NCCL asynchronous error ") + where + ": " + (es ? es(code) : "error");
    if (ab) { ab(R->comm); R->comm = nullptr; }                         // the communicator is unusable from here on
    return fail(ctx, SPB_ERR_CUDA, msg);
}

static int filmReduceNow(spb_ctx* ctx, RenderState* R, int32_t root) {
    cudaSetDevice(ctx->device);
    if (!R->comm) return fail(ctx, SPB_ERR_INVALID, "spb_film_reduce: call spb_comm_init first");
    cudaStream_t st = ctx->stream;
    SPB_CUDA(ctx, cudaEventRecord(R->ev_r0, st));
    int rc;
    if (root < 0) {
        ncclAllReduce_t f = (ncclAllReduce_t)dlsym(R->nccl_lib, "ncclAllReduce");
        if (!f) return fail(ctx, SPB_ERR_UNSUPPORTED, "ncclAllReduce not found");
        rc = f(R->d_film, R->d_film, (size_t)R->film_pixels * 4, 7, 0, R->comm, st);
    } else {
        ncclReduce_t f = (ncclReduce_t)dlsym(R->nccl_lib, "ncclReduce");
        if (!f) return fail(ctx, SPB_ERR_UNSUPPORTED, "ncclReduce not found");
        rc = f(R->d_film, R->d_film, (size_t)R->film_pixels * 4, 7, 0, root, R->comm, st);
    }
    if (rc != 0) {
        ncclGetErrorString_t es = (ncclGetErrorString_t)dlsym(R->nccl_lib, "ncclGetErrorString");
        return fail(ctx, SPB_ERR_CUDA, std::string(root < 0 ? "ncclAllReduce: " : "ncclReduce: ") + (es ? es(rc) : "error"));
    }
    SPB_CUDA(ctx, cudaEventRecord(R->ev_r1, st));
    if ((rc = ncclAsyncCheck(ctx, R, "after the film reduce was enqueued"))) return rc;
    // wait with a watchdog on the communicator instead of a blind synchronize: a dead peer would hang it for ever.  The
    // communicator is only asked every 5 ms: ncclCommGetAsyncError in a tight loop competes with NCCL's own proxy thread
    // (r02j: 8.9 ms for the 33 MB reduce of C3 on 8 GPUs with a check per spin).
    auto lastCheck = std::chrono::steady_clock::now();
    for (;;) {
        const cudaError_t q = cudaEventQuery(R->ev_r1);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) { cudaOk(ctx, q, "cudaEventQuery(film reduce)"); return SPB_ERR_CUDA; }
        const auto now = std::chrono::steady_clock::now();
        if (now - lastCheck > std::chrono::milliseconds(5)) {
            lastCheck = now;
            if ((rc = ncclAsyncCheck(ctx, R, "while the film reduce was running"))) return rc;
        }
        std::this_thread::yield();
    }
    if ((rc = ncclAsyncCheck(ctx, R, "after the film reduce"))) return rc;
    float ms = 0.f;
    SPB_CUDA(ctx, cudaEventElapsedTime(&ms, R->ev_r0, R->ev_r1));
    R->reduce_ms += ms;
    return SPB_OK;
}

static int reduceArgs(spb_ctx* ctx, const char* who) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    RenderState* R = ctx->render;
    if (!R || !R->d_film) return fail(ctx, SPB_ERR_INVALID, std::string(who) + ": no film");
    if (!R->comm) return fail(ctx, SPB_ERR_INVALID, std::string(who) + ": call spb_comm_init first");
    return SPB_OK;
}

int spb_film_reduce(spb_ctx* ctx, int32_t root) {
    int rc = reduceArgs(ctx, "spb_film_reduce");
    if (rc) return rc;
    if ((rc = workerDrain(ctx, ctx->render))) return rc;
    return filmReduceNow(ctx, ctx->render, root);
}
int spb_film_reduce_async(spb_ctx* ctx, int32_t root) {
    const int rc = reduceArgs(ctx, "spb_film_reduce_async");
    if (rc) return rc;
    RenderState* R = ctx->render;
    workerSubmit(ctx, R, [ctx, R, root]() { return filmReduceNow(ctx, R, root); });
    return SPB_OK;
}
int spb_film_allreduce(spb_ctx* ctx) { return spb_film_reduce(ctx, -1); }

// ---- K7b: the film sum of one process driving several GPUs, over peer memory -----------------------------------
// One kernel on the root's GPU: every thread keeps one RGBW texel of every other film in flight (loads through NVLink /
// NVSwitch peer mappings; plain loads when the other context sits on the same GPU) and adds them to its own.  33 MB x 7
// peers at 1080p = 0.23 GB into one GPU's NVLink ingress; no communicator to bring up or tear down.
struct PeerFilms { const float4* p[15]; int n; };
__global__ void filmPeerSumKernel(float4* __restrict__ film, PeerFilms pf, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v[15];
#pragma unroll
        for (int k = 0; k < 15; k++) if (k < pf.n) v[k] = __ldcs(pf.p[k] + i);
        float4 a = film[i];
#pragma unroll
        for (int k = 0; k < 15; k++) if (k < pf.n) { a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w; }
        film[i] = a;
    }
}

int spb_film_reduce_peers(spb_ctx* root, spb_ctx* const* others, int32_t n_others) {
    if (!root) return fail(nullptr, SPB_ERR_INVALID, "spb_film_reduce_peers: root is NULL");
    RenderState* R = root->render;
    if (!R || !R->d_film || R->film_pixels <= 0) return fail(root, SPB_ERR_INVALID, "spb_film_reduce_peers: no film (call spb_render_begin)");
    if (n_others < 0 || (n_others > 0 && !others)) return fail(root, SPB_ERR_INVALID, "spb_film_reduce_peers: bad argument");
    for (int k = 0; k < n_others; k++) {
        spb_ctx* o = others[k];
        if (!o || o == root) return fail(root, SPB_ERR_INVALID, "spb_film_reduce_peers: others[] holds NULL or the root itself");
        if (!o->render || !o->render->d_film || o->render->film_pixels != R->film_pixels)
            return fail(root, SPB_ERR_INVALID, "spb_film_reduce_peers: every context needs a film of the root's size");
        for (int j = 0; j < k; j++) if (others[j] == o) return fail(root, SPB_ERR_INVALID, "spb_film_reduce_peers: a context is listed twice");
    }
    // everything queued on any of the contexts is finished first (the other films are only read here)
    int rc = workerDrain(root, R);
    if (rc) return rc;
    for (int k = 0; k < n_others; k++) {
        spb_ctx* o = others[k];
        if ((rc = workerDrain(o, o->render))) return fail(root, rc, std::string("spb_film_reduce_peers: a queued call of context ") + std::to_string(k) + " failed: " + o->err);
        cudaSetDevice(o->device);
        SPB_CUDA(o, cudaStreamSynchronize(o->stream));
    }
    cudaSetDevice(root->device);
    cudaStream_t st = root->stream;
    SPB_CUDA(root, cudaEventRecord(R->ev_r0, st));
    const int64_t n = R->film_pixels;
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)root->sm_count * 8);
    PeerFilms pf; pf.n = 0;
    auto flush = [&]() {
        if (pf.n > 0) filmPeerSumKernel<<<grid, 256, 0, st>>>(R->d_film, pf, n);
        pf.n = 0;
    };
    for (int k = 0; k < n_others; k++) {
        spb_ctx* o = others[k];
        bool direct = o->device == root->device;
        if (!direct) {
            int can = 0;
            SPB_CUDA(root, cudaDeviceCanAccessPeer(&can, root->device, o->device));
            if (can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else SPB_CUDA(root, e);
                direct = true;
            }
        }
        if (direct) {
            pf.p[pf.n++] = o->render->d_film;
            if (pf.n == 15) flush();
        } else {
            // not peers: the other film is staged in the root's scratch buffer and added from there
            void* d = nullptr;
            if ((rc = scratch(root, R, (size_t)n * sizeof(float4), &d))) return rc;
            SPB_CUDA(root, cudaMemcpyPeerAsync(d, root->device, o->render->d_film, o->device, (size_t)n * sizeof(float4), st));
            filmAddKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(R->d_film, (const float4*)d, n);
        }
    }
    flush();
    SPB_CUDA(root, cudaGetLastError());
    SPB_CUDA(root, cudaEventRecord(R->ev_r1, st));
    SPB_CUDA(root, cudaStreamSynchronize(st));
    float ms = 0.f;
    SPB_CUDA(root, cudaEventElapsedTime(&ms, R->ev_r0, R->ev_r1));
    R->reduce_ms += ms;
    return SPB_OK;
}

// ---- the same sum across PROCESSES (one per GPU, torchrun / MPI style): the other ranks' films mapped through CUDA IPC ----
// handle = cudaIpcMemHandle_t (64 B) | film pixels (int64) | magic
static const uint64_t kFilmHandleMagic = 0x53504246494c4d31ull;       // "SPBFILM1"
int spb_film_export_handle(spb_ctx* ctx, char handle[SPB_FILM_HANDLE_BYTES]) {
    int rc = filmArgs(ctx, "spb_film_export_handle");
    if (rc) return rc;
    if (!handle) return fail(ctx, SPB_ERR_INVALID, "spb_film_export_handle: handle is NULL");
    static_assert(sizeof(cudaIpcMemHandle_t) + 16 <= SPB_FILM_HANDLE_BYTES, "film handle size");
    RenderState* R = ctx->render;
    cudaSetDevice(ctx->device);
    cudaIpcMemHandle_t h;
    SPB_CUDA(ctx, cudaIpcGetMemHandle(&h, R->d_film));
    std::memset(handle, 0, SPB_FILM_HANDLE_BYTES);
    std::memcpy(handle, &h, sizeof(h));
    const int64_t px = R->film_pixels;
    std::memcpy(handle + sizeof(h), &px, 8);
    std::memcpy(handle + sizeof(h) + 8, &kFilmHandleMagic, 8);
    return SPB_OK;
}

int spb_film_import_handles(spb_ctx* root, const char* handles, int32_t n) {
    int rc = filmArgs(root, "spb_film_import_handles");
    if (rc) return rc;
    if (n < 0 || (n > 0 && !handles)) return fail(root, SPB_ERR_INVALID, "spb_film_import_handles: bad argument");
    RenderState* R = root->render;
    if ((rc = workerDrain(root, R))) return rc;
    cudaSetDevice(root->device);
    // all or nothing: every handle is checked before the first one is opened, and what was mapped before goes either way
    for (void* p : R->ipc_films) cudaIpcCloseMemHandle(p);
    R->ipc_films.clear();
    for (int k = 0; k < n; k++) {
        const char* hb = handles + (size_t)k * SPB_FILM_HANDLE_BYTES;
        int64_t px = 0; uint64_t magic = 0;
        std::memcpy(&px, hb + sizeof(cudaIpcMemHandle_t), 8);
        std::memcpy(&magic, hb + sizeof(cudaIpcMemHandle_t) + 8, 8);
        if (magic != kFilmHandleMagic) return fail(root, SPB_ERR_INVALID, "spb_film_import_handles: not a handle of spb_film_export_handle");
        if (px != R->film_pixels) return fail(root, SPB_ERR_INVALID, "spb_film_import_handles: every film needs the root's size");
    }
    for (int k = 0; k < n; k++) {
        const char* hb = handles + (size_t)k * SPB_FILM_HANDLE_BYTES;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, hb, sizeof(h));
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (void* q : R->ipc_films) cudaIpcCloseMemHandle(q);
            R->ipc_films.clear();
            return fail(root, SPB_ERR_UNSUPPORTED, std::string("spb_film_import_handles: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e) +
                                                   " (a film of this process, or GPUs that are not peers: use spb_film_reduce)");
        }
        R->ipc_films.push_back(p);
    }
    return SPB_OK;
}

int spb_film_reduce_imported(spb_ctx* root) {
    int rc = filmArgs(root, "spb_film_reduce_imported");
    if (rc) return rc;
    RenderState* R = root->render;
    if ((rc = workerDrain(root, R))) return rc;
    cudaSetDevice(root->device);
    cudaStream_t st = root->stream;
    SPB_CUDA(root, cudaEventRecord(R->ev_r0, st));
    const int64_t n = R->film_pixels;
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)root->sm_count * 8);
    PeerFilms pf; pf.n = 0;
    for (void* p : R->ipc_films) {
        pf.p[pf.n++] = (const float4*)p;
        if (pf.n == 15) { filmPeerSumKernel<<<grid, 256, 0, st>>>(R->d_film, pf, n); pf.n = 0; }
    }
    if (pf.n > 0) filmPeerSumKernel<<<grid, 256, 0, st>>>(R->d_film, pf, n);
    SPB_CUDA(root, cudaGetLastError());
    SPB_CUDA(root, cudaEventRecord(R->ev_r1, st));
    SPB_CUDA(root, cudaStreamSynchronize(st));
    float ms = 0.f;
    SPB_CUDA(root, cudaEventElapsedTime(&ms, R->ev_r0, R->ev_r1));
    R->reduce_ms += ms;
    return SPB_OK;
}

int spb_comm_destroy(spb_ctx* ctx) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    RenderState* R = ctx->render;
    if (!R || !R->comm) return SPB_OK;
    workerDrain(ctx, R);
    ncclCommDestroy_t f = (ncclCommDestroy_t)dlsym(R->nccl_lib, "ncclCommDestroy");
    if (f) f(R->comm);
    R->comm = nullptr;
    return SPB_OK;
}

}  // extern "C"
