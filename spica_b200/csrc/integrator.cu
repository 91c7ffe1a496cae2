// The unidirectional path tracer as a wavefront pipeline of sm_100a kernels.
//
// Replaces SamplerIntegrator::render (reference core/integrator.cc:46-110) and PathIntegrator::Li
// (integrators/path/path.cc:42-125) with uniformSampleOneLight / estimateDirectLight
// (core/mis.cc:21-136) for whole waves of paths:
//
//   K3 generate   camera rays for a wave of (pixel, sample) pairs          integrator.cc:80-86, perspective.cc:53-74
//   K1 extend     closest hit for the ray queue (trace_kernels.cuh)        scene.cc:37 -> bvh.cc:331-360
//   K4 shade      emission, NEE light sample + MIS BSDF sample, BSDF       path.cc:58-94, mis.cc:35-136
//                 sample for the next segment, Russian roulette, queue compaction
//   K2 connect    any-hit for the shadow queue; the unoccluded             visibility_tester.cc:21-24 -> bvh.cc:362-387
//                 contribution is added inside the traversal kernel
//   K1' mis       closest hit for the (rare) MIS rays; the emission is     mis.cc:113-130
//                 added inside the traversal kernel when the expected light is hit
//   K5 film       adds every finished path into the RGBW film              film.cc:65-74, integrator.cc:88-90
//
// Queue sizes never visit the host: every kernel reads its count from device memory, so a wave is
// one uninterrupted stream of launches.  A path keeps one slot for its whole life (beta, L, pixel,
// sampler key); queues carry (ray, slot) records.
#include <chrono>
#include <cstring>
#include <vector>

#include <dlfcn.h>

#include "context.h"
#include "shading.cuh"
#include "trace_kernels.cuh"

namespace spb {

// ---- device state -----------------------------------------------------------------------------------
struct PathSoA {
    float* beta[3];
    float* L[3];
    uint32_t* pixel;      // pixel index y * width + x (unflipped)
    uint32_t* key;        // sampler key of (pixel, sample)
    uint32_t* flags;      // [7:0] bounces, [8] specularBounce, [31:16] path vertex counter (sampler dimension base)
};

struct Queues {
    spb_ray_f32* ray[2];  // extend queue (double buffered); ray.tmin carries the slot id (bit pattern)
    spb_hit*     hit;     // parallel to ray[cur]
    spb_ray_f32* shadow;  // connect queue
    float4*      shadowC; // contribution rgb
    spb_ray_f32* mis;     // MIS closest-hit queue
    float4*      misC;    // contribution rgb, w = expected primitive id (bits) or -1 for "escapes"
    uint32_t*    count;   // [0],[1] extend queue sizes, [2] shadow, [3] mis
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

struct SinkShadow {      // connect: add the contribution when NOTHING was hit
    const spb_ray_f32* rays; const float4* contrib; PathSoA paths;
    __device__ __forceinline__ void store(int64_t i, const RayState& r) const {
        if (r.best_prim >= 0) return;
        const uint32_t slot = __float_as_uint(rays[i].tmin);
        const float4 c = contrib[i];
        paths.L[0][slot] += c.x; paths.L[1][slot] += c.y; paths.L[2][slot] += c.z;
    }
};
struct SinkMis {         // MIS: add the contribution when exactly the expected light (or nothing) was hit
    const spb_ray_f32* rays; const float4* contrib; PathSoA paths;
    __device__ __forceinline__ void store(int64_t i, const RayState& r) const {
        const float4 c = contrib[i];
        if (r.best_prim != (int32_t)__float_as_uint(c.w)) return;
        const uint32_t slot = __float_as_uint(rays[i].tmin);
        paths.L[0][slot] += c.x; paths.L[1][slot] += c.y; paths.L[2][slot] += c.z;
    }
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
    bool sort_materials = false;        // more than one BSDF type in the scene: shade in material order
    uint32_t type_mask = 0x7fu;         // lobe types the scene's materials can produce: picks the shade kernel instance
    // envmap
    std::vector<float> env_rgb; int env_w = 0, env_h = 0; double env_l2w[16]; double env_scale = 1.0, env_center[3] = {0, 0, 0}, env_radius = 2.0;
    bool env_present = false, env_dirty = false;
    float4* d_env_texels = nullptr; float* d_env_floats = nullptr;
    DeviceScene ds{};
    // film + paths
    float4* d_film = nullptr; int64_t film_pixels = 0;
    int64_t slots = 0;
    void* d_pool = nullptr;
    PathSoA paths{}; Queues q{};
    unsigned long long* d_stats = nullptr;   // [0] paths [1] closest [2] shadow [3] mis
    unsigned long long* d_cursor = nullptr;  // 3 x 4 u64 cursors for the three trace launches of a bounce
    int64_t launches = 0, paths_total = 0; double render_ms = 0.0;
    // nccl
    void* nccl_lib = nullptr; void* comm = nullptr;
};

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

__global__ void __launch_bounds__(256) generateKernel(RenderParamsPOD rp, CameraPOD cam, PathSoA paths, Queues q,
                                                    int64_t item0, int64_t n, int first, int stride) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t item = item0 + i;
    const int64_t npix = (int64_t)rp.width * rp.height;
    const uint32_t pixel = (uint32_t)(item % npix);
    const uint32_t sample = (uint32_t)(first + (int)(item / npix) * stride);
    const uint32_t key = samplerKey(rp.seed, pixel, sample);
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
    spb_ray_f32 r;
    r.ox = (float)ow[0]; r.oy = (float)ow[1]; r.oz = (float)ow[2];
    r.dx = (float)(dw[0] * s); r.dy = (float)(dw[1] * s); r.dz = (float)(dw[2] * s);
    r.tmin = __uint_as_float((uint32_t)i); r.tmax = kRayInf;
    q.ray[0][i] = r;
    paths.beta[0][i] = 1.f; paths.beta[1][i] = 1.f; paths.beta[2][i] = 1.f;
    paths.L[0][i] = 0.f; paths.L[1][i] = 0.f; paths.L[2][i] = 0.f;
    paths.pixel[i] = pixel; paths.key[i] = key; paths.flags[i] = 0u;
    if (i == 0) { q.count[0] = (uint32_t)n; q.count[1] = 0u; q.count[2] = 0u; q.count[3] = 0u; }
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

// warp-aggregated queue push: one atomic per warp per queue
__device__ __forceinline__ uint32_t queuePush(uint32_t* counter, bool want) {
    const unsigned mask = __ballot_sync(__activemask(), want);
    if (!want) return 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

__device__ __forceinline__ void writeRay(spb_ray_f32* q, uint32_t idx, V3 o, V3 d, uint32_t slot, float tmax) {
    float4* p = (float4*)(q + idx);
    p[0] = make_float4(o.x, o.y, o.z, d.x);
    p[1] = make_float4(d.y, d.z, __uint_as_float(slot), tmax);
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

// MINB: resident CTAs per SM the kernel is compiled for.  4 (128 registers, 300 B of spills) beats 3 (167
// registers, no spills) by 4 % on the diffuse Cornell box and ties on the glossy one; 5 and 6 lose 4-11 %
// (tools/sweep_shade.py on a measurement build).
// TYPES: the lobe types (bit t = SPB_MAT_t) that can occur in the scene.  The kernel is instantiated for the sets
// the BASELINE scenes need -- Lambertian only (C1, C3, C5's torus is rough dielectric: generic), the glossy Cornell
// box's {diffuse, dielectric, rough conductor, conductor} -- and for all seven; every lobe outside the set is
// compiled out (shading.cuh, lobeLive), which is what the registers of this kernel are spent on.
constexpr int kShadeMinbDiffuse = 4;      // measured on the diffuse Cornell box: 4 -> 1729, 5 -> 1687, 6 -> 1625 Msamples/s (generic kernel: 1462)
constexpr uint32_t kTypesAll = 0x7fu;
constexpr uint32_t kTypesDiffuse = 1u << SPB_MAT_DIFFUSE;
constexpr uint32_t kTypesGlossy = (1u << SPB_MAT_DIFFUSE) | (1u << SPB_MAT_DIELECTRIC) | (1u << SPB_MAT_ROUGHCONDUCTOR) | (1u << SPB_MAT_CONDUCTOR);

template <bool SORT, int MINB = 4, uint32_t TYPES = kTypesAll>
__global__ void __launch_bounds__(128, MINB) shadeKernel(RenderParamsPOD rp, DeviceScene sc, PathSoA paths, Queues q, int cur) {
    const uint32_t n = q.count[cur];
    const spb_ray_f32* rays = q.ray[cur];
    spb_ray_f32* nextQ = q.ray[cur ^ 1];
    uint32_t* nextCount = q.count + (cur ^ 1);
    // Material-sorted shading: the block takes a tile of kTile queue entries, buckets them by the BSDF type of
    // the surface that was hit (counting sort in shared memory, one warp-aggregated atomic per key and warp)
    // and shades them in bucket order, so a warp runs ONE material's code instead of the sum of all of them
    // (the microfacet lobes cost ~10x the Lambertian one).  SORT = false (every material has the same BSDF
    // type) shades in queue order.  Whole warps iterate together so that the warp-aggregated queue pushes see
    // converged lanes.
    constexpr uint32_t kTile = SORT ? 512 : 128, kPer = kTile / 128;
    __shared__ uint32_t s_cnt[16], s_start[16];
    __shared__ uint16_t s_order[kTile];
    const uint32_t nTiles = (n + kTile - 1) / kTile;
    for (uint32_t tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
      const uint32_t tileBase = tile * kTile;
      if (SORT) {
      if (threadIdx.x < 16) s_cnt[threadIdx.x] = 0u;
      __syncthreads();
      uint32_t keys[kPer], ranks[kPer];
#pragma unroll
      for (uint32_t k = 0; k < kPer; k++) {
          const uint32_t e = tileBase + k * 128u + threadIdx.x;
          uint32_t key = 15u;                                        // past the end of the queue: sorted last, skipped
          if (e < n) {
              const int prim = __float_as_int(((const float*)(q.hit + e))[1]);
              key = 0u;                                              // escaped rays
              if (prim >= 0) {
                  const int mat = __ldg((const int*)(sc.tris + prim) + 19);       // ShadeTri::material
                  key = (mat >= 0 && mat < sc.n_mats) ? (uint32_t)(sc.mats[mat].type + 2) & 15u : 1u;
              }
          }
          const unsigned peers = __match_any_sync(0xffffffffu, key);
          const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
          uint32_t r = 0u;
          if (lane == leader) r = atomicAdd(&s_cnt[key], (uint32_t)__popc(peers));
          r = __shfl_sync(0xffffffffu, r, leader);
          keys[k] = key; ranks[k] = r + (uint32_t)__popc(peers & ((1u << lane) - 1u));
      }
      __syncthreads();
      if (threadIdx.x == 0) { uint32_t run = 0u; for (int j = 0; j < 16; j++) { s_start[j] = run; run += s_cnt[j]; } }
      __syncthreads();
#pragma unroll
      for (uint32_t k = 0; k < kPer; k++) s_order[s_start[keys[k]] + ranks[k]] = (uint16_t)(k * 128u + threadIdx.x);
      __syncthreads();
      }
      for (uint32_t k = 0; k < kPer; k++) {
        const uint32_t i = tileBase + (SORT ? (uint32_t)s_order[k * 128u + threadIdx.x] : k * 128u + threadIdx.x);
        const bool valid = i < n;
        bool pushNext = false, pushShadow = false, pushMis = false;
        V3 nO = v3(0.f), nD = v3(0.f), sO = v3(0.f), sD = v3(0.f), mO = v3(0.f), mD = v3(0.f);
        float sT = 0.f;
        V3 sC = v3(0.f), mC = v3(0.f);
        int mExpect = -1;
        uint32_t slot = 0;
        if (valid) {
            const float4 r0 = __ldcs((const float4*)(rays + i)), r1 = __ldcs((const float4*)(rays + i) + 1);
            const float4 hv = __ldcs((const float4*)(q.hit + i));
            slot = __float_as_uint(r1.z);
            const V3 d = v3(r0.w, r1.x, r1.y);
            const int prim = __float_as_int(hv.y);
            uint32_t flags = paths.flags[slot];
            const int bounces = (int)(flags & 0xffu);
            const bool specularBounce = (flags >> 8) & 1u;
            const uint32_t vertex = flags >> 16;
            V3 beta = v3(paths.beta[0][slot], paths.beta[1][slot], paths.beta[2][slot]);
            V3 Ladd = v3(0.f);
            const uint32_t key = paths.key[slot];
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
                                flags = (uint32_t)(bounces + 1) | 0x100u | ((vertex + 1u) << 16);
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
                                flags = (uint32_t)(bounces + 1) | ((sampled & kBxSpecular) ? 0x100u : 0u) | ((vertex + 1u) << 16);
                            }
                        }
                    }
                }
            }
            if (!isBlack(Ladd)) { paths.L[0][slot] += Ladd.x; paths.L[1][slot] += Ladd.y; paths.L[2][slot] += Ladd.z; }
            if (pushNext) {
                paths.beta[0][slot] = beta.x; paths.beta[1][slot] = beta.y; paths.beta[2][slot] = beta.z;
                paths.flags[slot] = flags;
            }
        }
        __syncwarp();
        const uint32_t in = queuePush(nextCount, pushNext);
        if (pushNext) writeRay(nextQ, in, nO, nD, slot, kRayInf);
        __syncwarp();
        const uint32_t is = queuePush(q.count + 2, pushShadow);
        if (pushShadow) { writeRay(q.shadow, is, sO, sD, slot, sT); q.shadowC[is] = make_float4(sC.x, sC.y, sC.z, 0.f); }
        __syncwarp();
        const uint32_t im = queuePush(q.count + 3, pushMis);
        if (pushMis) { writeRay(q.mis, im, mO, mD, slot, kRayInf); q.misC[im] = make_float4(mC.x, mC.y, mC.z, __uint_as_float((uint32_t)mExpect)); }
        __syncwarp();
      }
      if (SORT) __syncthreads();      // s_order / s_cnt are rewritten by the next tile
    }
}

// after a bounce: fold the queue sizes into the statistics and clear the consumed counters
__global__ void bounceEndKernel(Queues q, int cur, unsigned long long* stats) {
    stats[1] += q.count[cur]; stats[2] += q.count[2]; stats[3] += q.count[3];
    q.count[cur] = 0u; q.count[2] = 0u; q.count[3] = 0u;
}

// ---- K5: film --------------------------------------------------------------------------------------
__device__ __forceinline__ float filterWeight(const RenderParamsPOD& rp, float dx, float dy) {
    switch (rp.filter) {
    case SPB_FILTER_TENT: return fmaxf(0.f, rp.frx - fabsf(dx)) * fmaxf(0.f, rp.fry - fabsf(dy));                 // filters/tent.cc:27-30
    case SPB_FILTER_GAUSSIAN: return fmaxf(0.f, expf(-rp.fbeta * dx * dx) - rp.fexpx) * fmaxf(0.f, expf(-rp.fbeta * dy * dy) - rp.fexpy);  // gaussian.cc:34-40
    default: return 1.f;                                                                                             // filters/box.cc:22
    }
}
__global__ void __launch_bounds__(256) filmKernel(RenderParamsPOD rp, PathSoA paths, float4* film, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t pixel = paths.pixel[i], key = paths.key[i];
    const int x = pixel % rp.width, y = pixel / rp.width;
    const float w = filterWeight(rp, sample1D(key, kDimFilm) - 0.5f, sample1D(key, kDimFilm + 1) - 0.5f);
    float Lr = paths.L[0][i], Lg = paths.L[1][i], Lb = paths.L[2][i];
    float4* dst = film + (size_t)y * rp.width + (rp.width - 1 - x);      // core/integrator.cc:88: pixel (width - x - 1, y)
    atomicAdd(dst, make_float4(w * Lr, w * Lg, w * Lb, w));
}
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
    freeScene(R); freeEnv(R);
    if (R->d_film) cudaFree(R->d_film);
    if (R->d_pool) cudaFree(R->d_pool);
    if (R->d_stats) cudaFree(R->d_stats);
    if (R->d_cursor) cudaFree(R->d_cursor);
    if (R->comm && R->nccl_lib) {
        typedef int (*destroy_t)(void*);
        destroy_t f = (destroy_t)dlsym(R->nccl_lib, "ncclCommDestroy");
        if (f) f(R->comm);
    }
    delete R;
    ctx->render = nullptr;
}

void renderSceneChanged(spb_ctx* ctx) { if (ctx->render) ctx->render->scene_dirty = true; }

namespace {
struct D3 { double x, y, z; };
inline D3 sub(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 crossd(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double normd(D3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline D3 normalized(D3 a) { const double n = normd(a); return n > 0 ? D3{a.x / n, a.y / n, a.z / n} : a; }
inline D3 scaled(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline D3 addd(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline void coordSys(D3 w, D3* u, D3* v) {          // core/vect_math.h:115-123
    if (std::abs(w.x) > std::abs(w.y)) *u = normalized({-w.z, 0, w.x}); else *u = normalized({0, w.z, -w.y});
    *v = normalized(crossd(w, *u));
}
inline void put3(float* d, D3 a) { d[0] = (float)a.x; d[1] = (float)a.y; d[2] = (float)a.z; }
}  // namespace

// Builds the per-triangle shading records the way the reference's Triangle constructors and
// Triangle::intersect do (core/triangle.cc:17-67, 119-137), in double, then narrows to float32.
static int uploadScene(spb_ctx* ctx, RenderState* R) {
    freeScene(R);
    const int64_t n = ctx->n_tris;
    R->n_tris = n;
    std::vector<ShadeTri> tris((size_t)n);
    const bool hasN = !ctx->normals.empty(), hasUV = !ctx->uvs.empty();
    for (int64_t i = 0; i < n; i++) {
        const double* v = ctx->verts.data() + i * 9;
        const D3 p0{v[0], v[1], v[2]}, p1{v[3], v[4], v[5]}, p2{v[6], v[7], v[8]};
        const D3 e1 = sub(p1, p0), e2 = sub(p2, p0);
        D3 fn = crossd(e1, e2);
        bool triHasN = false;
        if (hasN) {
            const float* nn = ctx->normals.data() + i * 9;
            triHasN = !(nn[0] == 0.f && nn[1] == 0.f && nn[2] == 0.f && nn[3] == 0.f && nn[4] == 0.f && nn[5] == 0.f &&
                        nn[6] == 0.f && nn[7] == 0.f && nn[8] == 0.f);
            if (triHasN && normd(fn) < 1e-12) fn = scaled(D3{(double)nn[0] + nn[3] + nn[6], (double)nn[1] + nn[4] + nn[7], (double)nn[2] + nn[5] + nn[8]}, 1.0 / 3.0);
        }
        fn = normalized(fn);
        D3 dpdu, dpdv;
        double detUV = 0.0, duv01[2] = {0, 0}, duv02[2] = {0, 0};
        if (hasUV) {
            const float* t = ctx->uvs.data() + i * 6;
            duv01[0] = (double)t[2] - t[0]; duv01[1] = (double)t[3] - t[1];
            duv02[0] = (double)t[4] - t[0]; duv02[1] = (double)t[5] - t[1];
            detUV = duv01[0] * duv02[1] - duv01[1] * duv02[0];
        }
        if (detUV == 0.0) coordSys(fn, &dpdu, &dpdv);
        else {
            const double inv = 1.0 / detUV;
            dpdu = addd(scaled(e1, duv02[1] * inv), scaled(e2, -duv01[1] * inv));
            dpdv = addd(scaled(e1, -duv02[0] * inv), scaled(e2, duv01[0] * inv));
        }
        const D3 ng = normalized(crossd(dpdu, dpdv));
        ShadeTri& s = tris[(size_t)i];
        std::memset(&s, 0, sizeof(s));
        put3(s.p0, p0);
        s.e1x = (float)e1.x; s.e1y = (float)e1.y; s.e1z = (float)e1.z;
        s.e2[0] = (float)e2.x; s.e2[1] = (float)e2.y; s.e2z = (float)e2.z;
        put3(s.ng, ng); put3(s.fn, fn);
        s.area = (float)(0.5 * normd(crossd(e1, e2)));
        put3(s.ss, normalized(dpdu)); put3(s.ts, normalized(dpdv));
        s.material = ctx->material_id[(size_t)i];
        s.light = ctx->light_id[(size_t)i];
        s.has_normals = triHasN ? 1 : 0;
    }
    if (n > 0) {
        SPB_CUDA(ctx, cudaMalloc(&R->d_tris, (size_t)n * sizeof(ShadeTri)));
        SPB_CUDA(ctx, cudaMemcpy(R->d_tris, tris.data(), (size_t)n * sizeof(ShadeTri), cudaMemcpyHostToDevice));
        if (hasN) {
            SPB_CUDA(ctx, cudaMalloc(&R->d_vnormals, (size_t)n * 9 * sizeof(float)));
            SPB_CUDA(ctx, cudaMemcpy(R->d_vnormals, ctx->normals.data(), (size_t)n * 9 * sizeof(float), cudaMemcpyHostToDevice));
        }
    }
    const size_t nm = std::max<size_t>(R->mats.size(), 1), nl = std::max<size_t>(R->lights.size(), 1);
    SPB_CUDA(ctx, cudaMalloc(&R->d_mats, nm * sizeof(spb_material)));
    SPB_CUDA(ctx, cudaMalloc(&R->d_lights, nl * sizeof(spb_light)));
    if (!R->mats.empty()) SPB_CUDA(ctx, cudaMemcpy(R->d_mats, R->mats.data(), R->mats.size() * sizeof(spb_material), cudaMemcpyHostToDevice));
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
        if (!ctx->uvs.empty()) uv = ctx->uvs;
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
    R->sort_materials = false;
    for (const spb_material& m : R->mats) if (m.type != R->mats[0].type) R->sort_materials = true;
    // every lobe type a material can turn into (makeBsdf: alpha 0 -> the specular lobe, a black coating -> Lambertian)
    R->type_mask = 0u;
    for (const spb_material& m : R->mats) {
        if (m.type < 0 || m.type > 6) continue;
        R->type_mask |= 1u << m.type;
        if (m.type == SPB_MAT_ROUGHCONDUCTOR) R->type_mask |= 1u << SPB_MAT_CONDUCTOR;
        if (m.type == SPB_MAT_ROUGHDIELECTRIC) R->type_mask |= 1u << SPB_MAT_DIELECTRIC;
        if (m.type == SPB_MAT_ROUGHPLASTIC) R->type_mask |= 1u << SPB_MAT_DIFFUSE;
    }
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

static int allocWave(spb_ctx* ctx, RenderState* R, int64_t slots) {
    if (R->d_pool && R->slots >= slots) return SPB_OK;
    if (R->d_pool) cudaFree(R->d_pool);
    R->d_pool = nullptr; R->slots = 0;
    // 9 u32/float per slot + 4 queues x 32 B + hits 16 B + 2 contributions x 16 B
    const size_t per = 9 * 4 + 4 * 32 + 16 + 2 * 16;
    const size_t bytes = (size_t)slots * per + 4096;
    cudaError_t e = cudaMalloc(&R->d_pool, bytes);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(ctx, SPB_ERR_OOM, "out of device memory for the wavefront queues"); }
    SPB_CUDA(ctx, e);
    char* p = (char*)R->d_pool;
    auto take = [&](size_t b) { void* r = p; p += (b + 255) & ~(size_t)255; return r; };
    R->q.ray[0] = (spb_ray_f32*)take((size_t)slots * 32); R->q.ray[1] = (spb_ray_f32*)take((size_t)slots * 32);
    R->q.shadow = (spb_ray_f32*)take((size_t)slots * 32); R->q.mis = (spb_ray_f32*)take((size_t)slots * 32);
    R->q.hit = (spb_hit*)take((size_t)slots * 16);
    R->q.shadowC = (float4*)take((size_t)slots * 16); R->q.misC = (float4*)take((size_t)slots * 16);
    for (int k = 0; k < 3; k++) { R->paths.beta[k] = (float*)take((size_t)slots * 4); R->paths.L[k] = (float*)take((size_t)slots * 4); }
    R->paths.pixel = (uint32_t*)take((size_t)slots * 4); R->paths.key = (uint32_t*)take((size_t)slots * 4);
    R->paths.flags = (uint32_t*)take((size_t)slots * 4);
    R->q.count = (uint32_t*)take(64);
    if ((size_t)(p - (char*)R->d_pool) > bytes + 0) { /* the 256 B rounding of 14 arrays fits in the 4 KiB slack */ }
    R->slots = slots;
    return SPB_OK;
}

}  // namespace spb

using namespace spb;

extern "C" {

int spb_scene_set_materials(spb_ctx* ctx, const spb_material* mats, int32_t n) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    if (n < 0 || (n > 0 && !mats)) return fail(ctx, SPB_ERR_INVALID, "spb_scene_set_materials: bad arguments");
    RenderState* R = rs(ctx);
    R->mats.assign(mats, mats + n);
    R->mat_tex.clear();
    R->scene_dirty = true;
    return SPB_OK;
}

int spb_scene_set_textures(spb_ctx* ctx, const spb_texture* texs, int32_t n, const float* texels_rgb, int64_t n_texels) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    if (n < 0 || (n > 0 && !texs) || n_texels < 0 || (n_texels > 0 && !texels_rgb)) return fail(ctx, SPB_ERR_INVALID, "spb_scene_set_textures: bad arguments");
    for (int32_t i = 0; i < n; i++)
        if (texs[i].type != SPB_TEX_BITMAP && texs[i].type != SPB_TEX_CHECKERBOARD) return fail(ctx, SPB_ERR_INVALID, "spb_scene_set_textures: unknown texture type");
    RenderState* R = rs(ctx);
    R->texs.assign(texs, texs + n);
    R->texels.assign(texels_rgb, texels_rgb + n_texels * 3);
    R->scene_dirty = true;
    return SPB_OK;
}

int spb_scene_set_material_textures(spb_ctx* ctx, const int32_t* tex_ids, int32_t n_mats) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    RenderState* R = rs(ctx);
    if (!tex_ids) { R->mat_tex.clear(); R->scene_dirty = true; return SPB_OK; }
    if (n_mats != (int32_t)R->mats.size()) return fail(ctx, SPB_ERR_INVALID, "spb_scene_set_material_textures: call spb_scene_set_materials first (counts differ)");
    R->mat_tex.assign(tex_ids, tex_ids + (size_t)n_mats * 2);
    R->scene_dirty = true;
    return SPB_OK;
}

int spb_scene_set_lights(spb_ctx* ctx, const spb_light* lights, int32_t n) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    if (n < 0 || (n > 0 && !lights)) return fail(ctx, SPB_ERR_INVALID, "spb_scene_set_lights: bad arguments");
    RenderState* R = rs(ctx);
    R->lights.assign(lights, lights + n);
    R->scene_dirty = true;
    return SPB_OK;
}

int spb_scene_set_envmap(spb_ctx* ctx, const float* rgb, int32_t w, int32_t h, const double l2w[16], double scale,
                         const double center[3], double radius) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    RenderState* R = rs(ctx);
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
    if (desc->integrator != SPB_INTEGRATOR_PATH && desc->integrator != SPB_INTEGRATOR_DIRECT) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: unknown integrator");
    if (!ctx->bvh_ready) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: no acceleration structure (call spb_bvh_build first)");
    cudaSetDevice(ctx->device);
    RenderState* R = rs(ctx);
    for (int32_t m : ctx->material_id)
        if (m >= (int32_t)R->mats.size()) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: a triangle refers to a material that was not set");
    for (size_t i = 0; i < R->lights.size(); i++) {
        const spb_light& l = R->lights[i];
        if (l.type == SPB_LIGHT_AREA && (l.prim < 0 || l.prim >= ctx->n_tris)) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: area light refers to a missing triangle");
        if (l.type == SPB_LIGHT_ENVMAP && !R->env_present) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: envmap light without spb_scene_set_envmap");
    }
    for (int32_t l : ctx->light_id)
        if (l >= (int32_t)R->lights.size()) return fail(ctx, SPB_ERR_INVALID, "spb_render_begin: a triangle refers to a light that was not set");
    int rc;
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
    if (R->film_pixels != npix) {
        if (R->d_film) cudaFree(R->d_film);
        R->d_film = nullptr;
        SPB_CUDA(ctx, cudaMalloc(&R->d_film, (size_t)npix * sizeof(float4)));
        R->film_pixels = npix;
    }
    SPB_CUDA(ctx, cudaMemsetAsync(R->d_film, 0, (size_t)npix * sizeof(float4), ctx->stream));
    if (!R->d_stats) { SPB_CUDA(ctx, cudaMalloc(&R->d_stats, 8 * sizeof(unsigned long long))); }
    if (!R->d_cursor) { SPB_CUDA(ctx, cudaMalloc(&R->d_cursor, 16 * sizeof(unsigned long long))); }
    SPB_CUDA(ctx, cudaMemsetAsync(R->d_stats, 0, 8 * sizeof(unsigned long long), ctx->stream));
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    R->launches = 0; R->render_ms = 0.0; R->paths_total = 0;
    // Set-up that would otherwise land in the first spb_render_samples: the wavefront pool (capacity for 64 spp
    // or one full wave, whichever is smaller) and the lazy loading of the loop's kernels -- with several
    // contexts driven from one process the driver serialises those loads across the GPUs.
    if ((rc = allocWave(ctx, R, std::min<int64_t>(ctx->opt_wave_slots, npix * 64)))) return rc;
    {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, generateKernel);
        if (!(R->type_mask & ~kTypesDiffuse)) cudaFuncGetAttributes(&fa, shadeKernel<false, kShadeMinbDiffuse, kTypesDiffuse>);
        else if (!(R->type_mask & ~kTypesGlossy)) { cudaFuncGetAttributes(&fa, shadeKernel<true, 4, kTypesGlossy>); cudaFuncGetAttributes(&fa, shadeKernel<false, 4, kTypesGlossy>); }
        else { cudaFuncGetAttributes(&fa, shadeKernel<true>); cudaFuncGetAttributes(&fa, shadeKernel<false>); }
        cudaFuncGetAttributes(&fa, bounceEndKernel);
        cudaFuncGetAttributes(&fa, filmKernel);
        if (ctx->sp.tri_format == 0) {
            cudaFuncGetAttributes(&fa, traceCoopKernel<0, false, false, spb_ray_f32, HitOut, 8>);
            cudaFuncGetAttributes(&fa, traceCoopKernel<0, true, false, spb_ray_f32, SinkShadow, 8>);
            cudaFuncGetAttributes(&fa, traceCoopKernel<0, false, false, spb_ray_f32, SinkMis, 8>);
        } else {
            cudaFuncGetAttributes(&fa, traceCoopKernel<1, false, false, spb_ray_f32, HitOut, 8>);
            cudaFuncGetAttributes(&fa, traceCoopKernel<1, true, false, spb_ray_f32, SinkShadow, 8>);
            cudaFuncGetAttributes(&fa, traceCoopKernel<1, false, false, spb_ray_f32, SinkMis, 8>);
        }
        cudaGetLastError();
    }
    R->begun = true;
    return SPB_OK;
}

int spb_render_samples(spb_ctx* ctx, int32_t first, int32_t count, int32_t stride) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    RenderState* R = ctx->render;
    if (!R || !R->begun) return fail(ctx, SPB_ERR_INVALID, "spb_render_samples: call spb_render_begin first");
    if (count < 0 || stride <= 0 || first < 0) return fail(ctx, SPB_ERR_INVALID, "spb_render_samples: bad sample range");
    if (count == 0) return SPB_OK;
    cudaSetDevice(ctx->device);
    const int64_t npix = (int64_t)R->rp.width * R->rp.height;
    const int64_t total = npix * count;
    const int64_t maxSlots = ctx->opt_wave_slots;
    const int64_t slots = std::min(total, maxSlots);
    int rc = allocWave(ctx, R, slots);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    SPB_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    const int shadeGrid = ctx->sm_count * 8;
    for (int64_t item0 = 0; item0 < total; item0 += slots) {
        const int64_t n = std::min(slots, total - item0);
        generateKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(R->rp, R->cam, R->paths, R->q, item0, n, first, stride);
        R->launches++;
        int cur = 0;
        for (int bounce = 0; bounce <= R->rp.max_depth; bounce++) {
            // extend
            if ((rc = launchTrace<false>(ctx, R->q.ray[cur], n, R->q.count + cur, HitOut{R->q.hit}, R->d_cursor, st))) return rc;
            if (!(R->type_mask & ~kTypesDiffuse) && ctx->opt_shade_generic == 0) {
                const int mb = ctx->opt_shade_minb > 0 ? ctx->opt_shade_minb : kShadeMinbDiffuse;
                if (mb == 4) shadeKernel<false, 4, kTypesDiffuse><<<shadeGrid, 128, 0, st>>>(R->rp, R->ds, R->paths, R->q, cur);
                else if (mb == 6) shadeKernel<false, 6, kTypesDiffuse><<<shadeGrid, 128, 0, st>>>(R->rp, R->ds, R->paths, R->q, cur);
                else shadeKernel<false, 5, kTypesDiffuse><<<shadeGrid, 128, 0, st>>>(R->rp, R->ds, R->paths, R->q, cur);
            }
            else if (!(R->type_mask & ~kTypesGlossy) && ctx->opt_shade_generic == 0) {
                if (R->sort_materials) shadeKernel<true, 4, kTypesGlossy><<<shadeGrid, 128, 0, st>>>(R->rp, R->ds, R->paths, R->q, cur);
                else shadeKernel<false, 4, kTypesGlossy><<<shadeGrid, 128, 0, st>>>(R->rp, R->ds, R->paths, R->q, cur);
            } else if (R->sort_materials) shadeKernel<true><<<shadeGrid, 128, 0, st>>>(R->rp, R->ds, R->paths, R->q, cur);
            else shadeKernel<false><<<shadeGrid, 128, 0, st>>>(R->rp, R->ds, R->paths, R->q, cur);
            // connect + MIS (skipped by their own zero counts when empty)
            if ((rc = launchTrace<true>(ctx, R->q.shadow, n, R->q.count + 2, SinkShadow{R->q.shadow, R->q.shadowC, R->paths}, R->d_cursor + 4, st))) return rc;
            if ((rc = launchTrace<false>(ctx, R->q.mis, n, R->q.count + 3, SinkMis{R->q.mis, R->q.misC, R->paths}, R->d_cursor + 8, st))) return rc;
            bounceEndKernel<<<1, 1, 0, st>>>(R->q, cur, R->d_stats);
            R->launches += 5;
            cur ^= 1;
        }
        filmKernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(R->rp, R->paths, R->d_film, n);
        R->launches++;
        SPB_CUDA(ctx, cudaGetLastError());
    }
    SPB_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    SPB_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    SPB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    R->render_ms += ms;
    R->paths_total += total;
    return SPB_OK;
}

int spb_film_read(spb_ctx* ctx, float* rgbw) {
    if (!ctx || !rgbw) return fail(ctx, SPB_ERR_INVALID, "spb_film_read: NULL argument");
    RenderState* R = ctx->render;
    if (!R || !R->begun) return fail(ctx, SPB_ERR_INVALID, "spb_film_read: no film (call spb_render_begin)");
    cudaSetDevice(ctx->device);
    SPB_CUDA(ctx, cudaMemcpyAsync(rgbw, R->d_film, (size_t)R->film_pixels * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPB_OK;
}

int spb_film_resolve(spb_ctx* ctx, float* rgb) {
    if (!ctx || !rgb) return fail(ctx, SPB_ERR_INVALID, "spb_film_resolve: NULL argument");
    RenderState* R = ctx->render;
    if (!R || !R->begun) return fail(ctx, SPB_ERR_INVALID, "spb_film_resolve: no film (call spb_render_begin)");
    cudaSetDevice(ctx->device);
    float* d = nullptr;
    SPB_CUDA(ctx, cudaMalloc(&d, (size_t)R->film_pixels * 3 * sizeof(float)));
    resolveKernel<<<(unsigned)((R->film_pixels + 255) / 256), 256, 0, ctx->stream>>>(R->d_film, d, R->film_pixels);
    R->launches++;
    cudaError_t e = cudaMemcpyAsync(rgb, d, (size_t)R->film_pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    SPB_CUDA(ctx, e);
    return SPB_OK;
}

static int filmEncode(spb_ctx* ctx, void* host, size_t bytesPerPixel, double invGamma, const char* what) {
    if (!ctx || !host) return fail(ctx, SPB_ERR_INVALID, std::string(what) + ": NULL argument");
    RenderState* R = ctx->render;
    if (!R || !R->begun) return fail(ctx, SPB_ERR_INVALID, std::string(what) + ": no film (call spb_render_begin)");
    cudaSetDevice(ctx->device);
    unsigned char* d = nullptr;
    SPB_CUDA(ctx, cudaMalloc(&d, (size_t)R->film_pixels * bytesPerPixel));
    const unsigned grid = (unsigned)((R->film_pixels + 255) / 256);
    if (bytesPerPixel == 4) encodeRgbeKernel<<<grid, 256, 0, ctx->stream>>>(R->d_film, (uchar4*)d, R->film_pixels);
    else encodeLdrKernel<<<grid, 256, 0, ctx->stream>>>(R->d_film, d, invGamma, R->film_pixels);
    R->launches++;
    cudaError_t e = cudaMemcpyAsync(host, d, (size_t)R->film_pixels * bytesPerPixel, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    SPB_CUDA(ctx, e);
    return SPB_OK;
}
int spb_film_resolve_rgbe(spb_ctx* ctx, uint8_t* rgbe) { return filmEncode(ctx, rgbe, 4, 0.0, "spb_film_resolve_rgbe"); }
int spb_film_resolve_ldr(spb_ctx* ctx, double gamma, uint8_t* rgb8) {
    if (!(gamma >= 1.0e-12)) return fail(ctx, SPB_ERR_INVALID, "spb_film_resolve_ldr: too small gamma (core/tmo.cc:54)");
    return filmEncode(ctx, rgb8, 3, 1.0 / gamma, "spb_film_resolve_ldr");
}

int spb_film_add(spb_ctx* ctx, const float* rgbw) {
    if (!ctx || !rgbw) return fail(ctx, SPB_ERR_INVALID, "spb_film_add: NULL argument");
    RenderState* R = ctx->render;
    if (!R || !R->begun) return fail(ctx, SPB_ERR_INVALID, "spb_film_add: no film (call spb_render_begin)");
    cudaSetDevice(ctx->device);
    float4* d = nullptr;
    SPB_CUDA(ctx, cudaMalloc(&d, (size_t)R->film_pixels * sizeof(float4)));
    cudaError_t e = cudaMemcpyAsync(d, rgbw, (size_t)R->film_pixels * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        filmAddKernel<<<(unsigned)((R->film_pixels + 255) / 256), 256, 0, ctx->stream>>>(R->d_film, d, R->film_pixels);
        R->launches++;
        e = cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(d);
    SPB_CUDA(ctx, e);
    return SPB_OK;
}

int spb_render_get_stats(spb_ctx* ctx, spb_render_stats* out) {
    if (!ctx || !out) return fail(ctx, SPB_ERR_INVALID, "spb_render_get_stats: NULL argument");
    RenderState* R = ctx->render;
    std::memset(out, 0, sizeof(*out));
    if (!R || !R->begun) return SPB_OK;
    cudaSetDevice(ctx->device);
    unsigned long long h[8];
    SPB_CUDA(ctx, cudaMemcpy(h, R->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
    out->paths = R->paths_total;
    out->rays_closest = (int64_t)h[1]; out->rays_shadow = (int64_t)h[2]; out->rays_mis = (int64_t)h[3];
    out->kernel_launches = R->launches; out->render_ms = R->render_ms;
    return SPB_OK;
}

// ---- NCCL (loaded lazily: libnccl.so.2 is only needed by multi-GPU jobs) -----------------------------
typedef struct { char internal[128]; } spbNcclUniqueId;
typedef int (*ncclGetUniqueId_t)(spbNcclUniqueId*);
typedef int (*ncclCommInitRank_t)(void**, int, spbNcclUniqueId, int);
typedef int (*ncclAllReduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*ncclCommDestroy_t)(void*);
typedef const char* (*ncclGetErrorString_t)(int);

static void* ncclLib() {
    static void* lib = nullptr;
    if (lib) return lib;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nme : names) { lib = dlopen(nme, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
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
    ncclAllReduce_t ar = (ncclAllReduce_t)dlsym(lib, "ncclAllReduce");
    if (ar) {
        float* d_warm = nullptr;
        SPB_CUDA(ctx, cudaMalloc(&d_warm, 1024 * sizeof(float)));
        SPB_CUDA(ctx, cudaMemsetAsync(d_warm, 0, 1024 * sizeof(float), ctx->stream));
        const int wrc = ar(d_warm, d_warm, 1024, 7, 0, R->comm, ctx->stream);
        SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(d_warm);
        if (wrc != 0) return fail(ctx, SPB_ERR_CUDA, "ncclAllReduce (communicator warm-up) failed");
    }
    return SPB_OK;
}

int spb_film_allreduce(spb_ctx* ctx) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    RenderState* R = ctx->render;
    if (!R || !R->begun) return fail(ctx, SPB_ERR_INVALID, "spb_film_allreduce: no film");
    if (!R->comm) return fail(ctx, SPB_ERR_INVALID, "spb_film_allreduce: call spb_comm_init first");
    cudaSetDevice(ctx->device);
    ncclAllReduce_t f = (ncclAllReduce_t)dlsym(R->nccl_lib, "ncclAllReduce");
    if (!f) return fail(ctx, SPB_ERR_UNSUPPORTED, "ncclAllReduce not found");
    // K7: one sum over the RGBW film per frame (ncclFloat32 = 7, ncclSum = 0)
    const int rc = f(R->d_film, R->d_film, (size_t)R->film_pixels * 4, 7, 0, R->comm, ctx->stream);
    if (rc != 0) {
        ncclGetErrorString_t es = (ncclGetErrorString_t)dlsym(R->nccl_lib, "ncclGetErrorString");
        return fail(ctx, SPB_ERR_CUDA, std::string("ncclAllReduce: ") + (es ? es(rc) : "error"));
    }
    SPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SPB_OK;
}

int spb_comm_destroy(spb_ctx* ctx) {
    if (!ctx) return fail(nullptr, SPB_ERR_INVALID, "ctx is NULL");
    RenderState* R = ctx->render;
    if (!R || !R->comm) return SPB_OK;
    ncclCommDestroy_t f = (ncclCommDestroy_t)dlsym(R->nccl_lib, "ncclCommDestroy");
    if (f) f(R->comm);
    R->comm = nullptr;
    return SPB_OK;
}

}  // extern "C"
