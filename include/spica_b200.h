/* spica_b200 -- thin C ABI between the C++17 host (spica's plugin surface) and the sm_100a CUDA
 * library (libspica_b200.so).  POD only, extern "C", no exceptions, no STL, no torch types.
 *
 * This header is the drop-in boundary for ONE hot path of tatsy/spica: BVH ray traversal plus the
 * unidirectional path-tracing integrator loop.  Every entry point names the
 * reference interface it replaces (paths relative to /root/reference/sources).  INTEGRATION.md
 * shows the reference-side bindings (the `bvh` accelerator plugin and the `path` integrator
 * plugin) a maintainer would add on top of these calls.
 *
 * Conventions
 *   - every call returns SPB_OK (0) or a negative spb_status; spb_last_error(ctx) gives the text.
 *     The reference aborts on error (core/common.h:71-83,109-115); the host shim maps non-zero to
 *     the same FatalError behaviour.
 *   - one context per GPU; calls on a context are stream-ordered and are synchronous on return
 *     unless the name ends in _async.  The caller owns every host buffer.
 *   - names ending in _dev take DEVICE pointers (already resident in HBM on the context's GPU).
 *   - there is NO CPU fallback: without a CUDA device spb_ctx_create fails with SPB_ERR_NO_DEVICE.
 */
#ifndef SPICA_B200_H_
#define SPICA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPB_VERSION 1

typedef enum spb_status {
    SPB_OK = 0,
    SPB_ERR_INVALID = -1,     /* bad argument / call order                                   */
    SPB_ERR_NO_DEVICE = -2,   /* no CUDA device: the product has no CPU path                 */
    SPB_ERR_CUDA = -3,        /* a CUDA runtime call or kernel failed                        */
    SPB_ERR_OOM = -4,
    SPB_ERR_UNSUPPORTED = -5
} spb_status;

typedef struct spb_ctx spb_ctx; /* opaque; owns all device memory, freed by spb_ctx_destroy */

/* ---- ray / hit records ------------------------------------------------------------------ */

/* Ray as the reference's Ray constructor receives it (core/ray.h:26-27): origin, UN-normalised
 * direction, maxDist.  The library applies the constructor's arithmetic itself, in double
 * (core/ray.cc:11-19,43-47: dir *= 1.0/sqrt(d.d)).  tmin is reserved (the reference has none; hits
 * need t > 1e-12, core/triangle.cc:117).  A zero direction yields a miss (the reference aborts). */
typedef struct spb_ray_f32 { float ox, oy, oz, dx, dy, dz, tmin, tmax; } spb_ray_f32;    /* 32 B */
typedef struct spb_ray_f64 { double ox, oy, oz, dx, dy, dz, tmin, tmax; } spb_ray_f64;   /* 64 B */

/* Closest hit.  prim = index into the triangle array given to spb_scene_set_triangles (the
 * reference's primitive index, accelerators/bvh.cc:343), -1 on a miss.  (u, v) are the
 * Moeller-Trumbore barycentrics of core/triangle.cc:106-113. */
typedef struct spb_hit { float t; int32_t prim; float u, v; } spb_hit;                   /* 16 B */
typedef struct spb_hit_f64 { double t, u, v; int32_t prim; int32_t pad_; } spb_hit_f64;  /* 32 B */

/* ---- context ------------------------------------------------------------------------------ */

/* Replaces nothing in the reference (it has no device): one per GPU, created by the `bvh`
 * accelerator plugin's constructor (accelerators/bvh.cc:112-131). */
int spb_ctx_create(int device, spb_ctx** out);
void spb_ctx_destroy(spb_ctx* ctx);
const char* spb_last_error(const spb_ctx* ctx); /* ctx may be NULL: last global error        */
int spb_version(void);

/* ---- geometry ------------------------------------------------------------------------------ */

/* World-space triangle soup, 9 doubles per triangle (p0, p1, p2), exactly the points_ the
 * reference's Triangle stores after applying objectToWorld (core/triangle.cc:17-22).
 * normals (9 floats per triangle, per-vertex, may be NULL = face normal; core/triangle.cc:25-30),
 * material_id / light_id (per triangle, may be NULL = 0 / -1) are only used by spb_render.
 * Replaces: the std::vector<std::shared_ptr<Primitive>> handed to the accelerator factory
 * (core/cobject.h:66-75, core/accelerator.h:26-27). */
int spb_scene_set_triangles(spb_ctx* ctx, const double* verts, const float* normals,
                            const int32_t* material_id, const int32_t* light_id, int64_t n_tris);

/* ---- acceleration structure ---------------------------------------------------------------- */

typedef struct spb_build_opts {
    int32_t builder;        /* 0 = host binned-SAH (default), 1 = GPU LBVH                      */
    int32_t max_leaf_tris;  /* 1..3, default 3                                                  */
    int32_t sah_bins;       /* default 32                                                       */
    int32_t reserved_;
} spb_build_opts;

/* Replaces BVHAccel::construct / constructRec (accelerators/bvh.cc:139-237). opts may be NULL.
 * Builds a binary SAH tree, collapses it to the 8-wide compressed layout and uploads it. */
int spb_bvh_build(spb_ctx* ctx, const spb_build_opts* opts);

/* One node of a reference-built binary BVH (accelerators/bvh.h:28-52) with pointers replaced by
 * indices; this is what oracle/raycast_ref --dump-bvh writes. */
typedef struct spb_import_node {
    double lo[3], hi[3];
    int32_t left, right; /* -1 when absent                                                    */
    int32_t prim;        /* >= 0 for a leaf                                                    */
    int32_t axis;
} spb_import_node;

/* "import the reference-built BVH for exact comparison" (BASELINE.json north_star): uses the given
 * topology instead of building one; the wide collapse happens on top of it.  With an imported
 * tree exact-t ties resolve as in the reference (the leaf that is leftmost in tree order wins,
 * accelerators/bvh.cc:351-356 + core/triangle.cc:117); with an own-built tree the lower primitive
 * index wins. */
int spb_bvh_import_binary(spb_ctx* ctx, const spb_import_node* nodes, int64_t n_nodes, int32_t root);

typedef struct spb_bvh_stats {
    int64_t n_tris, n_wide_nodes, n_binary_nodes;
    int64_t node_bytes, tri_bytes;
    double sah_cost, build_seconds;
    int32_t tri_format;     /* 0: float32-exact vertices (48 B/tri), 1: float64 (80 B/tri)       */
    int32_t max_depth;
    double world_lo[3], world_hi[3]; /* Accelerator::worldBound (accelerators/bvh.cc:135-137)   */
} spb_bvh_stats;
int spb_bvh_get_stats(const spb_ctx* ctx, spb_bvh_stats* out);

/* ---- ray casting ----------------------------------------------------------------------------- */

/* Replaces BVHAccel::intersect(Ray&, SurfaceInteraction*) (accelerators/bvh.cc:315-321,331-360)
 * over a batch.  Host buffers; the copies are part of the call. */
int spb_trace_closest(spb_ctx* ctx, const spb_ray_f32* rays, int64_t n, spb_hit* hits);
int spb_trace_closest_f64(spb_ctx* ctx, const spb_ray_f64* rays, int64_t n, spb_hit_f64* hits);
/* Replaces BVHAccel::intersect(Ray&) (accelerators/bvh.cc:323-329,362-387): 1 = occluded. */
int spb_trace_any(spb_ctx* ctx, const spb_ray_f32* rays, int64_t n, uint8_t* occluded);
int spb_trace_any_f64(spb_ctx* ctx, const spb_ray_f64* rays, int64_t n, uint8_t* occluded);

/* Same, on buffers already resident in HBM (device pointers), on the context's stream. */
int spb_trace_closest_dev(spb_ctx* ctx, const spb_ray_f32* d_rays, int64_t n, spb_hit* d_hits);
int spb_trace_any_dev(spb_ctx* ctx, const spb_ray_f32* d_rays, int64_t n, uint8_t* d_occluded);

/* Device timing of the most recent trace kernel on this context (CUDA events on the launching
 * stream), and traversal counters when enabled with spb_set_option("counters", 1). */
typedef struct spb_counters {
    double last_kernel_ms;
    int64_t kernel_launches;      /* kernels launched by this context since creation            */
    int64_t rays, node_visits, tri_tests; /* only with the "counters" option                    */
} spb_counters;
int spb_get_counters(spb_ctx* ctx, spb_counters* out);

/* Tunables: "counters" (0/1), "trace_block" (threads per CTA), "trace_ctas_per_sm",
 * "trace_variant" (kernel variant id). Unknown names return SPB_ERR_INVALID. */
int spb_set_option(spb_ctx* ctx, const char* name, int64_t value);

/* raw device memory helpers so that a host without the CUDA runtime (a plugin, ctypes) can keep
 * buffers resident between calls */
int spb_dev_alloc(spb_ctx* ctx, size_t bytes, void** d_ptr);
int spb_dev_free(spb_ctx* ctx, void* d_ptr);
int spb_dev_upload(spb_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int spb_dev_download(spb_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
int spb_dev_sync(spb_ctx* ctx);
/* the CUDA stream (cudaStream_t) the context launches on, for event timing by the caller */
void* spb_ctx_stream(spb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SPICA_B200_H_ */
