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

#define SPB_VERSION 3

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
 * normals (9 floats per triangle, per-vertex, world space, may be NULL = face normal;
 * core/triangle.cc:25-30,40-42), uvs (6 floats per triangle, may be NULL = all zero; core/triangle.cc:64),
 * material_id / light_id (per triangle, may be NULL = 0 / -1; index into spb_scene_set_materials /
 * spb_scene_set_lights) are only used by spb_render_*.
 * Replaces: the std::vector<std::shared_ptr<Primitive>> handed to the accelerator factory
 * (core/cobject.h:66-75, core/accelerator.h:26-27). */
int spb_scene_set_triangles(spb_ctx* ctx, const double* verts, const float* normals, const float* uvs,
                            const int32_t* material_id, const int32_t* light_id, int64_t n_tris);

/* Replaces only the per-triangle material / light indices (either may be NULL = keep), leaving the
 * geometry and the acceleration structure untouched: the accelerator is built from the
 * primitives before the light list is final (spica/sceneparser.cc:91-95). */
int spb_scene_set_triangle_attributes(spb_ctx* ctx, const int32_t* material_id, const int32_t* light_id, int64_t n_tris);

/* ---- materials, lights (device PODs of the reference's bsdf / emitter plugins) ------------------ */

enum {
    SPB_MAT_NONE = -1,            /* no BSDF: the path passes straight through (integrators/path/path.cc:71-75) */
    SPB_MAT_DIFFUSE = 0,          /* bsdfs/diffuse.cc:23-32      -> LambertianReflection (core/bxdf.cc:34-58)   */
    SPB_MAT_DIELECTRIC = 1,       /* bsdfs/dielectric.cc:30-42   -> FresnelSpecular (core/bxdf.cc:156-248)      */
    SPB_MAT_ROUGHCONDUCTOR = 2,   /* bsdfs/roughconductor.cc:38-75 -> MicrofacetReflection (core/bxdf.cc:255-297) */
    SPB_MAT_ROUGHDIELECTRIC = 3,  /* bsdfs/roughdielectric.cc:41-83 -> MicrofacetTransmission (core/bxdf.cc:304-424) */
    SPB_MAT_CONDUCTOR = 4,        /* alpha == 0 roughconductor / bsdfs/conductor.cc -> SpecularReflection        */
    SPB_MAT_PLASTIC = 5,          /* bsdfs/plastic.cc:20-128       -> PlasticBRDF (smooth coating over a diffuse base) */
    SPB_MAT_ROUGHPLASTIC = 6      /* bsdfs/roughplastic.cc:20-165  -> RoughPlasticBRDF (microfacet coating)          */
};
enum { SPB_DISTR_BECKMANN = 0, SPB_DISTR_GGX = 1 };

typedef struct spb_material {
    int32_t type;           /* SPB_MAT_*                                                              */
    int32_t distribution;   /* SPB_DISTR_* (rough materials; default "beckmann")                       */
    float   kr[3];          /* diffuse: reflectance; dielectrics, plastics: specularReflectance; conductors: 1 */
    float   kt[3];          /* dielectrics: specularTransmittance; plastics: diffuseReflectance        */
    float   eta[3];         /* conductors: eta (rgb); dielectrics, plastics: intIOR in eta[0] (extIOR is 1.0) */
    float   k[3];           /* conductors: absorption k (rgb)                                          */
    float   alpha_u, alpha_v; /* microfacet roughness (no remap, bsdfs/roughconductor.cc:33); roughplastic: alpha_u */
} spb_material;             /* 64 B */
int spb_scene_set_materials(spb_ctx* ctx, const spb_material* mats, int32_t n);

/* Textures (textures/bitmap.cc:16-26, textures/checkerboard.cc:23-40): the reflectance-type parameters of a
 * material (kr, kt) may come from a texture evaluated at the hit point's interpolated uv
 * (core/triangle.cc:120) instead of the constant in spb_material.
 *   bitmap       : MipMap::lookup(st, 0) = bilinear on level 0 with Repeat wrap (core/mipmap.cc:67-93), st = (u, 1 - v)
 *                  (UVMapping2D's default invertHorizontal, core/texture.cc:24-31);
 *   checkerboard : color0 when int((u*uscale+uoffset)*2) + int((v*vscale+voffset)*2) is odd, else color1. */
enum { SPB_TEX_BITMAP = 0, SPB_TEX_CHECKERBOARD = 1 };
typedef struct spb_texture {
    int32_t type;               /* SPB_TEX_*                                                        */
    int32_t width, height;      /* bitmap                                                           */
    int32_t reserved_;
    int64_t texel_offset;       /* bitmap: index of its first texel in the texel array (rgb triples) */
    float   color0[3], color1[3];
    float   uoffset, voffset, uscale, vscale;
} spb_texture;                  /* 64 B */
/* texels_rgb: 3 floats per texel, row-major per bitmap; copied. */
int spb_scene_set_textures(spb_ctx* ctx, const spb_texture* texs, int32_t n, const float* texels_rgb, int64_t n_texels);
/* tex_ids: 2 ints per material {texture for kr, texture for kt}, -1 = use the constant.  NULL clears all bindings.
 * Call after spb_scene_set_materials (which resets the bindings). */
int spb_scene_set_material_textures(spb_ctx* ctx, const int32_t* tex_ids, int32_t n_mats);

enum { SPB_LIGHT_AREA = 0, SPB_LIGHT_ENVMAP = 1 };
/* One entry per light in Scene::lights() order (spica/sceneparser.cc:168-179: ONE AreaLight per
 * emitter triangle; lights/area.cc).  The list order only matters for the uniform light pick
 * (core/mis.cc:27).  An SPB_LIGHT_ENVMAP entry refers to the map given to spb_scene_set_envmap. */
typedef struct spb_light {
    int32_t type;           /* SPB_LIGHT_*                                                            */
    int32_t prim;           /* area: the emitter triangle (index into the triangle array)              */
    float   radiance[3];    /* area: Lemit                                                             */
    float   pad_;
} spb_light;                /* 24 B */
int spb_scene_set_lights(spb_ctx* ctx, const spb_light* lights, int32_t n);

/* lat-long environment map (lights/envmap.cc:17-58): rgb = w*h*3 floats, row-major, already
 * multiplied by nothing (scale is applied here); light_to_world = the XML toWorld matrix (row-major
 * 4x4; the reference transposes it, envmap.cc:19); world_radius bounds shadow rays (envmap.cc:75). */
int spb_scene_set_envmap(spb_ctx* ctx, const float* rgb, int32_t w, int32_t h, const double light_to_world[16],
                         double scale, const double world_center[3], double world_radius);

/* ---- the path-tracing integrator ------------------------------------------------------------------ */

enum { SPB_FILTER_BOX = 0, SPB_FILTER_TENT = 1, SPB_FILTER_GAUSSIAN = 2 };
enum {
    SPB_INTEGRATOR_PATH = 0,    /* integrators/path/path.cc:42-125 (the hot path)                                */
    SPB_INTEGRATOR_DIRECT = 1   /* integrators/directlighting/directlighting.cc:21-57: same kernels, no indirect light */
};

/* Everything SamplerIntegrator::render (core/integrator.cc:46-110) + PathIntegrator::Li
 * (integrators/path/path.cc:42-125) read from the camera / film / sampler / params objects,
 * resolved once into a POD. Matrices are row-major 4x4 doubles. */
typedef struct spb_render_desc {
    int32_t width, height;          /* film resolution (films/hdrfilm.cc:15-18)                        */
    int32_t max_depth;              /* "maxDepth", default 16 (spica/sceneparser.cc:63)                */
    int32_t filter;                 /* SPB_FILTER_* ; weight of the sample's own pixel (core/film.cc:65-74) */
    double  filter_radius[2];       /* filters: box.cc:22, tent.cc:27, gaussian.cc:34                    */
    double  filter_sigma;
    double  camera_to_world[16];    /* cameras/perspective.cc:53-74, core/camera.cc:20-38              */
    double  raster_to_camera[16];
    double  lens_radius, focal_distance;
    uint64_t seed;                  /* counter-based sampler key (replaces time(0) seeding, core/integrator.cc:51,71) */
    int32_t rr_start_bounce;        /* Russian roulette applies when bounces > this; reference: 3 (path.cc:117) */
    int32_t integrator;             /* SPB_INTEGRATOR_* (0 = path)                                       */
} spb_render_desc;

/* Allocates (or re-uses) and clears the device film and the wavefront queues, uploads whatever changed in the scene,
 * and prepares the render loop (kernels loaded, its iteration captured into a CUDA graph).  max_depth <= 4095.
 * Any later change to the scene (triangles, attributes, materials, lights, textures, environment, a new
 * acceleration structure) invalidates the render: spb_render_samples* then return SPB_ERR_INVALID until
 * spb_render_begin is called again. */
int spb_render_begin(spb_ctx* ctx, const spb_render_desc* desc);
/* Renders sample indices first, first+stride, ... (count of them) for every pixel and accumulates
 * them into the device film: one iteration of the spp loop of SamplerIntegrator::render per index.
 * A rank of an N-GPU job passes (first = rank, stride = N).  Synchronous: returns when the samples are in the film. */
int spb_render_samples(spb_ctx* ctx, int32_t first, int32_t count, int32_t stride);
/* The same, asynchronous: returns at once; the context's render worker drives the loop.  Calls queue up in order
 * (so do spb_film_reduce_async calls between them).  spb_render_wait blocks until everything queued has finished
 * and returns the first error any of it reported (spb_last_error has the text).  Every other call that reads or
 * changes the film, the scene or the statistics waits for the queue first. */
int spb_render_samples_async(spb_ctx* ctx, int32_t first, int32_t count, int32_t stride);
int spb_render_wait(spb_ctx* ctx);
/* Raw accumulators, height x width x 4 floats {sum w*r, sum w*g, sum w*b, sum w}, pixel (x, y) as
 * Film::addPixel indexed them (the horizontal flip of core/integrator.cc:88 already applied). */
int spb_film_read(spb_ctx* ctx, float* rgbw);
/* Film::save's normalisation (core/film.cc:23-29): rgb = sum / (wsum + 1e-12); height x width x 3. */
int spb_film_resolve(spb_ctx* ctx, float* rgb);
/* The film's output stage on the device: normalise as spb_film_resolve does, then encode
 * the pixel the way the reference's film plugins do before writing the file, so only the encoded bytes cross PCIe.
 *   rgbe : HDRFilm -> Image::saveHdr's HDRPixel (core/image.cc:60-88): m = frexp(max(r,g,b)), bytes = c * m*256/max,
 *          exponent + 128; height x width x 4 bytes.
 *   ldr  : LDRFilm -> GammaTmo (core/tmo.cc:53-71: clamp(pow(c, 1/gamma), 0, 1)) then Image::toByte
 *          (core/image.cc:484-487: uint8(255 * c)); height x width x 3 bytes. */
int spb_film_resolve_rgbe(spb_ctx* ctx, uint8_t* rgbe);
int spb_film_resolve_ldr(spb_ctx* ctx, double gamma, uint8_t* rgb8);
/* Adds raw accumulators (same layout as spb_film_read) into the device film: resume / merge. */
int spb_film_add(spb_ctx* ctx, const float* rgbw);

typedef struct spb_render_stats {
    int64_t paths, rays_closest, rays_shadow, rays_mis;   /* rays traced since spb_render_begin   */
    int64_t kernel_launches;
    double  render_ms;                                    /* device time of all spb_render_samples */
    double  reduce_ms;                                    /* device time of all film reductions (includes waiting for the slowest rank) */
    int64_t iterations;                                   /* iterations of the streaming loop (one extend + shade + connect + MIS each) */
} spb_render_stats;
int spb_render_get_stats(spb_ctx* ctx, spb_render_stats* out);

/* ---- multi-GPU: one context per GPU, films summed with one NCCL all-reduce over NVLink ----------- */
#define SPB_COMM_ID_BYTES 128
int spb_comm_get_unique_id(char id[SPB_COMM_ID_BYTES]);                 /* rank 0; ship to the others */
int spb_comm_init(spb_ctx* ctx, const char id[SPB_COMM_ID_BYTES], int32_t n_ranks, int32_t rank);
/* One sum (float32) over the RGBW films of all ranks per frame, in place, stream-ordered behind the frame's last kernel:
 * root < 0: ncclAllReduce, every rank gets the frame; root >= 0: ncclReduce, only rank `root` does (half the traffic;
 * the other ranks' films are left as they were).  The communicator's asynchronous error state is checked while the
 * call waits and after it: a peer that died turns into SPB_ERR_CUDA here (and the communicator is aborted), never into
 * a hang or a half-summed frame.  spb_film_reduce_async queues the reduction behind spb_render_samples_async calls. */
int spb_film_reduce(spb_ctx* ctx, int32_t root);
int spb_film_reduce_async(spb_ctx* ctx, int32_t root);
int spb_film_allreduce(spb_ctx* ctx);   /* = spb_film_reduce(ctx, -1)                                  */
/* ONE process driving several GPUs needs no communicator: sums the films of `others` into `root`'s film with one kernel
 * on the root's GPU that reads the other films through peer memory (NVLink / NVSwitch; a staging copy between GPUs that
 * are not peers; plain loads for a context on the same GPU).  Waits for everything queued on all the contexts first;
 * the other films are left as they were; every film must have the root's size.  spb_render_stats.reduce_ms of the root
 * counts the kernel.  (Replaces, like spb_film_reduce, the tile merge into the one Film of SamplerIntegrator::render,
 * core/integrator.cc:64-105, when the samples are spread over GPUs.) */
int spb_film_reduce_peers(spb_ctx* root, spb_ctx* const* others, int32_t n_others);
/* The same sum ACROSS PROCESSES (one process per GPU on one node): every other rank exports a handle of its film (CUDA IPC)
 * after spb_render_begin and ships the bytes to the root by its own means; the root maps them once per job
 * (spb_film_import_handles; SPB_ERR_UNSUPPORTED when the GPUs are not peers or a handle comes from this very process:
 * use spb_film_reduce then) and sums the mapped films into its own with the kernel of spb_film_reduce_peers
 * (spb_film_reduce_imported).  ORDERING IS THE CALLER'S: between the other ranks' last spb_render_samples (synchronous, or
 * waited for) and the root's spb_film_reduce_imported there must be a barrier of the job (MPI_Barrier,
 * torch.distributed.barrier), and another one before the other ranks touch their films again.  A handle dies with its
 * film: export again after a spb_render_begin that changes the film size. */
#define SPB_FILM_HANDLE_BYTES 96
int spb_film_export_handle(spb_ctx* ctx, char handle[SPB_FILM_HANDLE_BYTES]);
int spb_film_import_handles(spb_ctx* root, const char* handles /* n x SPB_FILM_HANDLE_BYTES */, int32_t n);
int spb_film_reduce_imported(spb_ctx* root);
int spb_comm_destroy(spb_ctx* ctx);

/* ---- acceleration structure ---------------------------------------------------------------- */

enum {
    SPB_BUILDER_DEVICE_SAH = 0,  /* default: top-down binned SAH (32 bins x 3 axes) + cost-optimal 8-wide collapse + quantisation, all on
                                  * the device; nodes and triangles never leave HBM.  The same tree as the host builder (2), byte for byte. */
    SPB_BUILDER_LBVH = 1,        /* Morton sort + Karras hierarchy + refit on the device, collapsed on the host: fastest binary build,
                                  * worst trees; kept for measurement                                                                */
    SPB_BUILDER_HOST_SAH = 2     /* the same binned SAH + collapse on the host cores (also used when sah_bins != 32 or SPICA_BVH_COLLAPSE=0) */
};
typedef struct spb_build_opts {
    int32_t builder;        /* SPB_BUILDER_*                                                      */
    int32_t max_leaf_tris;  /* 1..3, 0 = default (3)                                             */
    int32_t sah_bins;       /* default 32                                                       */
    int32_t reserved_;
} spb_build_opts;

/* Replaces BVHAccel::construct / constructRec (accelerators/bvh.cc:139-237). opts may be NULL.
 * Builds a binary SAH tree (1 primitive per leaf, like the reference), collapses it to the 8-wide compressed
 * layout and leaves it in HBM. */
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
    double inflate;         /* conservative slack of the quantised boxes (scene size x 2^-19)    */
    int32_t builder;        /* SPB_BUILDER_* that made this tree, -1: adopted (import_wide / clone) */
    int32_t reserved_;
} spb_bvh_stats;
int spb_bvh_get_stats(const spb_ctx* ctx, spb_bvh_stats* out);

/* One built tree for many contexts (replicas of a multi-GPU job must not each rebuild it).
 *   spb_bvh_export      copies the 8-wide BVH out of HBM: `nodes` (spb_bvh_stats::node_bytes) and `tris` (tri_bytes).
 *   spb_bvh_import_wide adopts an exported tree: `stats` is the exporter's spb_bvh_get_stats, unchanged; the context must have
 *                       been given the same triangles (spb_scene_set_triangles) -- they are what the integrator shades.
 *   spb_ctx_clone_scene same process: everything `src` holds -- triangles, attributes, the built tree, materials, lights,
 *                       textures, environment -- is replicated on `dst`'s GPU; the tree goes device to device (NVLink peer copy). */
int spb_bvh_export(spb_ctx* ctx, void* nodes, size_t node_capacity_bytes, void* tris, size_t tri_capacity_bytes);
int spb_bvh_import_wide(spb_ctx* ctx, const spb_bvh_stats* stats, const void* nodes, const void* tris);
int spb_ctx_clone_scene(spb_ctx* dst, spb_ctx* src);

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
 * "trace_variant" (traversal kernel: 0 one thread per ray, 1 persistent CTAs with dynamic refill,
 * 2 = 1 + warp-pooled float32 pre-test, 3 = 2 in visit / select / triangles order, 4 = 3 with the
 * stack in shared memory, 5 = 4 with three node visits per pooled triangle phase (default); every
 * variant returns the same records), "chunk_rays" (rays per pipelined chunk of the host-buffer calls,
 * default 524288), "wave_slots" (capacity of the integrator's queues = paths in flight, 240 B each, default 32 Mi but never more than 64 samples of the image, takes effect at
 * the next spb_render_begin), "render_graph" (0 = plain launches instead of the CUDA graph),
 * "shade_minb" (4/5/6: occupancy the Lambertian-only shade instance is compiled for),
 * "release_scratch" (any value: frees what the context keeps between calls to make them cheap -- the device builder's work
 * arena, about 460 B per triangle of the largest build, and the process-wide cache of freed queue memory of this GPU).
 * Unknown names return SPB_ERR_INVALID.
 * Environment read by spb_bvh_build / spb_bvh_import_binary: SPICA_BVH_COLLAPSE=0 selects the greedy
 * 8-wide collapse instead of the cost-optimal one (host builder). */
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
