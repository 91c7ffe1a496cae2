// Internal: the opaque context behind the C ABI (include/spica_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/spica_b200.h"
#include "bvh_host.h"
#include "bvh_layout.h"

namespace spb { struct RenderState; }

// The device builder's work arrays: a bump allocator over a few large plain allocations that the context keeps, so that a
// rebuild allocates nothing (26 driver allocations of odd sizes per build cost 3 .. 550 ms, whether from cudaMalloc or from a
// stream-ordered pool: profiles/r02ae_build_cold.txt, r02ag_build_cold.txt).  reset() at the start of a build; mark() / rewind()
// give the space of the arrays that die half-way to the second half.  spb_set_option(ctx, "release_scratch", 1) frees it.
struct ScratchArena {
    struct Block { char* p; size_t bytes; };
    std::vector<Block> blocks;
    size_t cur = 0, off = 0;
    struct Mark { size_t cur, off; };
    void reset() { cur = 0; off = 0; }
    Mark mark() const { return Mark{cur, off}; }
    void rewind(Mark m) { cur = m.cur; off = m.off; }
    cudaError_t alloc(void** out, size_t bytes, size_t growBytes) {
        bytes = (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
        for (; cur < blocks.size(); cur++, off = 0) {
            if (off + bytes <= blocks[cur].bytes) { *out = blocks[cur].p + off; off += bytes; return cudaSuccess; }
        }
        Block b{nullptr, std::max(bytes, growBytes)};
        cudaError_t e = cudaMalloc((void**)&b.p, b.bytes);
        if (e != cudaSuccess && b.bytes > bytes) { cudaGetLastError(); b.bytes = bytes; e = cudaMalloc((void**)&b.p, b.bytes); }
        if (e != cudaSuccess) { *out = nullptr; return e; }
        blocks.push_back(b);
        cur = blocks.size() - 1; off = bytes;
        *out = b.p;
        return cudaSuccess;
    }
    void release() { for (auto& b : blocks) cudaFree(b.p); blocks.clear(); reset(); }
};

struct spb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaStream_t kstream[4] = {};        // one per pipeline slot: the next chunk's CTAs fill SM slots as the previous kernel's tail drains
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;

    // host copy of the geometry: what the builders and the shading-record kernel are fed from.  Shared (not copied) between
    // a context and its clones (spb_ctx_clone_scene); replaced, never modified, by spb_scene_set_triangles.
    struct Geometry {
        std::vector<double>  verts;      // 9 per triangle
        std::vector<float>   normals;    // 9 per triangle or empty
        std::vector<float>   uvs;        // 6 per triangle or empty
    };
    std::shared_ptr<const Geometry> geo = std::make_shared<Geometry>();
    std::vector<int32_t> material_id, light_id;      // per context: the attributes may be re-set without touching the tree
    int64_t n_tris = 0;

    spb::BinaryBVH bin;
    spb::HostBVH   bvh;
    bool bvh_ready = false;
    double build_seconds = 0.0;
    int builder_used = 0;                // SPB_BUILDER_* of the current tree, -1: adopted (import / clone)

    void* d_nodes = nullptr;
    void* d_tris = nullptr;
    size_t d_nodes_bytes = 0, d_tris_bytes = 0;   // sizes of the two allocations (0 = unknown)
    void* spare_nodes = nullptr; void* spare_tris = nullptr; size_t spare_nodes_bytes = 0, spare_tris_bytes = 0;   // the previous tree's arrays during a device rebuild (re-used when large enough)
    ScratchArena build_arena;            // sah_build.cu's work arrays, kept between builds
    void* d_pre_tris = nullptr;          // TriF64 scenes: float32-rounded copy for the pre-test (SceneParams::pre_tris)
    spb::SceneParams sp{};

    // grow-only scratch for the host-buffer entry points (kPipe-deep ring)
    static constexpr int kPipe = 4;      // chunks in flight: H2D, kernel and D2H of neighbouring chunks overlap
    void* d_in[kPipe] = {};
    void* d_out[kPipe] = {};
    size_t in_cap = 0, out_cap = 0;
    cudaEvent_t ev_in[kPipe] = {}, ev_k[kPipe] = {}, ev_out[kPipe] = {};
    unsigned long long* d_work = nullptr;   // persistent-kernel work counter + traversal counters (4 x u64), one set per pipeline slot

    // options
    int opt_counters = 0;
    int opt_block = 128;
    int opt_ctas_per_sm = 0;             // 0 = occupancy query
    int opt_render_graph = 1;            // the integrator's iteration pair as a CUDA graph (0: plain launches; same image)
    int opt_shade_minb = 0;              // Lambertian-only shade instance compiled for 4, 5 or 6 resident CTAs per SM (0 = default, 4)
    int opt_variant = 5;                 // 0 = one thread per ray, 1 = persistent dynamic fetch, 2 = 1 + warp-cooperative pre-test, 3 = 2 in early-select order,
                                         // 4 = 3 with the stack in shared memory, 5 = 4 with three node visits per pooled triangle phase (default;
                                         // falls back to 3 on trees deeper than the shared stack and to 2 on float64 triangles)
    int64_t opt_chunk = 1 << 19;         // rays per pipelined chunk on the host-buffer path (measured: 256K 1412, 512K 1574, 1M 1403, 2M 1364 Mrays/s; PCIe floor 1574)
    int64_t opt_wave_slots = 1 << 25;    // capacity of the integrator's queues = paths in flight (240 B each); the streaming loop keeps them full, so the size only has to amortise the launches and kernel tails of an iteration: 8 Mi -> 2134, 16 Mi -> 2195, 32 Mi -> 2223 Msamples/s on the diffuse Cornell box (7.7 GB of the 180 GB)

    // resident CTAs per SM of each traversal kernel instantiation on THIS device (the occupancy query is slow enough to matter per chunk)
    std::unordered_map<const void*, int> occupancy;

    // counters
    double last_kernel_ms = 0.0;
    int64_t kernel_launches = 0;
    int64_t c_rays = 0, c_nodes = 0, c_tris = 0;

    // integrator state lives in integrator.cu
    spb::RenderState* render = nullptr;
};

namespace spb {
int  fail(spb_ctx* ctx, int code, const std::string& msg);
bool cudaOk(spb_ctx* ctx, cudaError_t e, const char* what);
void setGlobalError(const std::string& msg);
void renderStateEnsure(spb_ctx* ctx);    // integrator.cu: the render state exists from spb_ctx_create on (so that calls from two host threads never race to create it)
void renderStateDestroy(spb_ctx* ctx);   // integrator.cu
void renderSceneClone(spb_ctx* dst, spb_ctx* src);   // integrator.cu: materials, lights, textures, environment of src -> dst (host side; uploaded by the next spb_render_begin)
void renderSceneChanged(spb_ctx* ctx);   // integrator.cu: geometry / attributes changed, a new spb_render_begin is required
int  buildLbvhDevice(spb_ctx* ctx, BinaryBVH* out);   // lbvh.cu
// Scratch memory (the render queues): stream-ordered allocations from ONE pool per GPU and process
// that keeps what is freed (release threshold = max), so a new context re-uses the 8 GB of the one that went away instead of
// going through the driver's map / unmap (C1 through the host, a context per call: 0.033 / 0.67 / 0.055 s -> 0.027 / 0.028 / 0.025 s).
// spb_set_option(ctx, "release_scratch", 1) hands the cached memory back.  Not for memory another GPU or process maps.
cudaError_t scratchAlloc(spb_ctx* ctx, void** p, size_t bytes, cudaStream_t st);
void        scratchFree(void* p, cudaStream_t st);                    // p may be NULL
void        scratchRelease(int device);
int  buildSahDevice(spb_ctx* ctx, int maxLeaf);       // sah_build.cu: the default builder; leaves the 8-wide BVH in ctx->d_nodes / d_tris
}  // namespace spb

#define SPB_CUDA(ctx, call)                                                      \
    do {                                                                         \
        if (!spb::cudaOk((ctx), (call), #call)) return SPB_ERR_CUDA;             \
    } while (0)
