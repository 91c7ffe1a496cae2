// The two plugins that carry the hot path to the GPU, registered under the reference's names:
//   accelerator "bvh"  (accelerators/bvh.h:57-97)   -> spb_scene_set_triangles + spb_bvh_build, and the
//                                                      scalar intersect() calls answered from the same tree
//   integrator  "path" (integrators/path/path.h:21-41) -> spb_render_* ; the finished image goes back
//                                                      through Film::setImage + Film::save (core/film.cc:23-63)
// Nothing is computed on the CPU here; a missing device is a FatalError (no fallback).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <ctime>
#include <map>
#include <mutex>
#include <thread>

#include "core.h"
#include "plugins.h"

namespace spica {

namespace {
// SPICA_TIMING=1: wall-clock phases of the GPU plugins on stdout ([TIME] lines), measured from process start
struct PhaseClock {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), last = t0;
    bool on = getenv("SPICA_TIMING") != nullptr;
    void lap(const char* what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        printf("[TIME] %-34s %8.3f s  (+%.3f)\n", what, std::chrono::duration<double>(now - t0).count(), std::chrono::duration<double>(now - last).count());
        last = now;
    }
};
PhaseClock& phaseClock() { static PhaseClock c; return c; }

void check(spb_ctx* ctx, int rc, const char* what) {
    if (rc != SPB_OK) FatalError("%s failed (%d): %s", what, rc, spb_last_error(ctx));
}

// Everything about the GPUs that does not depend on the scene -- CUDA start-up (1.6 s and more on a box without the
// persistence daemon), one context per GPU, the NCCL communicator -- is started on background threads the moment the number
// of GPUs is known (hostGpuPrewarm, called from main() before the scene file is even opened) and picked up when needed.
// How the films of several GPUs are summed: over peer memory by default (spb_film_reduce_peers: one process owns all the
// contexts, nothing to bring up); SPICA_FILM_REDUCE=nccl takes the communicator path a process-per-GPU job uses.
static bool filmReduceNccl() {
    const char* e = getenv("SPICA_FILM_REDUCE");
    return e && std::strcmp(e, "nccl") == 0;
}

struct GpuPool {
    int first = 0, G = 0;
    bool started = false;
    char commId[SPB_COMM_ID_BYTES];
    std::vector<std::thread> th;
    std::vector<spb_ctx*> ctx;
    std::vector<int> ctxDone, commDone;         // guarded by mu
    std::mutex mu;
    std::condition_variable cv;
    ~GpuPool() { for (auto& t : th) if (t.joinable()) t.join(); }
};
GpuPool& gpuPool() { static GpuPool p; return p; }

void gpuPoolStart(int firstDevice, int G) {
    GpuPool& p = gpuPool();
    if (p.started) return;
    p.started = true; p.first = firstDevice; p.G = G;
    p.ctx.assign(G, nullptr); p.ctxDone.assign(G, 0); p.commDone.assign(G, 0);
    for (int g = 0; g < G; g++) {
        p.th.emplace_back([&p, g]() {
            spb_ctx* c = nullptr;
            check(nullptr, spb_ctx_create(p.first + g, &c), "spb_ctx_create");
            { std::lock_guard<std::mutex> lk(p.mu); p.ctx[g] = c; p.ctxDone[g] = 1; }
            p.cv.notify_all();
            if (p.G > 1 && filmReduceNccl()) {
                if (g == 0) {
                    check(nullptr, spb_comm_get_unique_id(p.commId), "spb_comm_get_unique_id");
                    { std::lock_guard<std::mutex> lk(p.mu); p.commDone[0] = -1; }       // the id exists
                    p.cv.notify_all();
                } else {
                    std::unique_lock<std::mutex> lk(p.mu);
                    p.cv.wait(lk, [&] { return p.commDone[0] != 0; });
                }
                check(c, spb_comm_init(c, p.commId, p.G, g), "spb_comm_init");
                { std::lock_guard<std::mutex> lk(p.mu); p.commDone[g] = 1; }
                p.cv.notify_all();
            }
        });
    }
}
// the context of GPU `first + g` (created here when nothing was pre-warmed, e.g. a plain library user)
spb_ctx* gpuPoolContext(int device, int g) {
    GpuPool& p = gpuPool();
    if (!p.started || device != p.first || g >= p.G) {
        spb_ctx* c = nullptr;
        check(nullptr, spb_ctx_create(device + g, &c), "spb_ctx_create");
        return c;
    }
    std::unique_lock<std::mutex> lk(p.mu);
    p.cv.wait(lk, [&] { return p.ctxDone[g] != 0; });
    return p.ctx[g];
}
// true when the pool brought (or is bringing) the communicator up for this job; waits for rank g's part of it
bool gpuPoolCommReady(int device, int G, int g) {
    GpuPool& p = gpuPool();
    if (!p.started || device != p.first || G != p.G || G < 2 || !filmReduceNccl()) return false;
    std::unique_lock<std::mutex> lk(p.mu);
    p.cv.wait(lk, [&] { return p.commDone[g] == 1; });
    return true;
}
void gpuPoolFinish() {
    GpuPool& p = gpuPool();
    for (auto& t : p.th) if (t.joinable()) t.join();
    p.th.clear(); p.started = false;
}

struct FlatScene {              // the primitives as the C ABI takes them
    RawArray<double> verts;                 // 9 per triangle
    RawArray<float> normals, uvs;           // 9 / 6 per triangle, only when some triangle carries normals
    std::vector<int32_t> material_id, light_id;
    std::vector<spb_material> materials;
    std::vector<spb_texture> textures;      // + the bitmaps' texels and {kr, kt} bindings per material
    std::vector<float> texels;
    std::vector<int32_t> matTex;
    bool anyNormals = false, anyUV = false;
};

void flatten(const std::vector<std::shared_ptr<Primitive>>& prims, FlatScene* f) {
    const size_t n = prims.size();
    std::atomic<bool> anyN(false), anyUV(false);
    parallelFor(n, [&](size_t b, size_t e) { bool h = false; for (size_t i = b; i < e; i++) h = h || prims[i]->tri.hasNormals; if (h) anyN = true; });
    const bool withN = anyN;
    f->verts.resize(n * 9);
    if (withN) { f->normals.resize(n * 9); f->uvs.resize(n * 6); } else { f->normals.resize(0); f->uvs.resize(0); }
    f->material_id.resize(n); f->light_id.assign(n, -1);
    std::map<const SurfaceMaterial*, int> matIndex;
    std::map<const Texture*, int> texIndex;
    auto bindTexture = [&](const Texture* t) -> int32_t {
        if (!t) return -1;
        auto it = texIndex.find(t);
        if (it == texIndex.end()) {
            spb_texture d; t->describe(&d, &f->texels);
            it = texIndex.emplace(t, (int)f->textures.size()).first;
            f->textures.push_back(d);
        }
        return it->second;
    };
    // materials and textures in order of first appearance (serial: a pointer comparison per triangle), then the arrays in parallel
    const SurfaceMaterial* last = nullptr; int lastIndex = -1;
    for (size_t i = 0; i < n; i++) {
        const SurfaceMaterial* m = prims[i]->material.get();
        if (!m) { f->material_id[i] = -1; continue; }            // no bsdf: the path passes through (path.cc:71-75)
        if (m != last) {
            auto it = matIndex.find(m);
            if (it == matIndex.end()) {
                spb_material d; m->describe(&d);
                const Texture *tkr, *tkt; m->textures(&tkr, &tkt);
                it = matIndex.emplace(m, (int)f->materials.size()).first;
                f->materials.push_back(d);
                f->matTex.push_back(bindTexture(tkr)); f->matTex.push_back(bindTexture(tkt));
            }
            last = m; lastIndex = it->second;
        }
        f->material_id[i] = lastIndex;
    }
    parallelFor(n, [&](size_t b, size_t e) {
        bool hasUV = false;
        for (size_t i = b; i < e; i++) {
            const Primitive& p = *prims[i];
            for (int k = 0; k < 3; k++) { f->verts[i * 9 + k * 3] = p.tri.p[k].x; f->verts[i * 9 + k * 3 + 1] = p.tri.p[k].y; f->verts[i * 9 + k * 3 + 2] = p.tri.p[k].z; }
            if (!withN) continue;
            if (p.tri.hasNormals) {
                for (int k = 0; k < 3; k++) { f->normals[i * 9 + k * 3] = (float)p.tri.n[k].x; f->normals[i * 9 + k * 3 + 1] = (float)p.tri.n[k].y; f->normals[i * 9 + k * 3 + 2] = (float)p.tri.n[k].z; }
                for (int k = 0; k < 3; k++) { f->uvs[i * 6 + k * 2] = (float)p.tri.uv[k][0]; f->uvs[i * 6 + k * 2 + 1] = (float)p.tri.uv[k][1]; if (p.tri.uv[k][0] != 0.0 || p.tri.uv[k][1] != 0.0) hasUV = true; }
            } else {
                for (int k = 0; k < 9; k++) f->normals[i * 9 + k] = 0.f;
                for (int k = 0; k < 6; k++) f->uvs[i * 6 + k] = 0.f;
            }
        }
        if (hasUV) anyUV = true;
    });
    f->anyNormals = withN; f->anyUV = anyUV;
}

void uploadGeometry(spb_ctx* ctx, const FlatScene& f) {
    const int64_t n = (int64_t)f.material_id.size();
    check(ctx, spb_scene_set_triangles(ctx, f.verts.data(), f.anyNormals ? f.normals.data() : nullptr, f.anyUV ? f.uvs.data() : nullptr,
                                       f.material_id.data(), f.light_id.data(), n), "spb_scene_set_triangles");
    spb_build_opts opts; std::memset(&opts, 0, sizeof(opts));
    if (const char* b = getenv("SPICA_BVH_BUILDER")) opts.builder = atoi(b);    // SPB_BUILDER_*: 0 device binned SAH (default), 1 device LBVH, 2 host binned SAH
    if (const char* b = getenv("SPICA_BVH_MAX_LEAF")) opts.max_leaf_tris = atoi(b);   // 1..3 triangles per leaf (default 3)
    check(ctx, spb_bvh_build(ctx, &opts), "spb_bvh_build");
}
}  // namespace

class BVHAccel : public Accelerator {
public:
    BVHAccel(const std::vector<std::shared_ptr<Primitive>>& prims, RenderParams& params) : Accelerator(prims) {
        params.getBool("useSIMD", false, true);                         // accelerators/bvh.cc:127-131 (no meaning on the GPU)
        const char* dev = getenv("SPICA_DEVICE");
        device_ = dev ? atoi(dev) : 0;
        phaseClock().lap("scene parsed, accelerator ctor");
        ctx_ = gpuPoolContext(device_, 0);
        phaseClock().lap("context of GPU 0 ready");
        flatten(prims, &flat_);
        uploadGeometry(ctx_, flat_);
        phaseClock().lap("flatten + set_triangles + build");
        spb_bvh_stats st;
        check(ctx_, spb_bvh_get_stats(ctx_, &st), "spb_bvh_get_stats");
        for (int k = 0; k < 3; k++) { lo_[k] = st.world_lo[k]; hi_[k] = st.world_hi[k]; }
        MsgInfo("BVH: %lld triangles, %lld wide nodes, built in %.3f s", (long long)st.n_tris, (long long)st.n_wide_nodes, st.build_seconds);
    }
    ~BVHAccel() override { spb_ctx_destroy(ctx_); }
    // scalar queries for callers outside the GPU integrator: same tree, same exact arithmetic
    bool intersect(Ray& ray, SurfaceInteraction* isect) const override {
        std::lock_guard<std::mutex> lk(mu_);
        const spb_ray_f64 r = {ray.org.x, ray.org.y, ray.org.z, ray.dir.x, ray.dir.y, ray.dir.z, 0.0, ray.maxDist};
        spb_hit_f64 h;
        check(ctx_, spb_trace_closest_f64(ctx_, &r, 1, &h), "spb_trace_closest_f64");
        if (h.prim < 0) return false;
        ray.maxDist = h.t;                                              // core/primitive.cc:52
        if (isect) { isect->pos = ray.org + ray.dir * h.t; isect->u = h.u; isect->v = h.v; isect->primitive = h.prim; }
        return true;
    }
    bool intersect(Ray& ray) const override {
        std::lock_guard<std::mutex> lk(mu_);
        const spb_ray_f64 r = {ray.org.x, ray.org.y, ray.org.z, ray.dir.x, ray.dir.y, ray.dir.z, 0.0, ray.maxDist};
        uint8_t occ = 0;
        check(ctx_, spb_trace_any_f64(ctx_, &r, 1, &occ), "spb_trace_any_f64");
        return occ != 0;
    }
    void worldBound(double lo[3], double hi[3]) const override { for (int k = 0; k < 3; k++) { lo[k] = lo_[k]; hi[k] = hi_[k]; } }
    spb_ctx* ctx() const { return ctx_; }
    int device() const { return device_; }
    FlatScene& flat() { return flat_; }
private:
    spb_ctx* ctx_ = nullptr;
    int device_ = 0;
    FlatScene flat_;
    double lo_[3], hi_[3];
    mutable std::mutex mu_;
};

class PathIntegrator : public Integrator {
public:
    // path.cc:34-36; directlighting.cc:17-19 (which consumes the sampler, remove = true)
    PathIntegrator(RenderParams& params, int mode) : sampler_(std::static_pointer_cast<Sampler>(params.getObject("sampler", mode == SPB_INTEGRATOR_DIRECT))), mode_(mode) {}

    void render(const std::shared_ptr<const Camera>& camera, const Scene& scene, RenderParams& params) override {
        auto* accel = dynamic_cast<BVHAccel*>(scene.accelerator().get());
        if (!accel) FatalError("the GPU path integrator needs the GPU `bvh` accelerator");
        const HostOptions& opt = hostOptions();
        Film& film = *camera->film;
        const int width = film.width(), height = film.height();
        const int numSamples = opt.sppOverride > 0 ? opt.sppOverride : params.getInt("sampleCount");   // core/integrator.cc:63
        const int maxDepth = params.getInt("maxDepth");                    // path.cc:68
        FlatScene& flat = accel->flat();

        // lights in Scene::lights() order (sceneparser.cc:172-179,329)
        std::vector<spb_light> lights;
        std::map<const Light*, int> lightIndex;
        const Envmap* env = nullptr;
        for (const auto& l : scene.lights()) {
            spb_light d; std::memset(&d, 0, sizeof(d));
            d.type = l->kind(); d.prim = -1;
            if (l->kind() == SPB_LIGHT_AREA) { const auto* a = static_cast<const AreaLight*>(l.get()); d.radiance[0] = (float)a->Lemit.r; d.radiance[1] = (float)a->Lemit.g; d.radiance[2] = (float)a->Lemit.b; }
            else { if (env) FatalError("more than one environment map is outside this host's scope"); env = static_cast<const Envmap*>(l.get()); }
            lightIndex[l.get()] = (int)lights.size();
            lights.push_back(d);
        }
        const auto& prims = accel->primitives();
        for (size_t i = 0; i < prims.size(); i++) {
            flat.light_id[i] = -1;
            if (prims[i]->light) { const int li = lightIndex.at(prims[i]->light.get()); flat.light_id[i] = li; lights[li].prim = (int32_t)i; }
        }

        spb_render_desc desc; std::memset(&desc, 0, sizeof(desc));
        desc.width = width; desc.height = height; desc.max_depth = maxDepth;
        desc.filter = film.filter()->kind();
        desc.filter_radius[0] = film.filter()->rx; desc.filter_radius[1] = film.filter()->ry; desc.filter_sigma = film.filter()->sigma;
        std::memcpy(desc.camera_to_world, camera->cameraToWorld.getMat().m, sizeof(double) * 16);
        std::memcpy(desc.raster_to_camera, camera->rasterToCamera.getMat().m, sizeof(double) * 16);
        desc.lens_radius = camera->lensRadius; desc.focal_distance = camera->focalLength;
        desc.seed = opt.seed ? opt.seed : (uint64_t)time(nullptr);         // the reference seeds from time(0) (integrator.cc:51)
        desc.rr_start_bounce = 3;                                          // path.cc:117
        desc.integrator = mode_;

        const int G = std::max(1, opt.gpus);
        std::vector<spb_ctx*> ctxs(G, nullptr);
        ctxs[0] = accel->ctx();
        const bool nccl = G > 1 && filmReduceNccl();
        const bool pooledComm = nccl && gpuPool().started && gpuPool().first == accel->device() && gpuPool().G == G;
        char commId[SPB_COMM_ID_BYTES];
        if (nccl && !pooledComm) check(nullptr, spb_comm_get_unique_id(commId), "spb_comm_get_unique_id");

        // replicas: scene + BVH on every GPU; the tree is built ONCE (by the accelerator) and copied device to device
        auto replicate = [&](int g) {
            if (g == 0) return;
            ctxs[g] = gpuPoolContext(accel->device(), g);
            check(ctxs[g], spb_ctx_clone_scene(ctxs[g], ctxs[0]), "spb_ctx_clone_scene");
        };
        auto setup = [&](int g) {
            spb_ctx* ctx = ctxs[g];
            check(ctx, spb_scene_set_triangle_attributes(ctx, flat.material_id.data(), flat.light_id.data(), (int64_t)flat.material_id.size()), "spb_scene_set_triangle_attributes");
            check(ctx, spb_scene_set_materials(ctx, flat.materials.data(), (int32_t)flat.materials.size()), "spb_scene_set_materials");
            if (!flat.textures.empty()) {
                check(ctx, spb_scene_set_textures(ctx, flat.textures.data(), (int32_t)flat.textures.size(), flat.texels.data(), (int64_t)(flat.texels.size() / 3)), "spb_scene_set_textures");
                check(ctx, spb_scene_set_material_textures(ctx, flat.matTex.data(), (int32_t)flat.materials.size()), "spb_scene_set_material_textures");
            }
            check(ctx, spb_scene_set_lights(ctx, lights.data(), (int32_t)lights.size()), "spb_scene_set_lights");
            if (env) {
                std::vector<float> rgb(env->image.rgb.begin(), env->image.rgb.end());
                const double c[3] = {env->worldCenter.x, env->worldCenter.y, env->worldCenter.z};
                check(ctx, spb_scene_set_envmap(ctx, rgb.data(), env->image.width, env->image.height, &env->toWorld.getMat().m[0][0], env->scale, c, env->worldRadius), "spb_scene_set_envmap");
            }
            if (const char* ws = getenv("SPICA_WAVE_SLOTS")) check(ctx, spb_set_option(ctx, "wave_slots", atoll(ws)), "spb_set_option(wave_slots)");
            check(ctx, spb_render_begin(ctx, &desc), "spb_render_begin");
        };
        auto forEachGpu = [&](const std::function<void(int)>& fn) {
            std::vector<std::thread> th;
            for (int g = 1; g < G; g++) th.emplace_back(fn, g);
            fn(0);
            for (auto& t : th) t.join();
        };
        phaseClock().lap("integrator: lights resolved");
        forEachGpu(replicate);
        phaseClock().lap("replicas: ctx_create + clone");
        // (SPICA_FILM_REDUCE=nccl) the communicator takes a second or two to come up: it does so on its own threads while the scene is uploaded
        std::vector<std::thread> commThreads;
        if (nccl && !pooledComm) for (int g = 0; g < G; g++) commThreads.emplace_back([&, g]() { check(ctxs[g], spb_comm_init(ctxs[g], commId, G, g), "spb_comm_init"); });
        forEachGpu(setup);
        for (auto& t : commThreads) t.join();
        if (pooledComm) for (int g = 0; g < G; g++) gpuPoolCommReady(accel->device(), G, g);
        phaseClock().lap(nccl ? "scene upload, comm_init, begin" : "scene upload, begin");

        std::vector<float> rgb((size_t)width * height * 3);
        auto publish = [&](int id) {
            // output stage on the device when nobody needs the float image: only the encoded pixels cross PCIe
            if (!opt.onImage && !getenv("SPICA_HOST_ENCODE") && film.saveFromDevice(ctxs[0], id)) return;
            check(ctxs[0], spb_film_resolve(ctxs[0], rgb.data()), "spb_film_resolve");
            Image img(width, height);
            for (size_t i = 0; i < rgb.size(); i++) img.rgb[i] = rgb[i];
            film.setImage(img);                                             // core/film.cc:60-63
            film.save(id);
            if (opt.onImage) opt.onImage(img);
        };

        const auto t0 = std::chrono::steady_clock::now();
        if (opt.savePasses && G == 1) {
            // reference behaviour: one film save per spp pass (core/integrator.cc:64-105)
            for (int i = 0; i < numSamples; i++) {
                check(ctxs[0], spb_render_samples(ctxs[0], i, 1, 1), "spb_render_samples");
                printf("[ %d / %d ] 100.00 %% processed...\n", i + 1, numSamples);
                publish(i + 1);
            }
        } else {
            // GPU g renders sample indices g, g+G, ... ; then ONE sum of the RGBW films into GPU 0's: a kernel on GPU 0 reading
            // the other films over NVLink peer memory (or, SPICA_FILM_REDUCE=nccl, an ncclReduce)
            std::vector<double> tRender(G, 0.0), tReduce(G, 0.0);
            forEachGpu([&](int g) {
                const int count = (numSamples - g + G - 1) / G;
                const auto a = std::chrono::steady_clock::now();
                check(ctxs[g], spb_render_samples(ctxs[g], g, std::max(count, 0), G), "spb_render_samples");
                const auto b = std::chrono::steady_clock::now();
                if (nccl) check(ctxs[g], spb_film_reduce(ctxs[g], 0), "spb_film_reduce");      // only GPU 0 publishes the frame: ncclReduce, half the traffic of an all-reduce
                tRender[g] = std::chrono::duration<double>(b - a).count();
                tReduce[g] = std::chrono::duration<double>(std::chrono::steady_clock::now() - b).count();   // includes waiting for the slowest GPU
            });
            if (G > 1 && !nccl) {
                const auto b = std::chrono::steady_clock::now();
                check(ctxs[0], spb_film_reduce_peers(ctxs[0], ctxs.data() + 1, G - 1), "spb_film_reduce_peers");
                const double r = std::chrono::duration<double>(std::chrono::steady_clock::now() - b).count();
                for (int g = 0; g < G; g++) tReduce[g] = r;
            }
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            double rmax = 0, rmin = 1e30, amin = 1e30;
            if (G > 1) {
                for (int g = 0; g < G; g++) { rmax = std::max(rmax, tRender[g]); rmin = std::min(rmin, tRender[g]); amin = std::min(amin, tReduce[g]); }
            }
            spb_render_stats st;
            check(ctxs[0], spb_render_get_stats(ctxs[0], &st), "spb_render_get_stats");
            if (G > 1) MsgInfo("per-GPU render %.3f .. %.3f s, film sum (%s) %.4f s on the host's clock, %.3f ms on GPU 0", rmin, rmax, nccl ? "ncclReduce" : "peer memory", amin, st.reduce_ms);
            MsgInfo("rendered %d spp at %dx%d on %d GPU(s) in %.3f s: %.2f Msamples/s; GPU0: %.1f Mrays/s", numSamples, width, height, G, sec,
                    1e-6 * width * height * (double)numSamples / sec,
                    st.render_ms > 0 ? 1e-3 * (double)(st.rays_closest + st.rays_shadow + st.rays_mis) / st.render_ms : 0.0);
            phaseClock().lap("render + reduce");
            publish(numSamples);
            phaseClock().lap("publish (resolve, encode, save)");
        }
        forEachGpu([&](int g) { if (nccl) spb_comm_destroy(ctxs[g]); if (g > 0) spb_ctx_destroy(ctxs[g]); });
        phaseClock().lap("replicas destroyed");
        printf("Finish!!\n");
    }
private:
    std::shared_ptr<Sampler> sampler_;
    int mode_;
};

void hostPhaseLap(const char* what) { phaseClock().lap(what); }
void hostGpuPrewarm(int gpus) {
    const char* dev = getenv("SPICA_DEVICE");
    gpuPoolStart(dev ? atoi(dev) : 0, std::max(1, gpus));
}
void hostGpuFinish() { gpuPoolFinish(); }

void registerGpuPlugins() {
    PluginManager& pm = PluginManager::getInstance();
    pm.registerAccelerator("bvh", [](const std::vector<std::shared_ptr<Primitive>>& prims, RenderParams& p) -> Accelerator* { return new BVHAccel(prims, p); });
    pm.registerPlugin("path", [](RenderParams& p) -> CObject* { return new PathIntegrator(p, SPB_INTEGRATOR_PATH); });
    // integrators/directlighting: the same kernels without indirect light
    pm.registerPlugin("directlighting", [](RenderParams& p) -> CObject* { return new PathIntegrator(p, SPB_INTEGRATOR_DIRECT); });
}

}  // namespace spica
