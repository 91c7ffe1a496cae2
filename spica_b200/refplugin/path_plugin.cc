// plugins/path.so for the UNMODIFIED reference host: replaces integrators/path/path.cc behind the
// same factory symbols (core/cobject.h:56-64).  render() is overridden wholesale
// (core/integrator.h:29-37): the scene the reference parsed is resolved into the PODs of
// include/spica_b200.h once, every sample is traced and shaded on the GPU (spb_render_samples), and
// the finished image goes back through the reference's own Film::setImage + Film::save
// (core/film.cc:23-40,60-63), so the output file is written by the reference's film plugin.
//
// Environment (the reference CLI has no flags for these): SPICA_DEVICE (first GPU, default 0),
// SPICA_GPUS (G contexts, samples interleaved, the films summed over peer memory; SPICA_FILM_REDUCE=nccl: ncclReduce), SPICA_SEED
// (default time(0), as core/integrator.cc:51), SPICA_SPP (overrides sampleCount),
// SPICA_SAVE_PASSES=1 (save after every pass like core/integrator.cc:98).
#include <cstring>

#include "gpu_scene.h"

// the same source builds plugins/path.so (default) and plugins/directlighting.so (-DSPB_REFPLUGIN_INTEGRATOR=1,
// replaces integrators/directlighting/directlighting.cc)
#ifndef SPB_REFPLUGIN_INTEGRATOR
#define SPB_REFPLUGIN_INTEGRATOR SPB_INTEGRATOR_PATH
#endif

namespace spica {
namespace b200 {

static long envInt(const char* name, long dflt) { const char* v = getenv(name); return v && *v ? atol(v) : dflt; }

static void rowMajor(const Transform& t, double out[16]) {
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out[i * 4 + j] = t.getMat()(i, j);
}

class GpuPathIntegrator : public Integrator {
public:
    explicit GpuPathIntegrator(RenderParams& params)                                                    // path.cc:35-37, directlighting.cc:17-19
        : sampler_(std::static_pointer_cast<Sampler>(params.getObject("sampler", SPB_REFPLUGIN_INTEGRATOR == SPB_INTEGRATOR_DIRECT))) {}

    void render(const std::shared_ptr<const Camera>& camera, const Scene& scene, RenderParams& params) override {
        // the tree: shared with the GPU `bvh` accelerator when the scene uses it, else built here from the same primitives
        std::unique_ptr<GpuScene> own;
        GpuScene* gs = nullptr;
        const Accelerator* accel = scene.aggregate_.get();
        if (isType(accel, kGpuBVHAccelName)) gs = static_cast<const GpuBVHAccel*>(accel)->gpu();
        if (!gs) {
            if (scene.primitives().empty()) FatalError("the scene has no primitives");
            own = std::make_unique<GpuScene>(scene.primitives());
            gs = own.get();
        }
        FlatScene& flat = gs->flat;
        const auto& prims = scene.primitives();

        if (!isType(camera.get(), "N5spica17PerspectiveCameraE"))
            FatalError("camera %s is outside the GPU path's scope (perspective only)", typeid(*camera).name());
        Film& film = *camera->film_;
        const int width = film.resolution_.x(), height = film.resolution_.y();
        const int numSamples = (int)envInt("SPICA_SPP", params.getInt("sampleCount"));                  // core/integrator.cc:63
        const int maxDepth = params.getInt("maxDepth");                                                 // path.cc:68

        // lights in Scene::lights() order (spica/sceneparser.cc:172-179): the order matters for the uniform pick (core/mis.cc:27)
        std::vector<spb_light> lights;
        std::map<const Light*, int> lightIndex;
        const Envmap* env = nullptr;
        for (const auto& l : scene.lights()) {
            spb_light d; std::memset(&d, 0, sizeof(d));
            d.prim = -1;
            if (isType(l.get(), "N5spica9AreaLightE")) {
                d.type = SPB_LIGHT_AREA; setv(d.radiance, static_cast<const AreaLight*>(l.get())->Lemit_);
            } else if (isType(l.get(), "N5spica6EnvmapE")) {
                if (env) FatalError("more than one environment map is outside the GPU path's scope");
                d.type = SPB_LIGHT_ENVMAP; env = static_cast<const Envmap*>(l.get());
            } else {
                FatalError("emitter %s is outside the GPU path's scope (area, envmap)", typeid(*l).name());
            }
            lightIndex[l.get()] = (int)lights.size();
            lights.push_back(d);
        }
        for (size_t i = 0; i < prims.size(); i++) {
            flat.light_id[i] = -1;
            if (const Light* l = prims[i]->light()) {
                auto it = lightIndex.find(l);
                if (it == lightIndex.end()) FatalError("primitive %zu emits through a light that is not in the scene's list", i);
                flat.light_id[i] = it->second; lights[it->second].prim = (int32_t)i;
            }
        }

        spb_render_desc desc; std::memset(&desc, 0, sizeof(desc));
        desc.width = width; desc.height = height; desc.max_depth = maxDepth;
        const Filter* filter = film.filter_.get();
        desc.filter_radius[0] = filter->radius_.x(); desc.filter_radius[1] = filter->radius_.y();
        if (isType(filter, "N5spica9BoxFilterE")) desc.filter = SPB_FILTER_BOX;
        else if (isType(filter, "N5spica10TentFilterE")) desc.filter = SPB_FILTER_TENT;
        else if (isType(filter, "N5spica14GaussianFilterE")) { desc.filter = SPB_FILTER_GAUSSIAN; desc.filter_sigma = 1.0 / static_cast<const GaussianFilter*>(filter)->beta_; }
        else FatalError("rfilter %s is outside the GPU path's scope (box, tent, gaussian)", typeid(*filter).name());
        rowMajor(camera->cameraToWorld_, desc.camera_to_world);
        rowMajor(camera->rasterToCamera_, desc.raster_to_camera);
        desc.lens_radius = camera->lensRadius_; desc.focal_distance = camera->focalLength_;
        desc.seed = (uint64_t)envInt("SPICA_SEED", (long)time(nullptr));                               // core/integrator.cc:51
        desc.rr_start_bounce = 3;                                                                       // path.cc:117
        desc.integrator = SPB_REFPLUGIN_INTEGRATOR;

        // environment map: level 0 of the reference's pyramid is the scaled image (lights/envmap.cc:27-33);
        // Light keeps the transposed XML matrix (envmap.cc:19), the C ABI takes the XML one
        std::vector<float> envRgb; int envW = 0, envH = 0; double envL2W[16], envCenter[3] = {0, 0, 0};
        if (env) {
            const Image& im = env->mipmap_->pyramid_[0];
            envW = im.width(); envH = im.height();
            envRgb.resize((size_t)envW * envH * 3);
            for (int y = 0; y < envH; y++) for (int x = 0; x < envW; x++) {
                const RGBSpectrum& c = im(x, y);
                float* o = &envRgb[((size_t)y * envW + x) * 3];
                o[0] = (float)c.red(); o[1] = (float)c.green(); o[2] = (float)c.blue();
            }
            for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) envL2W[i * 4 + j] = env->lightToWorld_.getMat()(j, i);
            envCenter[0] = env->worldCenter_.x(); envCenter[1] = env->worldCenter_.y(); envCenter[2] = env->worldCenter_.z();
        }

        if (const char* path = dumpPath()) {
            FILE* fp = fopen(path, "wb");
            if (!fp) FatalError("cannot write %s", path);
            const int64_t hdr[8] = {(int64_t)prims.size(), (int64_t)flat.materials.size(), (int64_t)lights.size(), flat.anyNormals, flat.anyUV,
                                    numSamples, envW, envH};
            fwrite(hdr, sizeof(hdr), 1, fp);
            fwrite(flat.verts.data(), sizeof(double), flat.verts.size(), fp);
            fwrite(flat.normals.data(), sizeof(float), flat.normals.size(), fp);
            fwrite(flat.uvs.data(), sizeof(float), flat.uvs.size(), fp);
            fwrite(flat.material_id.data(), sizeof(int32_t), flat.material_id.size(), fp);
            fwrite(flat.light_id.data(), sizeof(int32_t), flat.light_id.size(), fp);
            fwrite(flat.materials.data(), sizeof(spb_material), flat.materials.size(), fp);
            fwrite(lights.data(), sizeof(spb_light), lights.size(), fp);
            fwrite(&desc, sizeof(desc), 1, fp);
            if (env) {
                const double tail[4] = {envCenter[0], envCenter[1], envCenter[2], env->worldRadius_};
                fwrite(envL2W, sizeof(double), 16, fp); fwrite(tail, sizeof(double), 4, fp);
                fwrite(envRgb.data(), sizeof(float), envRgb.size(), fp);
            }
            const int64_t thdr[2] = {(int64_t)flat.textures.size(), (int64_t)(flat.texels.size() / 3)};     // trailer: textures
            fwrite(thdr, sizeof(thdr), 1, fp);
            fwrite(flat.textures.data(), sizeof(spb_texture), flat.textures.size(), fp);
            fwrite(flat.matTex.data(), sizeof(int32_t), flat.matTex.size(), fp);
            fwrite(flat.texels.data(), sizeof(float), flat.texels.size(), fp);
            fclose(fp);
            MsgInfo("scene PODs written to %s; not rendering", path);
            return;
        }

        const int G = (int)std::max(1L, envInt("SPICA_GPUS", 1));
        std::vector<spb_ctx*> ctxs(G, nullptr);
        ctxs[0] = gs->ctx;
        const char* how = getenv("SPICA_FILM_REDUCE");
        const bool nccl = G > 1 && how && std::strcmp(how, "nccl") == 0;     // default: spb_film_reduce_peers, no communicator
        char commId[SPB_COMM_ID_BYTES];
        if (nccl) check(nullptr, spb_comm_get_unique_id(commId), "spb_comm_get_unique_id");

        // replicas: scene + BVH on every GPU; the tree is built ONCE (by the accelerator) and copied device to device
        auto replicate = [&](int g) {
            if (g == 0) return;
            check(nullptr, spb_ctx_create(gs->device + g, &ctxs[g]), "spb_ctx_create");
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
            if (env) check(ctx, spb_scene_set_envmap(ctx, envRgb.data(), envW, envH, envL2W, 1.0, envCenter, env->worldRadius_), "spb_scene_set_envmap");
            if (const char* ws = getenv("SPICA_WAVE_SLOTS")) check(ctx, spb_set_option(ctx, "wave_slots", atoll(ws)), "spb_set_option(wave_slots)");
            check(ctx, spb_render_begin(ctx, &desc), "spb_render_begin");
        };
        auto forEachGpu = [&](const std::function<void(int)>& fn) {
            std::vector<std::thread> th;
            for (int g = 1; g < G; g++) th.emplace_back(fn, g);
            fn(0);
            for (auto& t : th) t.join();
        };
        forEachGpu(replicate);
        // (SPICA_FILM_REDUCE=nccl) the communicator takes a second or two to come up: it does so on its own threads while the scene is uploaded
        std::vector<std::thread> commThreads;
        if (nccl) for (int g = 0; g < G; g++) commThreads.emplace_back([&, g]() { check(ctxs[g], spb_comm_init(ctxs[g], commId, G, g), "spb_comm_init"); });
        forEachGpu(setup);
        for (auto& t : commThreads) t.join();

        std::vector<float> rgb((size_t)width * height * 3);
        auto publish = [&](int id) {
            check(ctxs[0], spb_film_resolve(ctxs[0], rgb.data()), "spb_film_resolve");
            Image img(width, height);
            for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) {
                const float* p = &rgb[((size_t)y * width + x) * 3];
                img.pixel(x, y) = RGBSpectrum(p[0], p[1], p[2]);
            }
            film.setImage(img);                                             // core/film.cc:60-63 (weights := 1)
            film.save(id);                                                  // core/film.cc:23-40
        };

        const auto t0 = std::chrono::steady_clock::now();
        if (envInt("SPICA_SAVE_PASSES", 0) && G == 1) {
            for (int i = 0; i < numSamples; i++) {                          // one film save per pass (core/integrator.cc:64-105)
                check(ctxs[0], spb_render_samples(ctxs[0], i, 1, 1), "spb_render_samples");
                printf("[ %d / %d ] 100.00 %% processed...\n", i + 1, numSamples);
                publish(i + 1);
            }
        } else {
            // GPU g renders sample indices g, g+G, ... ; then ONE sum of the RGBW films into GPU 0's (it alone publishes the frame)
            forEachGpu([&](int g) {
                const int count = (numSamples - g + G - 1) / G;
                check(ctxs[g], spb_render_samples(ctxs[g], g, std::max(count, 0), G), "spb_render_samples");
                if (nccl) check(ctxs[g], spb_film_reduce(ctxs[g], 0), "spb_film_reduce");
            });
            if (G > 1 && !nccl) check(ctxs[0], spb_film_reduce_peers(ctxs[0], ctxs.data() + 1, G - 1), "spb_film_reduce_peers");
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            spb_render_stats st;
            check(ctxs[0], spb_render_get_stats(ctxs[0], &st), "spb_render_get_stats");
            MsgInfo("rendered %d spp at %dx%d on %d GPU(s) in %.3f s: %.2f Msamples/s; GPU0: %.1f Mrays/s", numSamples, width, height, G, sec,
                    1e-6 * width * height * (double)numSamples / sec,
                    st.render_ms > 0 ? 1e-3 * (double)(st.rays_closest + st.rays_shadow + st.rays_mis) / st.render_ms : 0.0);
            publish(numSamples);
        }
        forEachGpu([&](int g) { if (nccl) spb_comm_destroy(ctxs[g]); if (g > 0) spb_ctx_destroy(ctxs[g]); });
        printf("Finish!!\n");
    }

private:
    std::shared_ptr<Sampler> sampler_;      // kept alive like path.h:37; the GPU sampler is counter-based
};

}  // namespace b200
}  // namespace spica

extern "C" {
spica::CObject* createInstance(spica::RenderParams& params) { return (spica::CObject*)(new spica::b200::GpuPathIntegrator(params)); }
const char* getDescription() {
    return SPB_REFPLUGIN_INTEGRATOR == SPB_INTEGRATOR_DIRECT ? "B200 wavefront direct-lighting integrator"
                                                             : "B200 wavefront path tracer (unidirectional, next-event estimation + MIS)";
}
}
