// Shared by plugins/bvh.so and plugins/path.so (the shim for the UNMODIFIED reference host):
// resolves the reference's live scene objects into the PODs of include/spica_b200.h and owns the
// GPU context.  Everything is computed by libspica_b200.so; there is no CPU path here -- a missing
// device or an object outside the path's scope is the reference's FatalError (print + abort,
// core/common.h:109-115).
#ifndef SPICA_B200_REFPLUGIN_GPU_SCENE_H_
#define SPICA_B200_REFPLUGIN_GPU_SCENE_H_

#include "ref_open.h"
#include "spica_b200.h"

namespace spica {
namespace b200 {

inline void check(spb_ctx* ctx, int rc, const char* what) {
    if (rc != SPB_OK) FatalError("%s failed (%d): %s", what, rc, spb_last_error(ctx));
}

// The plugins are dlopen'ed RTLD_LOCAL one by one (core/cobject.cc:32), so a plugin class has no
// linkable typeinfo from here: dynamic types are recognised by their mangled name (libstdc++
// compares typeinfo by name anyway) and then read through the layouts of the reference's headers.
template <class T>
inline bool isType(const T* obj, const char* mangled) { return obj && std::strcmp(typeid(*obj).name(), mangled) == 0; }

struct FlatScene {              // the primitives as spb_scene_set_triangles takes them
    std::vector<double> verts;
    std::vector<float> normals, uvs;
    std::vector<int32_t> material_id, light_id;
    std::vector<spb_material> materials;
    std::vector<spb_texture> textures;      // + the bitmaps' texels and {kr, kt} bindings per material
    std::vector<float> texels;
    std::vector<int32_t> matTex;
    std::map<const void*, int> texIndex;
    bool anyNormals = false, anyUV = false;
};

inline void setv(float dst[3], const Spectrum& s) { dst[0] = (float)s.red(); dst[1] = (float)s.green(); dst[2] = (float)s.blue(); }

// A reflectance-type parameter: constant, or a bitmap / checkerboard texture bound to the material slot
// (textures/bitmap.cc:16-26, textures/checkerboard.cc:23-40).  Returns the texture index or -1 (constant in *value).
inline int32_t reflectanceParam(FlatScene* f, const std::shared_ptr<Texture<Spectrum>>& tex, const char* what, float value[3]) {
    value[0] = value[1] = value[2] = 0.f;
    if (!tex) FatalError("%s is missing", what);
    if (isType(tex.get(), "N5spica15ConstantTextureINS_11RGBSpectrumEEE")) { setv(value, tex->evaluate(SurfaceInteraction())); return -1; }
    auto it = f->texIndex.find(tex.get());
    if (it != f->texIndex.end()) return it->second;
    spb_texture t; std::memset(&t, 0, sizeof(t));
    if (isType(tex.get(), "N5spica13BitmapTextureE")) {
        const auto* b = static_cast<const BitmapTexture*>(tex.get());
        if (!isType(b->texmap_.get(), "N5spica11UVMapping2DE")) FatalError("%s: only UV-mapped bitmaps are inside the GPU path's scope", what);
        const auto* m = static_cast<const UVMapping2D*>(b->texmap_.get());
        if (m->su_ != 1.0 || m->sv_ != 1.0 || m->du_ != 0.0 || m->dv_ != 0.0 || !m->invertHorizontal_) FatalError("%s: non-default UV mapping is outside the GPU path's scope", what);
        if (b->mipmap_->imageWrap_ != ImageWrap::Repeat) FatalError("%s: only the Repeat wrap mode is inside the GPU path's scope", what);
        const Image& im = b->mipmap_->pyramid_[0];
        t.type = SPB_TEX_BITMAP; t.width = im.width(); t.height = im.height(); t.texel_offset = (int64_t)(f->texels.size() / 3);
        for (int y = 0; y < im.height(); y++) for (int x = 0; x < im.width(); x++) {
            const RGBSpectrum& c = im(x, y);
            f->texels.push_back((float)c.red()); f->texels.push_back((float)c.green()); f->texels.push_back((float)c.blue());
        }
    } else if (isType(tex.get(), "N5spica12CheckerboardE")) {
        const auto* c = static_cast<const Checkerboard*>(tex.get());
        t.type = SPB_TEX_CHECKERBOARD; setv(t.color0, c->color0_); setv(t.color1, c->color1_);
        t.uoffset = (float)c->uOffset_; t.voffset = (float)c->vOffset_; t.uscale = (float)c->uScale_; t.vscale = (float)c->vScale_;
    } else {
        FatalError("%s: texture type %s is outside the GPU path's scope (constant, bitmap, checkerboard)", what, typeid(*tex).name());
    }
    const int32_t id = (int32_t)f->textures.size();
    f->textures.push_back(t);
    f->texIndex.emplace(tex.get(), id);
    return id;
}

// every other parameter must be constant
inline Spectrum constantValue(const std::shared_ptr<Texture<Spectrum>>& tex, const char* what) {
    if (!tex) FatalError("%s is missing", what);
    if (!isType(tex.get(), "N5spica15ConstantTextureINS_11RGBSpectrumEEE"))
        FatalError("%s: texture type %s is outside the GPU path's scope (constant textures only)", what, typeid(*tex).name());
    return tex->evaluate(SurfaceInteraction());
}
inline int distributionId(const std::string& name) {          // bsdfs/roughconductor.cc:62-69
    if (name == "beckmann") return SPB_DISTR_BECKMANN;
    if (name == "ggx") return SPB_DISTR_GGX;
    FatalError("Unknown microfacet distribution type: %s", name.c_str());
    return 0;
}

inline void describeMaterial(FlatScene* f, const SurfaceMaterial* sm, spb_material* m, int32_t tex[2]) {
    std::memset(m, 0, sizeof(*m));
    tex[0] = tex[1] = -1;
    auto noBump = [](const std::shared_ptr<Texture<Spectrum>>& b) { if (b) FatalError("bump maps are outside the GPU path's scope"); };
    if (isType(sm, "N5spica7DiffuseE")) {                                   // bsdfs/diffuse.cc:23-32
        const auto* d = static_cast<const Diffuse*>(sm);
        noBump(d->bumpMap_);
        m->type = SPB_MAT_DIFFUSE; tex[0] = reflectanceParam(f, d->Kd_, "diffuse reflectance", m->kr);
    } else if (isType(sm, "N5spica10DielectricE")) {                        // bsdfs/dielectric.cc:30-42
        const auto* d = static_cast<const Dielectric*>(sm);
        noBump(d->bumpMap_);
        m->type = SPB_MAT_DIELECTRIC;
        setv(m->kr, constantValue(d->Kr_, "specularReflectance")); setv(m->kt, constantValue(d->Kt_, "specularTransmittance"));
        m->eta[0] = m->eta[1] = m->eta[2] = (float)constantValue(d->index_, "intIOR").gray();
    } else if (isType(sm, "N5spica14RoughConductorE")) {                    // bsdfs/roughconductor.cc:38-75
        const auto* d = static_cast<const RoughConductor*>(sm);
        noBump(d->bumpMap_);
        if (d->remapRoughness_) FatalError("roughness remapping is outside the GPU path's scope");
        m->type = SPB_MAT_ROUGHCONDUCTOR; m->distribution = distributionId(d->distribution_);
        m->kr[0] = m->kr[1] = m->kr[2] = 1.f;
        setv(m->eta, constantValue(d->eta_, "eta")); setv(m->k, constantValue(d->k_, "k"));
        m->alpha_u = (float)constantValue(d->uRoughness_, "alpha").gray(); m->alpha_v = (float)constantValue(d->vRoughness_, "alpha").gray();
    } else if (isType(sm, "N5spica15RoughDielectricE")) {                   // bsdfs/roughdielectric.cc:41-83
        const auto* d = static_cast<const RoughDielectric*>(sm);
        noBump(d->bumpMap_);
        if (d->remapRoughness_) FatalError("roughness remapping is outside the GPU path's scope");
        m->type = SPB_MAT_ROUGHDIELECTRIC; m->distribution = distributionId(d->distribution_);
        setv(m->kr, constantValue(d->Kr_, "specularReflectance")); setv(m->kt, constantValue(d->Kt_, "specularTransmittance"));
        m->eta[0] = m->eta[1] = m->eta[2] = (float)constantValue(d->index_, "intIOR").gray();
        m->alpha_u = (float)constantValue(d->uRoughness_, "alpha").gray(); m->alpha_v = (float)constantValue(d->vRoughness_, "alpha").gray();
    } else if (isType(sm, "N5spica9ConductorE")) {                          // bsdfs/conductor.cc:28-38
        const auto* d = static_cast<const Conductor*>(sm);
        noBump(d->bumpMap_);
        m->type = SPB_MAT_CONDUCTOR; m->kr[0] = m->kr[1] = m->kr[2] = 1.f;
        setv(m->eta, constantValue(d->eta_, "eta")); setv(m->k, constantValue(d->k_, "k"));
    } else if (isType(sm, "N5spica7PlasticE")) {                            // bsdfs/plastic.cc:118-128
        const auto* d = static_cast<const Plastic*>(sm);
        noBump(d->bumpMap_);
        m->type = SPB_MAT_PLASTIC;
        tex[0] = reflectanceParam(f, d->Ks_, "specularReflectance", m->kr); tex[1] = reflectanceParam(f, d->Kd_, "diffuseReflectance", m->kt);
        m->eta[0] = m->eta[1] = m->eta[2] = (float)constantValue(d->eta_, "intIOR").gray();
    } else if (isType(sm, "N5spica12RoughPlasticE")) {                      // bsdfs/roughplastic.cc:130-165
        const auto* d = static_cast<const RoughPlastic*>(sm);
        noBump(d->bumpMap_);
        if (d->remapRoughness_) FatalError("roughness remapping is outside the GPU path's scope");
        m->type = SPB_MAT_ROUGHPLASTIC; m->distribution = distributionId(d->distribution_);
        tex[0] = reflectanceParam(f, d->Ks_, "specularReflectance", m->kr); tex[1] = reflectanceParam(f, d->Kd_, "diffuseReflectance", m->kt);
        m->eta[0] = m->eta[1] = m->eta[2] = (float)constantValue(d->index_, "intIOR").gray();
        m->alpha_u = m->alpha_v = (float)constantValue(d->roughness_, "alpha").gray();
    } else {
        FatalError("bsdf %s is outside the GPU path's scope (diffuse, dielectric, conductor, roughconductor, roughdielectric, plastic, roughplastic)", typeid(*sm).name());
    }
}

// Primitive i of the accelerator's list becomes triangle i (the reference keeps exactly one primitive
// per leaf, accelerators/bvh.cc:166-170), so hit ids are the reference's primitive indices.
inline void flatten(const std::vector<std::shared_ptr<Primitive>>& prims, FlatScene* f) {
    const size_t n = prims.size();
    f->verts.resize(n * 9); f->normals.assign(n * 9, 0.f); f->uvs.assign(n * 6, 0.f);
    f->material_id.assign(n, -1); f->light_id.assign(n, -1);
    std::map<const SurfaceMaterial*, int> matIndex;
    for (size_t i = 0; i < n; i++) {
        const auto* gp = static_cast<const GeometricPrimitive*>(prims[i].get());
        if (!isType(prims[i].get(), "N5spica18GeometricPrimitiveE") || !isType(gp->shape_.get(), "N5spica8TriangleE"))
            FatalError("primitive %zu is not a triangle (%s): analytic shapes are outside the GPU path's scope", i,
                       typeid(*prims[i]).name());
        if (gp->mediumInterface_) FatalError("participating media are outside the GPU path's scope");
        const auto& t = *static_cast<const Triangle*>(gp->shape_.get());
        bool ownNormals = false;
        for (int k = 0; k < 3; k++) {
            const Point3d& p = t[k];
            const Normal3d& nk = t.normal(k);
            const Point2d& uv = t.uv(k);
            f->verts[i * 9 + k * 3] = p.x(); f->verts[i * 9 + k * 3 + 1] = p.y(); f->verts[i * 9 + k * 3 + 2] = p.z();
            f->normals[i * 9 + k * 3] = (float)nk.x(); f->normals[i * 9 + k * 3 + 1] = (float)nk.y(); f->normals[i * 9 + k * 3 + 2] = (float)nk.z();
            f->uvs[i * 6 + k * 2] = (float)uv.x(); f->uvs[i * 6 + k * 2 + 1] = (float)uv.y();
            // a triangle built without vertex normals stores its face normal three times (core/triangle.cc:25-30)
            if (nk.x() != t.faceNormal_.x() || nk.y() != t.faceNormal_.y() || nk.z() != t.faceNormal_.z()) ownNormals = true;
            if (uv.x() != 0.0 || uv.y() != 0.0) f->anyUV = true;
        }
        if (ownNormals) f->anyNormals = true;
        const Material* mat = gp->material_.get();
        if (!mat || !mat->bsdf_) continue;                              // no bsdf: the path passes through (path.cc:71-75)
        if (mat->subsurface_) FatalError("subsurface materials are outside the GPU path's scope");
        const SurfaceMaterial* sm = mat->bsdf_.get();
        auto it = matIndex.find(sm);
        if (it == matIndex.end()) {
            spb_material m; int32_t tex[2];
            describeMaterial(f, sm, &m, tex);
            it = matIndex.emplace(sm, (int)f->materials.size()).first;
            f->materials.push_back(m);
            f->matTex.push_back(tex[0]); f->matTex.push_back(tex[1]);
        }
        f->material_id[i] = it->second;
    }
}

inline void uploadGeometry(spb_ctx* ctx, const FlatScene& f) {
    const int64_t n = (int64_t)f.material_id.size();
    check(ctx, spb_scene_set_triangles(ctx, f.verts.data(), f.anyNormals ? f.normals.data() : nullptr, f.anyUV ? f.uvs.data() : nullptr,
                                       f.material_id.data(), f.light_id.data(), n), "spb_scene_set_triangles");
    spb_build_opts opts; std::memset(&opts, 0, sizeof(opts));
    if (const char* b = getenv("SPICA_BVH_BUILDER")) opts.builder = atoi(b);    // SPB_BUILDER_*: 0 device binned SAH (default), 1 device LBVH, 2 host binned SAH
    if (const char* b = getenv("SPICA_BVH_MAX_LEAF")) opts.max_leaf_tris = atoi(b);   // 1..3 triangles per leaf (default 3)
    check(ctx, spb_bvh_build(ctx, &opts), "spb_bvh_build");
}

inline int deviceFromEnv() { const char* d = getenv("SPICA_DEVICE"); return d ? atoi(d) : 0; }
// SPICA_B200_DUMP_SCENE=<file>: write the resolved PODs to <file> and return without creating a device
// context or rendering -- lets the scene resolution be checked on a machine without a GPU (tests/test_refplugin.py).
inline const char* dumpPath() { const char* d = getenv("SPICA_B200_DUMP_SCENE"); return d && *d ? d : nullptr; }

// One GPU context with the scene's triangles and their BVH resident in HBM.
struct GpuScene {
    spb_ctx* ctx = nullptr;
    int device = 0;
    FlatScene flat;
    spb_bvh_stats stats;
    explicit GpuScene(const std::vector<std::shared_ptr<Primitive>>& prims) {
        device = deviceFromEnv();
        std::memset(&stats, 0, sizeof(stats));
        flatten(prims, &flat);
        if (dumpPath()) return;
        check(nullptr, spb_ctx_create(device, &ctx), "spb_ctx_create");
        uploadGeometry(ctx, flat);
        check(ctx, spb_bvh_get_stats(ctx, &stats), "spb_bvh_get_stats");
        MsgInfo("GPU BVH: %lld triangles, %lld wide nodes, built in %.3f s", (long long)stats.n_tris, (long long)stats.n_wide_nodes,
                stats.build_seconds);
    }
    ~GpuScene() { if (ctx) spb_ctx_destroy(ctx); }
    GpuScene(const GpuScene&) = delete;
    GpuScene& operator=(const GpuScene&) = delete;
};

// The `bvh` accelerator the reference host instantiates (core/cobject.h:66-75, core/accelerator.h:23-52).
// The class is defined in this header so that path.so recognises an instance made by bvh.so (by
// name, see isType) and renders from the same resident tree instead of building a second one.
class GpuBVHAccel : public Accelerator {
public:
    GpuBVHAccel(const std::vector<std::shared_ptr<Primitive>>& prims, RenderParams& params) : Accelerator{prims} {
        params.getBool("useSIMD", false, true);                         // accelerators/bvh.cc:127-131 (no meaning on the GPU)
        construct();
    }
    void construct() override {
        if (gpu_ || primitives_.empty()) return;
        gpu_ = std::make_unique<GpuScene>(primitives_);
        worldBound_ = Bounds3d(Point3d(gpu_->stats.world_lo[0], gpu_->stats.world_lo[1], gpu_->stats.world_lo[2]),
                               Point3d(gpu_->stats.world_hi[0], gpu_->stats.world_hi[1], gpu_->stats.world_hi[2]));
    }
    // Scalar queries from callers other than the GPU integrator (any reference integrator can sit on
    // top of this accelerator).  The GPU finds the primitive -- the search is the hot part -- and the
    // reference's own primitive then fills the SurfaceInteraction, exactly as the leaf visit of
    // accelerators/bvh.cc:341-349 does.  One ray per call is latency-bound (a launch per ray): it is
    // the compatibility path, the batched calls of include/spica_b200.h are the fast one.
    bool intersect(Ray& ray, SurfaceInteraction* isect) const override {
        if (!gpu_ || !gpu_->ctx) return false;
        spb_hit_f64 h;
        {
            std::lock_guard<std::mutex> lk(mu_);
            const spb_ray_f64 r = {ray.org().x(), ray.org().y(), ray.org().z(), ray.dir().x(), ray.dir().y(), ray.dir().z(), 0.0, ray.maxDist()};
            check(gpu_->ctx, spb_trace_closest_f64(gpu_->ctx, &r, 1, &h), "spb_trace_closest_f64");
        }
        if (h.prim < 0) return false;
        return primitives_[h.prim]->intersect(ray, isect);              // sets ray.maxDist = tHit (core/primitive.cc:49-62)
    }
    bool intersect(Ray& ray) const override {
        if (!gpu_ || !gpu_->ctx) return false;
        std::lock_guard<std::mutex> lk(mu_);
        const spb_ray_f64 r = {ray.org().x(), ray.org().y(), ray.org().z(), ray.dir().x(), ray.dir().y(), ray.dir().z(), 0.0, ray.maxDist()};
        uint8_t occ = 0;
        check(gpu_->ctx, spb_trace_any_f64(gpu_->ctx, &r, 1, &occ), "spb_trace_any_f64");
        return occ != 0;
    }
    std::vector<Triangle> triangulate() const override {               // accelerators/bvh.cc:237-246
        std::vector<Triangle> tris;
        for (const auto& p : primitives_) { auto t = p->triangulate(); tris.insert(tris.end(), t.begin(), t.end()); }
        return tris;
    }
    GpuScene* gpu() const { return gpu_.get(); }
private:
    std::unique_ptr<GpuScene> gpu_;
    mutable std::mutex mu_;
};
static const char* const kGpuBVHAccelName = "N5spica4b20011GpuBVHAccelE";

}  // namespace b200
}  // namespace spica

#endif  // SPICA_B200_REFPLUGIN_GPU_SCENE_H_
