// The plugins the path-tracing scenes use, under the reference's type names and with its XML
// parameter names, defaults and consume-on-read flags (each constructor cites its counterpart).
// These classes only carry parameters: they describe themselves as PODs for the device.
#include "plugins.h"

#include <algorithm>
#include <cstring>

namespace spica {

// ---- filters (filters/box.cc:14-24, tent.cc:14-30, gaussian.cc:19-40) ------------------------------
class BoxFilter : public Filter {
public:
    explicit BoxFilter(RenderParams& p) { rx = p.getDouble("radius", 1.0); ry = p.getDouble("radius", 1.0, true); }
    int kind() const override { return SPB_FILTER_BOX; }
    double evaluate(double, double) const override { return 1.0; }
};
class TentFilter : public Filter {
public:
    explicit TentFilter(RenderParams& p) { rx = p.getDouble("radius", 1.0); ry = p.getDouble("radius", 1.0, true); }
    int kind() const override { return SPB_FILTER_TENT; }
    double evaluate(double dx, double dy) const override { return std::max(0.0, rx - std::abs(dx)) * std::max(0.0, ry - std::abs(dy)); }
};
class GaussianFilter : public Filter {
public:
    explicit GaussianFilter(RenderParams& p) { rx = p.getDouble("radius", 1.0); ry = p.getDouble("radius", 1.0); sigma = p.getDouble("sigma", 0.5, true); }
    int kind() const override { return SPB_FILTER_GAUSSIAN; }
    double evaluate(double dx, double dy) const override {
        const double beta = 1.0 / sigma;
        return std::max(0.0, std::exp(-beta * dx * dx) - std::exp(-rx * rx * beta)) * std::max(0.0, std::exp(-beta * dy * dy) - std::exp(-ry * ry * beta));
    }
};

// ---- films (core/film.cc, films/hdrfilm.cc:15-24, films/ldrfilm.cc:18-33) --------------------------
Film::Film(int w, int h, std::shared_ptr<Filter> filter, std::string filename)
    : width_(w), height_(h), filter_(std::move(filter)), filename_(std::move(filename)), image_(w, h), weights_((size_t)w * h, 0.0) {}
void Film::addPixel(int px, int py, double fx, double fy, const Spectrum& c) {
    const double w = filter_->evaluate(fx - 0.5, fy - 0.5);
    double* q = image_.pixel(px, py);
    q[0] += w * c.r; q[1] += w * c.g; q[2] += w * c.b;
    weights_[(size_t)py * width_ + px] += w;
}
void Film::setImage(const Image& img) {
    image_ = img;
    std::fill(weights_.begin(), weights_.end(), 1.0);
}
void Film::save(int id) const {
    Image res = image_;
    for (int y = 0; y < height_; y++) for (int x = 0; x < width_; x++) {
        const double inv = 1.0 / (weights_[(size_t)y * width_ + x] + EPS);
        double* q = res.pixel(x, y);
        q[0] *= inv; q[1] *= inv; q[2] *= inv;
    }
    saveImage(fileNameFor(id), res);
    if (callback_) callback_(res);
}
std::string Film::fileNameFor(int id) const {
    char savefile[1024];
    snprintf(savefile, sizeof(savefile), filename_.c_str(), id);
    return savefile;
}
class HDRFilm : public Film {
public:
    explicit HDRFilm(RenderParams& p)
        : Film(p.getInt("width", true), p.getInt("height", true), std::static_pointer_cast<Filter>(p.getObject("rfilter", true)), p.getString("outputFile")) {}
    void saveImage(const std::string& filename, const Image& img) const override {
        const std::string out = filename + ".hdr";
        img.saveHdr(out);
        MsgInfo("Save: %s", out.c_str());
    }
    bool saveFromDevice(spb_ctx* ctx, int id) const override {       // the same file, pixels encoded by the GPU
        if (hasCallback()) return false;
        std::vector<unsigned char> rgbe((size_t)width_ * height_ * 4);
        if (spb_film_resolve_rgbe(ctx, rgbe.data()) != SPB_OK) FatalError("spb_film_resolve_rgbe failed: %s", spb_last_error(ctx));
        const std::string out = fileNameFor(id) + ".hdr";
        Image::writeHdrRgbe(out, width_, height_, rgbe.data());
        MsgInfo("Save: %s", out.c_str());
        return true;
    }
};
class LDRFilm : public Film {
public:
    explicit LDRFilm(RenderParams& p)
        : Film(p.getInt("width", true), p.getInt("height", true), std::static_pointer_cast<Filter>(p.getObject("rfilter", true)), p.getString("outputFile", std::string("image"))),
          gamma_(p.getDouble("gamma", 2.2)) {}
    void saveImage(const std::string& filename, const Image& img) const override {
        SpicaAssert(gamma_ >= EPS, "Too small gamma is specified!!");          // core/tmo.cc:53-71
        Image t(img.width, img.height);
        const double ig = 1.0 / gamma_;
        for (size_t i = 0; i < img.rgb.size(); i++) t.rgb[i] = std::min(1.0, std::max(0.0, std::pow(img.rgb[i], ig)));
        const std::string out = filename + ".png";
        t.savePng(out);
        MsgInfo("Save: %s", out.c_str());
    }
    bool saveFromDevice(spb_ctx* ctx, int id) const override {
        if (hasCallback()) return false;
        std::vector<unsigned char> rgb((size_t)width_ * height_ * 3);
        if (spb_film_resolve_ldr(ctx, gamma_, rgb.data()) != SPB_OK) FatalError("spb_film_resolve_ldr failed: %s", spb_last_error(ctx));
        const std::string out = fileNameFor(id) + ".png";
        Image::writePngRgb8(out, width_, height_, rgb.data());
        MsgInfo("Save: %s", out.c_str());
        return true;
    }
private:
    double gamma_;
};

// ---- samplers -----------------------------------------------------------------------------------------
// The reference's three samplers all end up as one MT19937 stream per thread seeded from time(0)
// (SURVEY.md section 2); the GPU integrator replaces that with a counter-based stream keyed by
// (seed, pixel, sample, dimension). The host-side object keeps the Sampler interface alive for
// CPU-side callers and carries sampleCount.
double CounterSampler::get1D() {
    uint64_t z = (seed_ + 0x9e3779b97f4a7c15ull * ++counter_);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
static CObject* makeIndependent(RenderParams&) { return new CounterSampler(0); }                       // independent.cc:10-12
static CObject* makeLowDiscrepancy(RenderParams& p) { p.getInt("sampleCount", 32); p.getInt("dimensions", 2, true); return new CounterSampler(0); }   // ldsampler.cc:124-127
static CObject* makeHalton(RenderParams& p) { p.getInt("ns", true); p.getBool("isPermite", true, true); p.getInt("seed", 0, true); return new CounterSampler(0); }   // halton.cc:108-112

// ---- perspective camera (cameras/perspective.cc:24-51, core/camera.cc:20-38) -------------------------
class PerspectiveCamera : public Camera {
public:
    explicit PerspectiveCamera(RenderParams& p) {
        cameraToWorld = p.getTransform("toWorld", true);
        lensRadius = p.getDouble("apertureRadius", 0.0, true);
        focalLength = p.getDouble("focusDistance", 50.0, true);
        const double fov = p.getDouble("fov", true) * PI / 180.0;
        film = std::static_pointer_cast<Film>(p.getObject("film", true));
        cameraToScreen = Transform::perspective(fov, film->aspect(), 1.0e-2, 1000.0);
        // screen window = [-1,1]^2
        screenToRaster = Transform::scale(film->width(), film->height(), 1.0) * Transform::scale(1.0 / 2.0, -1.0 / 2.0, 1.0) *
                         Transform::translate(Vector3d(1.0, -1.0, 0.0));
        rasterToScreen = screenToRaster.inverted();
        rasterToCamera = cameraToScreen.inverted() * rasterToScreen;
    }
};

// ---- materials ------------------------------------------------------------------------------------------
static void setv(float* d, const Spectrum& s) { d[0] = (float)s.r; d[1] = (float)s.g; d[2] = (float)s.b; }
static int distributionId(const std::string& d) {
    if (d == "beckmann") return SPB_DISTR_BECKMANN;
    if (d == "ggx") return SPB_DISTR_GGX;
    FatalError("Unknown micforacet distribution type: %s", d.c_str());
}
static void rejectBump(RenderParams& p, bool remove) {
    bool found; p.getTexture("bumpMap", remove, &found);
    if (found) FatalError("bumpMap is outside this host's scope");
}

// ---- textures (textures/bitmap.cc:16-26, textures/checkerboard.cc:14-40) -------------------------------
class BitmapTexture : public Texture {
public:
    explicit BitmapTexture(RenderParams& p) : image_(Image::fromFile(p.getString("filename", true))) {}
    void describe(spb_texture* t, std::vector<float>* texels) const override {
        std::memset(t, 0, sizeof(*t));
        t->type = SPB_TEX_BITMAP; t->width = image_.width; t->height = image_.height;
        t->texel_offset = (int64_t)(texels->size() / 3);
        texels->insert(texels->end(), image_.rgb.begin(), image_.rgb.end());
    }
private:
    Image image_;
};
class Checkerboard : public Texture {
public:
    explicit Checkerboard(RenderParams& p)
        : c0_(p.getSpectrum("color0")), c1_(p.getSpectrum("color1")), uo_(p.getDouble("uoffset")), vo_(p.getDouble("voffset")),
          us_(p.getDouble("uscale")), vs_(p.getDouble("vscale")) {}
    void describe(spb_texture* t, std::vector<float>* texels) const override;
private:
    Spectrum c0_, c1_; double uo_, vo_, us_, vs_;
};

void Checkerboard::describe(spb_texture* t, std::vector<float>*) const {
    std::memset(t, 0, sizeof(*t));
    t->type = SPB_TEX_CHECKERBOARD; setv(t->color0, c0_); setv(t->color1, c1_);
    t->uoffset = (float)uo_; t->voffset = (float)vo_; t->uscale = (float)us_; t->vscale = (float)vs_;
}

// a reflectance-type parameter: a nested texture object, else the constant getTexture returns
struct TexParam { Spectrum value; std::shared_ptr<Texture> tex; };
static TexParam takeTexParam(RenderParams& p, const char* name, bool remove, bool* found) {
    TexParam t;
    if (auto obj = p.getTextureObject(name, remove)) {
        t.tex = std::dynamic_pointer_cast<Texture>(obj);
        if (!t.tex) FatalError("parameter \"%s\" is not a texture", name);
        *found = true;
        return t;
    }
    t.value = p.getTexture(name, remove, found);
    return t;
}

class Diffuse : public SurfaceMaterial {                  // bsdfs/diffuse.cc:18-32
public:
    explicit Diffuse(RenderParams& p) {
        bool found; kd_ = takeTexParam(p, "reflectance", true, &found);
        SpicaAssert(found, "Object not found: name = reflectance");
        rejectBump(p, true);
    }
    void describe(spb_material* m) const override { std::memset(m, 0, sizeof(*m)); m->type = SPB_MAT_DIFFUSE; setv(m->kr, kd_.value); }
    void textures(const Texture** kr, const Texture** kt) const override { *kr = kd_.tex.get(); *kt = nullptr; }
private:
    TexParam kd_;
};
class Dielectric : public SurfaceMaterial {               // bsdfs/dielectric.cc:23-42
public:
    explicit Dielectric(RenderParams& p) {
        bool f1, f2;
        kr_ = p.getTexture("specularReflectance", false, &f1);
        kt_ = p.getTexture("specularTransmittance", false, &f2);
        SpicaAssert(f1 && f2, "dielectric needs specularReflectance and specularTransmittance (the reference dereferences null otherwise, dielectric.cc:24-25)");
        ior_ = p.getTexture("intIOR", Spectrum(1.333)).gray();
        rejectBump(p, false);
    }
    void describe(spb_material* m) const override {
        std::memset(m, 0, sizeof(*m)); m->type = SPB_MAT_DIELECTRIC; setv(m->kr, kr_); setv(m->kt, kt_);
        m->eta[0] = m->eta[1] = m->eta[2] = (float)ior_;
    }
private:
    Spectrum kr_, kt_; double ior_;
};
class RoughConductor : public SurfaceMaterial {           // bsdfs/roughconductor.cc:29-75
public:
    explicit RoughConductor(RenderParams& p) {
        bool f1, f2;
        eta_ = p.getTexture("eta", true, &f1); k_ = p.getTexture("k", true, &f2);
        SpicaAssert(f1 && f2, "roughconductor needs eta and k");
        au_ = p.getTexture("alpha", Spectrum(0.1)).gray(); av_ = p.getTexture("alpha", Spectrum(0.1)).gray();
        distr_ = distributionId(p.getString("distribution", "beckmann", true));
        rejectBump(p, true);
    }
    void describe(spb_material* m) const override {
        std::memset(m, 0, sizeof(*m)); m->type = SPB_MAT_ROUGHCONDUCTOR; m->distribution = distr_;
        m->kr[0] = m->kr[1] = m->kr[2] = 1.f; setv(m->eta, eta_); setv(m->k, k_); m->alpha_u = (float)au_; m->alpha_v = (float)av_;
    }
private:
    Spectrum eta_, k_; double au_, av_; int distr_;
};
class Conductor : public SurfaceMaterial {                // bsdfs/conductor.cc:20-38
public:
    explicit Conductor(RenderParams& p) {
        bool f1, f2;
        eta_ = p.getTexture("eta", false, &f1); k_ = p.getTexture("k", false, &f2);
        SpicaAssert(f1 && f2, "conductor needs eta and k");
        rejectBump(p, false);
    }
    void describe(spb_material* m) const override {
        std::memset(m, 0, sizeof(*m)); m->type = SPB_MAT_CONDUCTOR; m->kr[0] = m->kr[1] = m->kr[2] = 1.f; setv(m->eta, eta_); setv(m->k, k_);
    }
private:
    Spectrum eta_, k_;
};
class RoughDielectric : public SurfaceMaterial {          // bsdfs/roughdielectric.cc:31-83
public:
    explicit RoughDielectric(RenderParams& p) {
        bool f1, f2;
        kr_ = p.getTexture("specularReflectance", false, &f1);
        kt_ = p.getTexture("specularTransmittance", false, &f2);
        SpicaAssert(f1 && f2, "roughdielectric needs specularReflectance and specularTransmittance");
        au_ = p.getTexture("alpha", Spectrum(0.1)).gray(); av_ = p.getTexture("alpha", Spectrum(0.1)).gray();
        ior_ = p.getTexture("intIOR", Spectrum(1.3333)).gray();
        distr_ = distributionId(p.getString("distribution", "beckmann", true));
        rejectBump(p, false);
    }
    void describe(spb_material* m) const override {
        std::memset(m, 0, sizeof(*m)); m->type = SPB_MAT_ROUGHDIELECTRIC; m->distribution = distr_;
        setv(m->kr, kr_); setv(m->kt, kt_); m->eta[0] = m->eta[1] = m->eta[2] = (float)ior_; m->alpha_u = (float)au_; m->alpha_v = (float)av_;
    }
private:
    Spectrum kr_, kt_; double au_, av_, ior_; int distr_;
};

class Plastic : public SurfaceMaterial {                  // bsdfs/plastic.cc:102-128
public:
    explicit Plastic(RenderParams& p) {
        bool f1, f2;
        kd_ = takeTexParam(p, "diffuseReflectance", false, &f1);
        ks_ = takeTexParam(p, "specularReflectance", false, &f2);
        SpicaAssert(f1 && f2, "plastic needs diffuseReflectance and specularReflectance (the reference dereferences null otherwise, plastic.cc:122-123)");
        ior_ = p.getTexture("intIOR", Spectrum(1.5)).gray();
        rejectBump(p, false);
    }
    void describe(spb_material* m) const override {
        std::memset(m, 0, sizeof(*m)); m->type = SPB_MAT_PLASTIC; setv(m->kr, ks_.value); setv(m->kt, kd_.value);
        m->eta[0] = m->eta[1] = m->eta[2] = (float)ior_;
    }
    void textures(const Texture** kr, const Texture** kt) const override { *kr = ks_.tex.get(); *kt = kd_.tex.get(); }
private:
    TexParam kd_, ks_; double ior_;
};
class RoughPlastic : public SurfaceMaterial {             // bsdfs/roughplastic.cc:118-165
public:
    explicit RoughPlastic(RenderParams& p) {
        bool f1, f2;
        kd_ = takeTexParam(p, "diffuseReflectance", false, &f1);
        ks_ = takeTexParam(p, "specularReflectance", false, &f2);
        SpicaAssert(f1 && f2, "roughplastic needs diffuseReflectance and specularReflectance");
        ior_ = p.getTexture("intIOR", Spectrum(1.5)).gray();
        alpha_ = p.getTexture("alpha", Spectrum(0.1)).gray();
        distr_ = distributionId(p.getString("distribution", "beckmann", true));
        rejectBump(p, false);
    }
    void describe(spb_material* m) const override {
        std::memset(m, 0, sizeof(*m)); m->type = SPB_MAT_ROUGHPLASTIC; m->distribution = distr_; setv(m->kr, ks_.value); setv(m->kt, kd_.value);
        m->eta[0] = m->eta[1] = m->eta[2] = (float)ior_; m->alpha_u = m->alpha_v = (float)alpha_;
    }
    void textures(const Texture** kr, const Texture** kt) const override { *kr = ks_.tex.get(); *kt = kd_.tex.get(); }
private:
    TexParam kd_, ks_; double ior_, alpha_; int distr_;
};

// ---- lights --------------------------------------------------------------------------------------------
static CObject* makeArea(RenderParams& p) {                // lights/area.cc:20-24
    p.getObject("shape", true);
    p.getTransform("toWorld", true);
    return new AreaLight(p.getSpectrum("radiance"));
}
static CObject* makeEnvmap(RenderParams& p) {              // lights/envmap.cc:49-55
    Envmap* e = new Envmap();
    e->worldCenter = p.getPoint3d("worldCenter", Point3d(0, 0, 0), true);
    e->worldRadius = p.getDouble("worldRadius", 2.0, true);
    e->image = Image::fromFile(p.getString("filename", true));
    e->toWorld = p.getTransform("toWorld", Transform(), true);
    e->scale = p.getDouble("scale", 1.0, true);
    return e;
}

void registerGpuPlugins();          // gpu_path.cc: `bvh` accelerator and `path` integrator

void registerBuiltinPlugins() {
    static bool done = false;
    if (done) return;
    done = true;
    PluginManager& pm = PluginManager::getInstance();
    pm.registerPlugin("box", [](RenderParams& p) -> CObject* { return new BoxFilter(p); });
    pm.registerPlugin("tent", [](RenderParams& p) -> CObject* { return new TentFilter(p); });
    pm.registerPlugin("gaussian", [](RenderParams& p) -> CObject* { return new GaussianFilter(p); });
    pm.registerPlugin("hdrfilm", [](RenderParams& p) -> CObject* { return new HDRFilm(p); });
    pm.registerPlugin("ldrfilm", [](RenderParams& p) -> CObject* { return new LDRFilm(p); });
    pm.registerPlugin("independent", makeIndependent);
    pm.registerPlugin("ldsampler", makeLowDiscrepancy);
    pm.registerPlugin("halton", makeHalton);
    pm.registerPlugin("perspective", [](RenderParams& p) -> CObject* { return new PerspectiveCamera(p); });
    pm.registerPlugin("diffuse", [](RenderParams& p) -> CObject* { return new Diffuse(p); });
    pm.registerPlugin("dielectric", [](RenderParams& p) -> CObject* { return new Dielectric(p); });
    pm.registerPlugin("roughconductor", [](RenderParams& p) -> CObject* { return new RoughConductor(p); });
    pm.registerPlugin("conductor", [](RenderParams& p) -> CObject* { return new Conductor(p); });
    pm.registerPlugin("roughdielectric", [](RenderParams& p) -> CObject* { return new RoughDielectric(p); });
    pm.registerPlugin("bitmap", [](RenderParams& p) -> CObject* { return new BitmapTexture(p); });
    pm.registerPlugin("checkerboard", [](RenderParams& p) -> CObject* { return new Checkerboard(p); });
    pm.registerPlugin("plastic", [](RenderParams& p) -> CObject* { return new Plastic(p); });
    pm.registerPlugin("roughplastic", [](RenderParams& p) -> CObject* { return new RoughPlastic(p); });
    pm.registerPlugin("area", makeArea);
    pm.registerPlugin("envmap", makeEnvmap);
    registerGpuPlugins();
}

}  // namespace spica
