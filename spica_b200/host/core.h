// C++17 host: the reference's plugin surface for the path-tracing hot path, re-implemented on top
// of the C ABI (include/spica_b200.h).  Same class roles, type names, XML parameter names, defaults
// and consume-on-read behaviour as the reference (paths below are relative to
// /root/reference/sources); nothing here computes radiance or intersections on the CPU -- the
// `bvh` accelerator and the `path` integrator forward to libspica_b200.so.
#pragma once
#include <algorithm>
#include <thread>
#include <type_traits>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/spica_b200.h"

namespace spica {

// n values of a plain type in storage that nobody touches before it is filled (std::vector::resize would zero-fill it on one thread)
template <class T>
struct RawArray {
    static_assert(std::is_trivially_copyable<T>::value, "plain values only");
    T* p = nullptr; size_t n = 0;
    RawArray() = default;
    RawArray(const RawArray&) = delete;
    RawArray& operator=(const RawArray&) = delete;
    ~RawArray() { std::free(p); }
    void resize(size_t count) { std::free(p); p = count ? (T*)std::malloc(sizeof(T) * count) : nullptr; n = p ? count : 0; if (count && !p) { fprintf(stderr, "out of memory\n"); abort(); } }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
};

// fn(begin, end) over [0, n) on the host cores (mesh loading and flattening of multi-million-triangle scenes)
template <class F>
inline void parallelFor(size_t n, F fn, size_t grain = 1 << 16) {
    size_t nt = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 32);
    nt = std::min(nt, (n + grain - 1) / grain);
    if (nt <= 1) { fn((size_t)0, n); return; }
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++) th.emplace_back([=]() { fn(n * t / nt, n * (t + 1) / nt); });
    for (auto& t : th) t.join();
}

// core/common.h:71-115: the reference reports errors by printing and aborting
[[noreturn]] void FatalError(const char* fmt, ...);
void Warning(const char* fmt, ...);
void MsgInfo(const char* fmt, ...);
#define SpicaAssert(cond, ...) do { if (!(cond)) ::spica::FatalError(__VA_ARGS__); } while (0)

constexpr double PI = 3.14159265358979323846;
constexpr double EPS = 1.0e-12;       // core/common.h:55
constexpr double INFTY = 1.0e32;      // core/common.h:54

struct Vector3d {
    double x = 0, y = 0, z = 0;
    Vector3d() = default;
    Vector3d(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    explicit Vector3d(const std::string& s);      // core/vector3d_detail.h:30-44: "a b c" | "a, b, c" | "a"
    Vector3d operator+(const Vector3d& o) const { return {x + o.x, y + o.y, z + o.z}; }
    Vector3d operator-(const Vector3d& o) const { return {x - o.x, y - o.y, z - o.z}; }
    Vector3d operator-() const { return {-x, -y, -z}; }
    Vector3d operator*(double s) const { return {x * s, y * s, z * s}; }
    double dot(const Vector3d& o) const { return x * o.x + y * o.y + z * o.z; }
    Vector3d cross(const Vector3d& o) const { return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x}; }
    double norm() const { return std::sqrt(dot(*this)); }
    Vector3d normalized() const { const double s = 1.0 / norm(); return {x * s, y * s, z * s}; }   // vector3d_detail.h:196-200
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
using Point3d = Vector3d;
using Normal3d = Vector3d;

struct Spectrum {           // RGBSpectrum (core/spectrum.h)
    double r = 0, g = 0, b = 0;
    Spectrum() = default;
    explicit Spectrum(double v) : r(v), g(v), b(v) {}
    Spectrum(double r_, double g_, double b_) : r(r_), g(g_), b(b_) {}
    double gray() const { return 0.2126 * r + 0.7152 * g + 0.0722 * b; }      // spectrum.cc:677
};

struct Matrix4x4 {
    double m[4][4];
    Matrix4x4();                                  // identity
    explicit Matrix4x4(const double v[4][4]);
    Matrix4x4 operator*(const Matrix4x4& o) const;
    Matrix4x4 transposed() const;
    Matrix4x4 inverted() const;                   // Gauss-Jordan, like core/matrix4x4.cc
};

class Transform {           // core/transform.h
public:
    Transform() = default;
    explicit Transform(const Matrix4x4& m) : m_(m), mInv_(m.inverted()) {}
    Transform(const Matrix4x4& m, const Matrix4x4& mInv) : m_(m), mInv_(mInv) {}
    Transform operator*(const Transform& t) const { return Transform(m_ * t.m_, t.mInv_ * mInv_); }
    Point3d applyPoint(const Point3d& p) const;   // transform.cc:60-76 (w-divide by w + EPS when w != 1)
    Vector3d applyVector(const Vector3d& v) const;
    Normal3d applyNormal(const Normal3d& n) const;
    Transform inverted() const { return Transform(mInv_, m_); }
    const Matrix4x4& getMat() const { return m_; }
    static Transform translate(const Vector3d& d);
    static Transform scale(double x, double y, double z);
    static Transform rotate(double theta, const Vector3d& axis);     // theta in RADIANS (transform.cc:136)
    static Transform lookAt(const Point3d& eye, const Point3d& look, const Vector3d& up);
    static Transform perspective(double fov, double aspect, double n, double f);
private:
    Matrix4x4 m_, mInv_;
};

// ---- core/cobject.h, core/renderparams.h ----------------------------------------------------------
class CObject {
public:
    virtual ~CObject() = default;
};

class RenderParams {        // typed, string-keyed bag; get*(name[, default][, remove])
public:
    static RenderParams& getInstance();
    void clear();
    void add(const std::string& n, bool v) { bools_[n] = v; }
    void add(const std::string& n, int v) { ints_[n] = v; }
    void add(const std::string& n, double v) { doubles_[n] = v; }
    void add(const std::string& n, const std::string& v) { strings_[n] = v; }
    void add(const std::string& n, const Spectrum& v) { spectrums_[n] = v; }
    void add(const std::string& n, const Vector3d& v) { vectors_[n] = v; }
    void add(const std::string& n, const Transform& v) { transforms_[n] = v; }
    void add(const std::string& n, const std::shared_ptr<CObject>& v) { objects_[n] = v; }

    bool getBool(const std::string& n, bool def, bool remove = false);
    int getInt(const std::string& n, bool remove = false);
    int getInt(const std::string& n, int def, bool remove = false);
    double getDouble(const std::string& n, bool remove = false);
    double getDouble(const std::string& n, double def, bool remove = false);
    std::string getString(const std::string& n, bool remove = false);
    std::string getString(const std::string& n, const std::string& def, bool remove = false);
    Spectrum getSpectrum(const std::string& n, bool remove = false);
    Point3d getPoint3d(const std::string& n, const Point3d& def, bool remove = false);
    Transform getTransform(const std::string& n, bool remove = false);
    Transform getTransform(const std::string& n, const Transform& def, bool remove = false);
    std::shared_ptr<CObject> getObject(const std::string& n, bool remove = false);
    std::shared_ptr<CObject> getObject(const std::string& n, std::nullptr_t, bool remove = false);
    // getTexture (renderparams.cc:374-440): object, else spectrum, else double -> a constant; here
    // only constants are in scope, returned as a Spectrum. `found` tells whether anything was there.
    Spectrum getTexture(const std::string& n, bool remove, bool* found);
    // the texture OBJECT stored under `n` (a nested <texture type="bitmap|checkerboard" name=n>), or nullptr
    std::shared_ptr<CObject> getTextureObject(const std::string& n, bool remove);
    Spectrum getTexture(const std::string& n, const Spectrum& def, bool remove = false);
    bool hasObject(const std::string& n) const { return objects_.count(n) != 0; }
private:
    std::map<std::string, bool> bools_;
    std::map<std::string, int> ints_;
    std::map<std::string, double> doubles_;
    std::map<std::string, std::string> strings_;
    std::map<std::string, Spectrum> spectrums_;
    std::map<std::string, Vector3d> vectors_;
    std::map<std::string, Transform> transforms_;
    std::map<std::string, std::shared_ptr<CObject>> objects_;
};

// ---- render interfaces (core/{filter,film,sampler,camera,material,light,shape,primitive,accelerator,scene,integrator}.h)
class Filter : public CObject {
public:
    virtual int kind() const = 0;                 // SPB_FILTER_*
    virtual double evaluate(double dx, double dy) const = 0;
    double rx = 1.0, ry = 1.0, sigma = 0.5;
};

struct Image {              // row-major RGB doubles
    int width = 0, height = 0;
    std::vector<double> rgb;
    Image() = default;
    Image(int w, int h) : width(w), height(h), rgb((size_t)w * h * 3, 0.0) {}
    double* pixel(int x, int y) { return &rgb[((size_t)y * width + x) * 3]; }
    const double* pixel(int x, int y) const { return &rgb[((size_t)y * width + x) * 3]; }
    void saveHdr(const std::string& file) const;  // core/image.cc:390-435 (RGBE, literal runs)
    void savePng(const std::string& file) const;  // 8-bit, values clamped to [0,1]
    static void writeHdrRgbe(const std::string& file, int width, int height, const unsigned char* rgbe);   // already-encoded pixels
    static void writePngRgb8(const std::string& file, int width, int height, const unsigned char* rgb);
    static Image loadHdr(const std::string& file);
    static Image fromFile(const std::string& file);
};

class Film : public CObject {       // core/film.h:23-86
public:
    Film(int w, int h, std::shared_ptr<Filter> filter, std::string filename);
    int width() const { return width_; }
    int height() const { return height_; }
    double aspect() const { return (double)width_ / height_; }
    const std::shared_ptr<Filter>& filter() const { return filter_; }
    void addPixel(int px, int py, double fx, double fy, const Spectrum& c);     // film.cc:65-74
    void setImage(const Image& img);                                            // film.cc:60-63
    void save(int id) const;                                                    // film.cc:23-40
    void setSaveCallback(std::function<void(const Image&)> cb) { callback_ = std::move(cb); }
    virtual void saveImage(const std::string& filename, const Image& img) const = 0;
    // Output stage on the device: fetch the ENCODED pixels of `ctx`'s film (spb_film_resolve_rgbe / _ldr) and write
    // the same file save() would.  false = this film cannot (or a save callback wants the float image).
    virtual bool saveFromDevice(spb_ctx* ctx, int id) const { (void)ctx; (void)id; return false; }
    std::string fileNameFor(int id) const;                                      // film.cc:31-33: sprintf(filename_, id)
    bool hasCallback() const { return (bool)callback_; }
protected:
    int width_, height_;
    std::shared_ptr<Filter> filter_;
    std::string filename_;
    Image image_;
    std::vector<double> weights_;
    std::function<void(const Image&)> callback_;
};

class Sampler : public CObject {    // core/sampler.h:24-36
public:
    virtual double get1D() = 0;
    virtual std::array<double, 2> get2D() { const double a = get1D(); return {a, get1D()}; }
    virtual void startPixel() {}
    virtual bool startNextSample() { return true; }
    virtual std::unique_ptr<Sampler> clone(unsigned int seed) const = 0;
};

class Camera : public CObject {     // core/camera.h
public:
    Transform cameraToWorld, cameraToScreen, rasterToCamera, screenToRaster, rasterToScreen;
    double lensRadius = 0.0, focalLength = 50.0;
    std::shared_ptr<Film> film;
};

class Texture : public CObject {    // core/texture.h:64-69, as far as the device needs it: the POD + (bitmap) its texels
public:
    virtual void describe(spb_texture* out, std::vector<float>* texels) const = 0;
};

class SurfaceMaterial : public CObject {
public:
    virtual void describe(spb_material* out) const = 0;   // the POD the device shades with
    // textures bound to the reflectance-type parameters (spb_material::kr / kt), nullptr = the constant
    virtual void textures(const Texture** kr, const Texture** kt) const { *kr = nullptr; *kt = nullptr; }
};

struct Triangle {                   // core/triangle.h: world-space points, optional normals / uvs
    Point3d p[3];
    Normal3d n[3];
    double uv[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    bool hasNormals = false;
};

class Light : public CObject {
public:
    virtual int kind() const = 0;   // SPB_LIGHT_*
};

struct Primitive {                  // GeometricPrimitive (core/primitive.h)
    Triangle tri;
    std::shared_ptr<SurfaceMaterial> material;
    std::shared_ptr<Light> light;
};

struct Ray {                        // core/ray.h:50-54 / ray.cc:11-19
    Point3d org; Vector3d dir; double maxDist = INFTY;
    Ray() = default;
    Ray(const Point3d& o, const Vector3d& d, double maxd = INFTY) : org(o), dir(d.normalized()), maxDist(maxd) {}
};
struct SurfaceInteraction { Point3d pos; double u = 0, v = 0; int primitive = -1; };

class Accelerator : public CObject {    // core/accelerator.h:23-52
public:
    explicit Accelerator(const std::vector<std::shared_ptr<Primitive>>& prims) : primitives_(prims) {}
    virtual bool intersect(Ray& ray, SurfaceInteraction* isect) const = 0;
    virtual bool intersect(Ray& ray) const = 0;
    virtual void worldBound(double lo[3], double hi[3]) const = 0;
    const std::vector<std::shared_ptr<Primitive>>& primitives() const { return primitives_; }
protected:
    std::vector<std::shared_ptr<Primitive>> primitives_;
};

class Scene {                           // core/scene.h
public:
    Scene(std::shared_ptr<Accelerator> a, std::vector<std::shared_ptr<Light>> l) : accel_(std::move(a)), lights_(std::move(l)) {}
    bool intersect(Ray& r, SurfaceInteraction* s) const { return accel_->intersect(r, s); }
    bool intersect(Ray& r) const { return accel_->intersect(r); }
    const std::vector<std::shared_ptr<Light>>& lights() const { return lights_; }
    const std::shared_ptr<Accelerator>& accelerator() const { return accel_; }
private:
    std::shared_ptr<Accelerator> accel_;
    std::vector<std::shared_ptr<Light>> lights_;
};

class Integrator : public CObject {     // core/integrator.h:29-37
public:
    virtual void render(const std::shared_ptr<const Camera>& camera, const Scene& scene, RenderParams& params) = 0;
};

// ---- plugin manager (core/cobject.h:30-75): type name -> factory; the reference dlopens
// plugins/<type>.so and calls its extern "C" createInstance; here the same factories are linked in
// and registered under the same names.
using CreateFunc = std::function<CObject*(RenderParams&)>;
using CreateAccelFunc = std::function<Accelerator*(const std::vector<std::shared_ptr<Primitive>>&, RenderParams&)>;
class PluginManager {
public:
    static PluginManager& getInstance();
    void registerPlugin(const std::string& type, CreateFunc f) { creators_[type] = std::move(f); }
    void registerAccelerator(const std::string& type, CreateAccelFunc f) { accels_[type] = std::move(f); }
    void initModule(const std::string& type);                      // FatalError when unknown, like a failed dlopen
    CObject* createObject(const std::string& type, RenderParams& params);
    Accelerator* createAccelerator(const std::string& type, const std::vector<std::shared_ptr<Primitive>>& prims, RenderParams& params);
private:
    std::map<std::string, CreateFunc> creators_;
    std::map<std::string, CreateAccelFunc> accels_;
};
void registerBuiltinPlugins();      // plugins.cc + gpu_path.cc

// ---- mesh loading (core/meshio.cc:76-261): one Triangle per face, transformed to world space
namespace meshio {
std::vector<Triangle> loadOBJ(const std::string& file, const Transform& objectToWorld);
// The triangles of a large mesh in storage that nobody touches before the loader's threads fill it (a std::vector would
// value-initialise 200 bytes per triangle on one thread first: seconds at ten million triangles).
struct TriangleBuffer {
    Triangle* p = nullptr; size_t n = 0;
    TriangleBuffer() = default;
    explicit TriangleBuffer(size_t count) : p((Triangle*)std::malloc(sizeof(Triangle) * std::max<size_t>(count, 1))), n(count) { if (!p) FatalError("out of memory for %zu triangles", count); }
    TriangleBuffer(TriangleBuffer&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    TriangleBuffer& operator=(TriangleBuffer&& o) noexcept { if (this != &o) { std::free(p); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    TriangleBuffer(const TriangleBuffer&) = delete;
    TriangleBuffer& operator=(const TriangleBuffer&) = delete;
    ~TriangleBuffer() { std::free(p); }
    static_assert(std::is_trivially_destructible<Triangle>::value && std::is_trivially_copyable<Triangle>::value, "Triangle must stay a plain value");
};
TriangleBuffer loadPLY(const std::string& file, const Transform& objectToWorld);
}

// ---- scene parser (spica/sceneparser.cc)
class SceneParser {
public:
    explicit SceneParser(const std::string& xmlFile);
    void parse();                       // loads, builds and renders, like SceneParser::parse
    void load();                        // the first half: XML -> objects (no accelerator, no device)
    void render();                      // the second half: integrator + accelerator + render
    const std::vector<std::shared_ptr<Primitive>>& primitives() const;
    const std::vector<std::shared_ptr<Light>>& lights() const;
    std::shared_ptr<Camera> camera() const;
private:
    struct Impl;
    std::shared_ptr<Impl> impl_;
};

// run-time options of this host that the reference has no flag for
struct HostOptions {
    int gpus = 1;                       // --gpus / SPICA_GPUS
    uint64_t seed = 0;                  // 0: seed from time(0) like the reference (core/integrator.cc:51)
    int sppOverride = 0;                // > 0: replaces the scene's sampleCount (bench / tests)
    bool savePasses = false;            // reference behaviour: save after every spp pass (integrator.cc:100)
    std::function<void(const Image&)> onImage;   // receives the final normalised image
};
HostOptions& hostOptions();
void hostGpuPrewarm(int gpus);           // start CUDA, one context per GPU and the NCCL communicator on background threads (call before parsing)
void hostGpuFinish();                    // join them (after the render)
void hostPhaseLap(const char* what);     // SPICA_TIMING=1: "[TIME] what  seconds since the first lap" on stdout

}  // namespace spica
