// Math, RenderParams and PluginManager of the host (see core.h).
#include "core.h"

#include <cstdarg>
#include <cstring>

namespace spica {

void FatalError(const char* fmt, ...) {          // core/common.h:109-115
    va_list ap; va_start(ap, fmt);
    fprintf(stderr, "[ERROR] "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n");
    va_end(ap);
    std::abort();
}
void Warning(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    fprintf(stderr, "[WARNING] "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n");
    va_end(ap);
}
void MsgInfo(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    printf("[INFO] "); vprintf(fmt, ap); printf("\n");
    va_end(ap);
}

Vector3d::Vector3d(const std::string& s) {
    double a, b, c;
    if (sscanf(s.c_str(), "%lf %lf %lf", &a, &b, &c) == 3 || sscanf(s.c_str(), "%lf, %lf, %lf", &a, &b, &c) == 3) { x = a; y = b; z = c; }
    else if (sscanf(s.c_str(), "%lf", &a) == 1) { x = y = z = a; }
    else FatalError("Cannot parse string \"%s\" for Vector3d", s.c_str());
}

Matrix4x4::Matrix4x4() { std::memset(m, 0, sizeof(m)); for (int i = 0; i < 4; i++) m[i][i] = 1.0; }
Matrix4x4::Matrix4x4(const double v[4][4]) { std::memcpy(m, v, sizeof(m)); }
Matrix4x4 Matrix4x4::operator*(const Matrix4x4& o) const {
    Matrix4x4 r;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
        double s = 0.0;
        for (int k = 0; k < 4; k++) s += m[i][k] * o.m[k][j];
        r.m[i][j] = s;
    }
    return r;
}
Matrix4x4 Matrix4x4::transposed() const {
    Matrix4x4 r;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = m[j][i];
    return r;
}
Matrix4x4 Matrix4x4::inverted() const {
    double a[4][8];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { a[i][j] = m[i][j]; a[i][j + 4] = i == j ? 1.0 : 0.0; }
    for (int c = 0; c < 4; c++) {
        int piv = c;
        for (int r = c + 1; r < 4; r++) if (std::abs(a[r][c]) > std::abs(a[piv][c])) piv = r;
        SpicaAssert(std::abs(a[piv][c]) > 0.0, "Matrix is singular!");
        if (piv != c) for (int j = 0; j < 8; j++) std::swap(a[piv][j], a[c][j]);
        const double inv = 1.0 / a[c][c];
        for (int j = 0; j < 8; j++) a[c][j] *= inv;
        for (int r = 0; r < 4; r++) if (r != c) {
            const double f = a[r][c];
            if (f != 0.0) for (int j = 0; j < 8; j++) a[r][j] -= f * a[c][j];
        }
    }
    Matrix4x4 r;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = a[i][j + 4];
    return r;
}

Point3d Transform::applyPoint(const Point3d& p) const {
    const double ps[4] = {p.x, p.y, p.z, 1.0};
    double r[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r[i] += m_.m[i][j] * ps[j];
    if (r[3] != 1.0) { r[0] /= (r[3] + EPS); r[1] /= (r[3] + EPS); r[2] /= (r[3] + EPS); }
    return {r[0], r[1], r[2]};
}
Vector3d Transform::applyVector(const Vector3d& v) const {
    return {m_.m[0][0] * v.x + m_.m[0][1] * v.y + m_.m[0][2] * v.z, m_.m[1][0] * v.x + m_.m[1][1] * v.y + m_.m[1][2] * v.z,
            m_.m[2][0] * v.x + m_.m[2][1] * v.y + m_.m[2][2] * v.z};
}
Normal3d Transform::applyNormal(const Normal3d& n) const {          // transform.cc:78-83
    return {mInv_.m[0][0] * n.x + mInv_.m[1][0] * n.y + mInv_.m[2][0] * n.z, mInv_.m[0][1] * n.x + mInv_.m[1][1] * n.y + mInv_.m[2][1] * n.z,
            mInv_.m[0][2] * n.x + mInv_.m[1][2] * n.y + mInv_.m[2][2] * n.z};
}
Transform Transform::translate(const Vector3d& d) {
    Matrix4x4 a, b;
    a.m[0][3] = d.x; a.m[1][3] = d.y; a.m[2][3] = d.z;
    b.m[0][3] = -d.x; b.m[1][3] = -d.y; b.m[2][3] = -d.z;
    return Transform(a, b);
}
Transform Transform::scale(double x, double y, double z) {
    SpicaAssert(x != 0.0 && y != 0.0 && z != 0.0, "Zero division!!");
    Matrix4x4 a, b;
    a.m[0][0] = x; a.m[1][1] = y; a.m[2][2] = z;
    b.m[0][0] = 1.0 / x; b.m[1][1] = 1.0 / y; b.m[2][2] = 1.0 / z;
    return Transform(a, b);
}
Transform Transform::rotate(double theta, const Vector3d& axis) {   // transform.cc:136-165 (Rodrigues)
    const Vector3d a = axis.normalized();
    const double s = std::sin(theta), c = std::cos(theta);
    Matrix4x4 r;
    r.m[0][0] = a.x * a.x + (1.0 - a.x * a.x) * c; r.m[0][1] = a.x * a.y * (1.0 - c) - a.z * s; r.m[0][2] = a.x * a.z * (1.0 - c) + a.y * s;
    r.m[1][0] = a.x * a.y * (1.0 - c) + a.z * s; r.m[1][1] = a.y * a.y + (1.0 - a.y * a.y) * c; r.m[1][2] = a.y * a.z * (1.0 - c) - a.x * s;
    r.m[2][0] = a.x * a.z * (1.0 - c) - a.y * s; r.m[2][1] = a.y * a.z * (1.0 - c) + a.x * s; r.m[2][2] = a.z * a.z + (1.0 - a.z * a.z) * c;
    return Transform(r, r.transposed());
}
Transform Transform::lookAt(const Point3d& eye, const Point3d& look, const Vector3d& up) {   // transform.cc:167-197
    const Vector3d dir = (look - eye).normalized();
    Vector3d left = up.normalized().cross(dir);
    SpicaAssert(left.norm() != 0.0, "Up vector and viewing direction are oriented the same direction!!");
    left = left.normalized();
    const Vector3d nu = dir.cross(left);
    Matrix4x4 c;
    c.m[0][0] = left.x; c.m[1][0] = left.y; c.m[2][0] = left.z;
    c.m[0][1] = nu.x; c.m[1][1] = nu.y; c.m[2][1] = nu.z;
    c.m[0][2] = dir.x; c.m[1][2] = dir.y; c.m[2][2] = dir.z;
    c.m[0][3] = eye.x; c.m[1][3] = eye.y; c.m[2][3] = eye.z;
    return Transform(c);
}
Transform Transform::perspective(double fov, double aspect, double n, double f) {            // transform.cc:204-212
    Matrix4x4 p;
    std::memset(p.m, 0, sizeof(p.m));
    p.m[0][0] = 1.0 / aspect; p.m[1][1] = 1.0; p.m[2][2] = f / (f - n); p.m[2][3] = -f * n / (f - n); p.m[3][2] = 1.0;
    const double it = 1.0 / std::tan(fov / 2.0);
    return scale(it, it, 1.0) * Transform(p);
}

// ---- RenderParams -----------------------------------------------------------------------------------
RenderParams& RenderParams::getInstance() { static RenderParams p; return p; }
void RenderParams::clear() { *this = RenderParams(); }

namespace {
template <class M>
typename M::mapped_type take(M& m, const std::string& n, bool remove, const char* what) {
    auto it = m.find(n);
    SpicaAssert(it != m.end(), "%s not found: name = %s", what, n.c_str());
    auto v = it->second;
    if (remove) m.erase(it);
    return v;
}
template <class M>
typename M::mapped_type takeOr(M& m, const std::string& n, const typename M::mapped_type& def, bool remove) {
    auto it = m.find(n);
    if (it == m.end()) return def;
    auto v = it->second;
    if (remove) m.erase(it);
    return v;
}
}  // namespace

bool RenderParams::getBool(const std::string& n, bool def, bool remove) { return takeOr(bools_, n, def, remove); }
int RenderParams::getInt(const std::string& n, bool remove) { return take(ints_, n, remove, "Int"); }
int RenderParams::getInt(const std::string& n, int def, bool remove) { return takeOr(ints_, n, def, remove); }
double RenderParams::getDouble(const std::string& n, bool remove) { return take(doubles_, n, remove, "Double"); }
double RenderParams::getDouble(const std::string& n, double def, bool remove) { return takeOr(doubles_, n, def, remove); }
std::string RenderParams::getString(const std::string& n, bool remove) { return take(strings_, n, remove, "String"); }
std::string RenderParams::getString(const std::string& n, const std::string& def, bool remove) { return takeOr(strings_, n, def, remove); }
Spectrum RenderParams::getSpectrum(const std::string& n, bool remove) { return take(spectrums_, n, remove, "Spectrum"); }
Point3d RenderParams::getPoint3d(const std::string& n, const Point3d& def, bool remove) { return takeOr(vectors_, n, def, remove); }
Transform RenderParams::getTransform(const std::string& n, bool remove) { return take(transforms_, n, remove, "Transform"); }
Transform RenderParams::getTransform(const std::string& n, const Transform& def, bool remove) { return takeOr(transforms_, n, def, remove); }
std::shared_ptr<CObject> RenderParams::getObject(const std::string& n, bool remove) { return take(objects_, n, remove, "Object"); }
std::shared_ptr<CObject> RenderParams::getObject(const std::string& n, std::nullptr_t, bool remove) {
    return takeOr(objects_, n, std::shared_ptr<CObject>(), remove);
}
std::shared_ptr<CObject> RenderParams::getTextureObject(const std::string& n, bool remove) {
    auto it = objects_.find(n);
    if (it == objects_.end()) return nullptr;
    auto v = it->second;
    if (remove) objects_.erase(it);
    return v;
}
Spectrum RenderParams::getTexture(const std::string& n, bool remove, bool* found) {
    *found = true;
    if (objects_.count(n)) FatalError("a texture object on parameter \"%s\" is outside this host's scope (textures bind to reflectance-type parameters only)", n.c_str());
    auto s = spectrums_.find(n);
    if (s != spectrums_.end()) { Spectrum v = s->second; if (remove) spectrums_.erase(s); return v; }
    auto d = doubles_.find(n);
    if (d != doubles_.end()) { const double v = d->second; if (remove) doubles_.erase(d); return Spectrum(v); }
    *found = false;
    return Spectrum(0.0);
}
Spectrum RenderParams::getTexture(const std::string& n, const Spectrum& def, bool remove) {
    bool found;
    const Spectrum v = getTexture(n, remove, &found);
    return found ? v : def;
}

// ---- PluginManager ----------------------------------------------------------------------------------
PluginManager& PluginManager::getInstance() { static PluginManager p; return p; }
void PluginManager::initModule(const std::string& type) {
    if (!creators_.count(type)) FatalError("Failed to load plugin: plugins/%s.so (this host implements the path-tracing hot path only)", type.c_str());
}
CObject* PluginManager::createObject(const std::string& type, RenderParams& params) {
    initModule(type);
    return creators_[type](params);
}
Accelerator* PluginManager::createAccelerator(const std::string& type, const std::vector<std::shared_ptr<Primitive>>& prims, RenderParams& params) {
    if (!accels_.count(type)) FatalError("Failed to load accelerator plugin: plugins/%s.so", type.c_str());
    return accels_[type](prims, params);
}

HostOptions& hostOptions() { static HostOptions o; return o; }

}  // namespace spica
