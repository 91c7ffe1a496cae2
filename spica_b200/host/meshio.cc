// Mesh loaders with the reference's behaviour (core/meshio.cc:76-261): one world-space Triangle per
// face; OBJ through a small reader that reproduces what the reference takes from tinyobjloader
// (float32 positions widened to double, fan triangulation, per-shape "has normals / has uvs").
#include <atomic>
#include <cstring>
#include <new>
#include <fstream>
#include <sstream>

#include "core.h"

namespace spica {
namespace meshio {
namespace {

Triangle makeTriangle(const Point3d p[3], const Normal3d n[3], const double uv[3][2], bool hasN, bool hasUV, const Transform& o2w) {
    Triangle t;
    for (int k = 0; k < 3; k++) t.p[k] = o2w.applyPoint(p[k]);                 // core/triangle.cc:20-22
    if (hasN) {
        // core/triangle.cc:40-42: normals go through Transform::apply(Vector3d), then normalized()
        for (int k = 0; k < 3; k++) t.n[k] = o2w.applyVector(n[k]).normalized();
        t.hasNormals = true;
        if (hasUV) for (int k = 0; k < 3; k++) { t.uv[k][0] = uv[k][0]; t.uv[k][1] = uv[k][1]; }   // triangle.cc:64
    }
    return t;
}

struct Idx { int v = -1, vt = -1, vn = -1; };

bool parseIndex(const char*& s, Idx* out, int nv, int nvt, int nvn) {
    auto fix = [](int i, int n) { return i > 0 ? i - 1 : (i < 0 ? n + i : -1); };
    char* e;
    long v = strtol(s, &e, 10);
    if (e == s) return false;
    out->v = fix((int)v, nv);
    s = e;
    if (*s == '/') {
        s++;
        if (*s != '/') { long t = strtol(s, &e, 10); if (e != s) out->vt = fix((int)t, nvt); s = e; }
        if (*s == '/') { s++; long n = strtol(s, &e, 10); if (e != s) out->vn = fix((int)n, nvn); s = e; }
    }
    return true;
}

}  // namespace

std::vector<Triangle> loadOBJ(const std::string& file, const Transform& o2w) {
    std::ifstream ifs(file);
    if (!ifs) FatalError("Failed to open OBJ file \"%s\" !!", file.c_str());
    std::vector<float> V, VN, VT;
    struct Shape { std::vector<Idx> idx; };
    std::vector<Shape> shapes(1);
    std::string line;
    while (std::getline(ifs, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        const char* s = line.c_str();
        while (*s == ' ' || *s == '\t') s++;
        if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
            double a = 0, b = 0, c = 0; sscanf(s + 2, "%lf %lf %lf", &a, &b, &c);
            V.push_back((float)a); V.push_back((float)b); V.push_back((float)c);
        } else if (s[0] == 'v' && s[1] == 'n') {
            double a = 0, b = 0, c = 0; sscanf(s + 3, "%lf %lf %lf", &a, &b, &c);
            VN.push_back((float)a); VN.push_back((float)b); VN.push_back((float)c);
        } else if (s[0] == 'v' && s[1] == 't') {
            double a = 0, b = 0; sscanf(s + 3, "%lf %lf", &a, &b);
            VT.push_back((float)a); VT.push_back((float)b);
        } else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
            std::vector<Idx> poly;
            const char* p = s + 2;
            for (;;) {
                while (*p == ' ' || *p == '\t') p++;
                if (!*p) break;
                Idx ix;
                if (!parseIndex(p, &ix, (int)V.size() / 3, (int)VT.size() / 2, (int)VN.size() / 3)) break;
                poly.push_back(ix);
            }
            for (size_t k = 2; k < poly.size(); k++) {      // fan, like tinyobjloader's triangulation
                shapes.back().idx.push_back(poly[0]); shapes.back().idx.push_back(poly[k - 1]); shapes.back().idx.push_back(poly[k]);
            }
        } else if ((s[0] == 'g' || s[0] == 'o') && (s[1] == ' ' || s[1] == '\t' || s[1] == 0)) {
            if (!shapes.back().idx.empty()) shapes.emplace_back();
        }
    }
    std::vector<Triangle> out;
    for (const Shape& sh : shapes) {
        bool hasN = true, hasUV = true;                       // core/meshio.cc:193-221: per shape
        for (const Idx& ix : sh.idx) { if (ix.vn < 0) hasN = false; if (ix.vt < 0) hasUV = false; }
        for (size_t i = 0; i + 2 < sh.idx.size(); i += 3) {
            Point3d p[3]; Normal3d n[3]; double uv[3][2] = {{0, 0}, {0, 0}, {0, 0}};
            for (int k = 0; k < 3; k++) {
                const Idx& ix = sh.idx[i + k];
                if (ix.v >= 0 && (size_t)ix.v * 3 + 2 < V.size()) p[k] = Point3d(V[ix.v * 3], V[ix.v * 3 + 1], V[ix.v * 3 + 2]);
                if (ix.vn >= 0 && (size_t)ix.vn * 3 + 2 < VN.size()) n[k] = Normal3d(VN[ix.vn * 3], VN[ix.vn * 3 + 1], VN[ix.vn * 3 + 2]);
                if (ix.vt >= 0 && (size_t)ix.vt * 2 + 1 < VT.size()) { uv[k][0] = VT[ix.vt * 2]; uv[k][1] = VT[ix.vt * 2 + 1]; }
            }
            out.push_back(makeTriangle(p, n, uv, hasN, hasN && hasUV, o2w));
        }
    }
    return out;
}

TriangleBuffer loadPLY(const std::string& file, const Transform& o2w) {
    std::ifstream ifs(file, std::ios::in | std::ios::binary);
    if (!ifs.is_open()) FatalError("failed to open file \"%s\" !!", file.c_str());
    std::string line;
    std::getline(ifs, line);
    if (!line.empty() && line.back() == '\r') line.pop_back();
    SpicaAssert(line == "ply", "Invalid format identifier");
    long nv = 0, nf = 0;
    for (;;) {
        if (!std::getline(ifs, line)) FatalError("PLY header has no end_header: %s", file.c_str());
        std::stringstream ss(line);
        std::string key, name, val;
        ss >> key;
        if (key == "format") { ss >> name; SpicaAssert(name == "binary_little_endian", "PLY must be binary little endian format!"); }
        else if (key == "element") {
            ss >> name;
            if (name == "vertex") ss >> nv; else if (name == "face") ss >> nf; else FatalError("Invalid element indentifier");
        } else if (key == "end_header") break;
    }
    SpicaAssert(nv > 0 && nf > 0, "numVerts and numFaces must be positive");
    std::vector<float> v((size_t)nv * 3);
    ifs.read((char*)v.data(), sizeof(float) * v.size());          // the reference reads xyz only (meshio.cc:137-140)
    if (!ifs) FatalError("PLY file is truncated: %s", file.c_str());
    // the face block in one read (13 bytes per triangle; the reference pulls it through the stream field by field, meshio.cc:143-160)
    const std::streampos at = ifs.tellg();
    ifs.seekg(0, std::ios::end);
    const size_t rest = (size_t)(ifs.tellg() - at);
    ifs.seekg(at);
    std::vector<unsigned char> fb(rest);
    ifs.read((char*)fb.data(), (std::streamsize)rest);
    if (!ifs) FatalError("PLY file is truncated: %s", file.c_str());
    std::vector<size_t> offset;                                  // only when some face is not a triangle
    bool allTriangles = rest >= (size_t)nf * 13;
    if (allTriangles) for (long i = 0; i < nf && allTriangles; i++) allTriangles = fb[(size_t)i * 13] == 3;
    if (!allTriangles) {                                         // polygons: the first three indices of each, like the reference
        offset.resize((size_t)nf);
        size_t o = 0;
        for (long i = 0; i < nf; i++) {
            if (o + 13 > rest) FatalError("PLY file is truncated: %s", file.c_str());
            const unsigned char vs = fb[o];
            offset[(size_t)i] = o;
            if (vs > 3) Warning("mesh contains non-triangle polygon (%d vertices) !!", (int)vs);
            o += 1 + sizeof(int) * (size_t)std::max<int>(vs, 3);
        }
    }
    TriangleBuffer out((size_t)nf);
    const double uv[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    std::atomic<bool> bad(false);
    parallelFor((size_t)nf, [&](size_t b, size_t e) {
        Normal3d n[3];
        for (size_t i = b; i < e; i++) {
            int ii[3];
            std::memcpy(ii, fb.data() + (allTriangles ? i * 13 : offset[i]) + 1, sizeof(int) * 3);
            Point3d p[3];
            for (int k = 0; k < 3; k++) {
                if (ii[k] < 0 || ii[k] >= nv) { bad = true; ii[k] = 0; }
                p[k] = Point3d(v[(size_t)ii[k] * 3], v[(size_t)ii[k] * 3 + 1], v[(size_t)ii[k] * 3 + 2]);
            }
            new (out.p + i) Triangle(makeTriangle(p, n, uv, false, false, o2w));
        }
    });
    SpicaAssert(!bad, "PLY vertex index out of range");
    return out;
}

}  // namespace meshio
}  // namespace spica
