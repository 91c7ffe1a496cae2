// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product library.
//
// Harness AROUND the unmodified reference (tatsy/spica): loads a mesh with the reference's own
// loaders, builds the reference's own BVHAccel and drives BVHAccel::intersect over a ray file.
// It is the "real reference" ray-cast oracle and, under bench.py, the
// cpu_baseline of kind "reference".
//
// The reference sources are compiled from where they lie (the Makefile passes REF_BVH_CC);
// bvh.cc is pulled in as a single TU because its header exports the plugin symbols
// (accelerators/bvh.h:97), and `private` is opened only around that include so the harness can
// dump root_/nodes_ (accelerators/bvh.h:90-91) for the import path.
//
// Ray file (little endian, no header):
//   f32 : n x 8 float32  { ox, oy, oz, dx, dy, dz, tmin(ignored), tmax }
//   f64 : n x 8 float64  { ox, oy, oz, dx, dy, dz, tmin(ignored), tmax }
// Every ray goes through the reference's Ray constructor (core/ray.cc:11-19), i.e. the direction
// is re-normalised by reciprocal-multiply and invdir derived from it.
// Output (closest): int32 prim[n] (-1 = miss) followed by float64 t[n]
// Output (any)    : uint8 occluded[n]
// BVH dump        : int32 n_nodes, int32 root; then n_nodes x { f64 lo[3], hi[3]; i32 left,right,prim,axis }

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stack>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "core/accelerator.h"
#include "core/bounds3d.h"
#include "core/interaction.h"
#include "core/meshio.h"
#include "core/primitive.h"
#include "core/ray.h"
#include "core/renderparams.h"
#include "core/transform.h"
#include "core/triangle.h"

#define private public
#include REF_BVH_CC
#undef private

using namespace spica;

namespace {

double now() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

template <class T>
std::vector<T> readAll(const std::string& path) {
    std::ifstream ifs(path, std::ios::binary | std::ios::ate);
    if (!ifs) { fprintf(stderr, "cannot open %s\n", path.c_str()); std::exit(2); }
    size_t bytes = (size_t)ifs.tellg();
    ifs.seekg(0);
    std::vector<T> v(bytes / sizeof(T));
    ifs.read((char*)v.data(), v.size() * sizeof(T));
    return v;
}

struct DumpNode {
    double lo[3], hi[3];
    int32_t left, right, prim, axis;
};

}  // namespace

int main(int argc, char** argv) {
    std::string ply, tris, raysFile, rayFormat = "f32", mode = "closest", out, dumpBvh;
    int threads = (int)std::thread::hardware_concurrency();
    int repeat = 1, simd = 0, warmup = 0;
    long stride = 1, limit = -1;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2);} return argv[++i]; };
        if (a == "--ply") ply = next();
        else if (a == "--tris") tris = next();
        else if (a == "--rays") raysFile = next();
        else if (a == "--ray-format") rayFormat = next();
        else if (a == "--mode") mode = next();
        else if (a == "--out") out = next();
        else if (a == "--dump-bvh") dumpBvh = next();
        else if (a == "--threads") threads = std::atoi(next().c_str());
        else if (a == "--repeat") repeat = std::atoi(next().c_str());
        else if (a == "--warmup") warmup = std::atoi(next().c_str());
        else if (a == "--simd") simd = std::atoi(next().c_str());
        else if (a == "--stride") stride = std::atol(next().c_str());
        else if (a == "--limit") limit = std::atol(next().c_str());
        else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 2; }
    }
    if (threads < 1) threads = 1;

    // ---- geometry, through the reference's own types ------------------------------------
    std::vector<std::shared_ptr<Primitive>> prims;
    if (!ply.empty()) {
        auto groups = meshio::loadPLY(ply, Transform());
        for (const auto& g : groups)
            for (const auto& s : g.shapes())
                prims.push_back(std::make_shared<GeometricPrimitive>(s, nullptr, nullptr, nullptr));
    } else if (!tris.empty()) {
        auto v = readAll<double>(tris);
        size_t n = v.size() / 9;
        for (size_t i = 0; i < n; i++) {
            const double* p = &v[i * 9];
            std::shared_ptr<Shape> s(new Triangle(Point3d(p[0], p[1], p[2]), Point3d(p[3], p[4], p[5]),
                                                  Point3d(p[6], p[7], p[8]), Transform()));
            prims.push_back(std::make_shared<GeometricPrimitive>(s, nullptr, nullptr, nullptr));
        }
    } else {
        fprintf(stderr, "need --ply or --tris\n");
        return 2;
    }
    std::unordered_map<const Primitive*, int> primIndex;
    primIndex.reserve(prims.size() * 2);
    for (size_t i = 0; i < prims.size(); i++) primIndex[prims[i].get()] = (int)i;

    double t0 = now();
    BVHAccel accel(prims, simd != 0);
    double buildSec = now() - t0;

    if (!dumpBvh.empty()) {
        std::unordered_map<const BVHNode*, int> nodeIndex;
        nodeIndex.reserve(accel.nodes_.size() * 2);
        for (size_t i = 0; i < accel.nodes_.size(); i++) nodeIndex[accel.nodes_[i].get()] = (int)i;
        std::ofstream ofs(dumpBvh, std::ios::binary);
        int32_t hdr[2] = { (int32_t)accel.nodes_.size(), accel.root_ ? nodeIndex[accel.root_] : -1 };
        ofs.write((const char*)hdr, sizeof(hdr));
        for (const auto& up : accel.nodes_) {
            const BVHNode* nd = up.get();
            DumpNode d;
            for (int k = 0; k < 3; k++) { d.lo[k] = nd->bounds.posMin()[k]; d.hi[k] = nd->bounds.posMax()[k]; }
            bool leaf = nd->isLeaf();
            d.left  = (!leaf && nd->left)  ? nodeIndex[nd->left]  : -1;
            d.right = (!leaf && nd->right) ? nodeIndex[nd->right] : -1;
            d.prim = nd->primIdx;
            d.axis = nd->splitAxis;
            ofs.write((const char*)&d, sizeof(d));
        }
    }

    long nRays = 0, nUsed = 0, nHit = 0;
    double traceSec = 0.0;
    if (!raysFile.empty()) {
        std::vector<float> rf;
        std::vector<double> rd;
        if (rayFormat == "f32") { rf = readAll<float>(raysFile); nRays = (long)(rf.size() / 8); }
        else { rd = readAll<double>(raysFile); nRays = (long)(rd.size() / 8); }
        std::vector<long> ids;
        for (long i = 0; i < nRays; i += stride) { ids.push_back(i); if (limit > 0 && (long)ids.size() >= limit) break; }
        nUsed = (long)ids.size();
        std::vector<int32_t> outPrim(nUsed, -1);
        std::vector<double> outT(nUsed, 0.0);
        std::vector<uint8_t> outAny(nUsed, 0);
        const bool closest = (mode == "closest");

        auto work = [&](long b, long e) {
            for (long k = b; k < e; k++) {
                long i = ids[k];
                double o[3], d[3], tmax;
                if (rayFormat == "f32") {
                    const float* p = &rf[i * 8];
                    for (int c = 0; c < 3; c++) { o[c] = (double)p[c]; d[c] = (double)p[3 + c]; }
                    tmax = (double)p[7];
                } else {
                    const double* p = &rd[i * 8];
                    for (int c = 0; c < 3; c++) { o[c] = p[c]; d[c] = p[3 + c]; }
                    tmax = p[7];
                }
                Ray ray(Point3d(o[0], o[1], o[2]), Vector3d(d[0], d[1], d[2]), tmax);
                if (closest) {
                    SurfaceInteraction isect;
                    if (accel.intersect(ray, &isect)) {
                        outPrim[k] = primIndex[isect.primitive()];
                        outT[k] = ray.maxDist();   // primitive.cc:52 stored tHit here
                    }
                } else {
                    outAny[k] = accel.intersect(ray) ? 1 : 0;
                }
            }
        };

        double best = 1e30, sum = 0.0;
        for (int r = 0; r < warmup + repeat; r++) {
            double ts = now();
            std::vector<std::thread> pool;
            // dynamic chunks so the timing is not dominated by one slow slice
            std::atomic<long> cursor(0);
            const long chunk = 4096;
            for (int t = 0; t < threads; t++) {
                pool.emplace_back([&]() {
                    for (;;) {
                        long b = cursor.fetch_add(chunk);
                        if (b >= nUsed) break;
                        work(b, std::min(nUsed, b + chunk));
                    }
                });
            }
            for (auto& th : pool) th.join();
            const double dt = now() - ts;
            if (r >= warmup) { best = std::min(best, dt); sum += dt; }
        }
        traceSec = sum / repeat;   // mean over the timed repeats
        for (long k = 0; k < nUsed; k++) nHit += closest ? (outPrim[k] >= 0) : outAny[k];

        if (!out.empty()) {
            std::ofstream ofs(out, std::ios::binary);
            if (closest) {
                ofs.write((const char*)outPrim.data(), outPrim.size() * sizeof(int32_t));
                ofs.write((const char*)outT.data(), outT.size() * sizeof(double));
            } else {
                ofs.write((const char*)outAny.data(), outAny.size());
            }
        }
    }

    printf("{\"impl\": \"reference\", \"n_prims\": %zu, \"n_nodes\": %zu, \"build_s\": %.6f, "
           "\"mode\": \"%s\", \"n_rays\": %ld, \"n_hit\": %ld, \"threads\": %d, \"trace_s\": %.6f, "
           "\"mrays_s\": %.6f, \"simd\": %d}\n",
           prims.size(), accel.nodes_.size(), buildSec, mode.c_str(), nUsed, nHit, threads, traceSec,
           traceSec > 0 ? nUsed / traceSec * 1e-6 : 0.0, simd);
    return 0;
}
