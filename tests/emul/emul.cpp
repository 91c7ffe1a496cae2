// TEST INFRASTRUCTURE ONLY.  g++ build of the product's host-side builder (bvh_host.cpp) plus the
// __host__ instantiation of the per-ray traversal (trace_core.h), so that node encoding, octant
// ordering, stack handling and the exact triangle test can be checked against the oracle on a
// machine without a GPU.  Nothing in the product loads this library.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../spica_b200/csrc/bvh_host.h"
#include "../../spica_b200/csrc/trace_core.h"

using namespace spb;

struct Emul {
    std::vector<double> verts;
    int64_t n = 0;
    BinaryBVH bin;
    HostBVH bvh;
    SceneParams sp{};
    std::vector<TriF32> pre;            // TriF64 trees: the rounded records of the pre-test (api.cu: roundTrisKernel)
    std::string err;
};

extern "C" {

Emul* emul_create(const double* verts, int64_t n, int max_leaf, int bins, const spb_import_node* imp,
                  int64_t n_imp, int32_t root) {
    Emul* e = new Emul();
    e->verts.assign(verts, verts + n * 9);
    e->n = n;
    bool ok = true;
    if (imp) ok = import_binary(imp, n_imp, root, e->verts.data(), n, &e->bin, &e->err);
    else build_binary_sah(e->verts.data(), n, bins, &e->bin);
    if (ok) ok = encode_wide(e->bin, e->verts.data(), n, max_leaf, &e->bvh, &e->err);
    if (!ok) return e;
    SceneParams& sp = e->sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.f32_one = 0x3f800000u;
    sp.nodes = e->bvh.nodes.data();
    sp.tris = e->bvh.tris.data();
    sp.empty = (n == 0 || e->bvh.nodes.empty()) ? 1 : 0;
    sp.n_tris = (int32_t)n;
    sp.tri_format = e->bvh.tri_format;
    sp.max_depth = e->bvh.max_depth;
    sp.inflate = (float)e->bvh.inflate;
    for (int k = 0; k < 3; k++) { sp.wlo[k] = e->bvh.wlo[k] - 2 * e->bvh.inflate; sp.whi[k] = e->bvh.whi[k] + 2 * e->bvh.inflate; }
    { double m = 0.0; for (int k = 0; k < 3; k++) m = std::max(m, std::max(std::abs(sp.wlo[k]), std::abs(sp.whi[k]))); sp.max_coord = (float)(m * 1.0000002); }
    sp.pre_tris = (const TriF32*)e->bvh.tris.data();
    sp.pre_round = 0.f;
    if (!sp.empty && e->bvh.tri_format == 1) {
        const TriF64* t64 = (const TriF64*)e->bvh.tris.data();
        e->pre.resize((size_t)n);
        for (int64_t i = 0; i < n; i++) {
            TriF32& o = e->pre[(size_t)i];
            for (int k = 0; k < 3; k++) { o.v0[k] = (float)t64[i].v[k]; o.v1[k] = (float)t64[i].v[3 + k]; o.v2[k] = (float)t64[i].v[6 + k]; }
            o.id = t64[i].id; o.rank = t64[i].rank; o.pad = 0;
        }
        sp.pre_tris = e->pre.data();
        sp.pre_round = sp.max_coord * 1.1920929e-07f;
    }
    return e;
}
const char* emul_error(Emul* e) { return e->err.c_str(); }
void emul_destroy(Emul* e) { delete e; }
void emul_stats(Emul* e, int64_t* out) {
    out[0] = (int64_t)e->bvh.nodes.size(); out[1] = e->bvh.tri_format; out[2] = e->bvh.max_depth;
    out[3] = (int64_t)e->bin.nodes.size();
}
double emul_sah(Emul* e) { return e->bvh.sah_cost; }

// rays: n x 8 (f32 or f64); out: prim i32[n], t f64[n], u f32[n], v f32[n]; counters u64[3]
void emul_trace(Emul* e, const void* rays, int f64, int64_t n, int any, int32_t* prim, double* t, float* u,
                float* v, unsigned long long* counters, int threads, int mode) {
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    std::vector<TraceCounters> ctrs((size_t)threads, TraceCounters{0, 0, 0});
    for (int th = 0; th < threads; th++) {
        pool.emplace_back([&, th]() {
            const int64_t b = n * th / threads, en = n * (th + 1) / threads;
            for (int64_t i = b; i < en; i++) {
                double o[3], d[3], tmax;
                if (f64) { const double* p = (const double*)rays + i * 8; for (int k = 0; k < 3; k++) { o[k] = p[k]; d[k] = p[3 + k]; } tmax = p[7]; }
                else { const float* p = (const float*)rays + i * 8; for (int k = 0; k < 3; k++) { o[k] = p[k]; d[k] = p[3 + k]; } tmax = p[7]; }
                RayState r;
                const bool valid = rayBegin(e->sp, o[0], o[1], o[2], d[0], d[1], d[2], tmax, r);
                if (mode == 2 && e->sp.tri_format == 0) {   // early select + deferred exact test (kernel variant 6)
                    if (any) traceRayDeferred<true>(e->sp, r, valid, &ctrs[th]); else traceRayDeferred<false>(e->sp, r, valid, &ctrs[th]);
                } else if (mode == 3) {   // three node visits per triangle phase (kernel variant 5)
                    if (e->sp.tri_format == 0) { if (any) traceRayKVisits<0, true, 3>(e->sp, r, valid, &ctrs[th]); else traceRayKVisits<0, false, 3>(e->sp, r, valid, &ctrs[th]); }
                    else { if (any) traceRayKVisits<1, true, 3>(e->sp, r, valid, &ctrs[th]); else traceRayKVisits<1, false, 3>(e->sp, r, valid, &ctrs[th]); }
                } else if (mode == 1) {   // early-select phase order (kernel variants 3, 4)
                    if (e->sp.tri_format == 0) { if (any) traceRayEarlySelect<0, true>(e->sp, r, valid, &ctrs[th]); else traceRayEarlySelect<0, false>(e->sp, r, valid, &ctrs[th]); }
                    else { if (any) traceRayEarlySelect<1, true>(e->sp, r, valid, &ctrs[th]); else traceRayEarlySelect<1, false>(e->sp, r, valid, &ctrs[th]); }
                } else if (e->sp.tri_format == 0) { if (any) traceRay<0, true>(e->sp, r, valid, &ctrs[th]); else traceRay<0, false>(e->sp, r, valid, &ctrs[th]); }
                else { if (any) traceRay<1, true>(e->sp, r, valid, &ctrs[th]); else traceRay<1, false>(e->sp, r, valid, &ctrs[th]); }
                prim[i] = r.best_prim;
                t[i] = r.best_prim >= 0 ? r.best_t : 0.0;
                u[i] = r.best_u; v[i] = r.best_v;
            }
        });
    }
    for (auto& th : pool) th.join();
    counters[0] = counters[1] = counters[2] = 0;
    for (auto& c : ctrs) { counters[0] += c.nodes; counters[1] += c.tris; counters[2] += c.exact; }
}

}  // extern "C"
