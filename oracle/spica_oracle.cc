// TEST INFRASTRUCTURE ONLY -- see spica_oracle.h for the contract and the parity status.
//
// CPU restatement of the reference's ray-cast hot path, written from the reference's behaviour
// (file:line cited per function, relative to /root/reference/sources).  C++ rather than C only
// because the reference's builder relies on libstdc++'s std::nth_element / std::partition
// (accelerators/bvh.cc:176,224) and the tree is only reproducible with the same algorithms.

#include "spica_oracle.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

const double kInfty = 1.0e32;   // core/common.h:54
const double kEps   = 1.0e-12;  // core/common.h:55

struct Ray {
    double o[3], d[3], inv[3], maxDist;
};

// core/ray.cc:11-19 (ctor normalises), :43-47 (invdir); core/vector3d_detail.h:134-137
// (operator/=(double) multiplies by 1.0/s), :185-200 (norm = sqrt(dot), dot = x*x + y*y + z*z).
bool rayInit(const double o[3], const double d[3], double maxDist, Ray* r) {
    const double sq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    const double nrm = ::sqrt(sq);
    if (nrm == 0.0) return false;  // Assertion(s != 0.0) aborts in the reference
    const double s = 1.0 / nrm;
    for (int k = 0; k < 3; k++) {
        r->o[k] = o[k];
        r->d[k] = d[k] * s;
        r->inv[k] = (r->d[k] == 0.0) ? kInfty : 1.0 / r->d[k];
    }
    r->maxDist = maxDist;
    return true;
}

// core/triangle.cc:98-117 and :159-178 (identical arithmetic).
inline bool triIntersect(const double* p, const Ray& r, double* tHit, double* uOut, double* vOut) {
    const double e1x = p[3] - p[0], e1y = p[4] - p[1], e1z = p[5] - p[2];
    const double e2x = p[6] - p[0], e2y = p[7] - p[1], e2z = p[8] - p[2];
    // pVec = cross(dir, e2)   (core/vector3d_detail.h:168-174)
    const double px = r.d[1] * e2z - r.d[2] * e2y;
    const double py = r.d[2] * e2x - r.d[0] * e2z;
    const double pz = r.d[0] * e2y - r.d[1] * e2x;
    const double det = e1x * px + e1y * py + e1z * pz;
    if (det > -kEps && det < kEps) return false;
    const double invdet = 1.0 / det;
    const double tx = r.o[0] - p[0], ty = r.o[1] - p[1], tz = r.o[2] - p[2];
    const double u = (tx * px + ty * py + tz * pz) * invdet;
    if (u < 0.0 || u > 1.0) return false;
    // qVec = cross(tVec, e1)
    const double qx = ty * e1z - tz * e1y;
    const double qy = tz * e1x - tx * e1z;
    const double qz = tx * e1y - ty * e1x;
    const double v = (r.d[0] * qx + r.d[1] * qy + r.d[2] * qz) * invdet;
    if (v < 0.0 || u + v > 1.0) return false;
    const double t = (e2x * qx + e2y * qy + e2z * qz) * invdet;
    if (t <= kEps || t > r.maxDist) return false;
    *tHit = t;
    if (uOut) *uOut = u;
    if (vOut) *vOut = v;
    return true;
}

// core/bounds3d_detail.h:64-79
inline bool boxIntersect(const double lo[3], const double hi[3], const Ray& r, double* tNear, double* tFar) {
    double t0 = 0.0, t1 = r.maxDist;
    for (int i = 0; i < 3; i++) {
        double tt0 = (lo[i] - r.o[i]) * r.inv[i];
        double tt1 = (hi[i] - r.o[i]) * r.inv[i];
        if (tt0 > tt1) std::swap(tt0, tt1);
        t0 = std::max(t0, tt0);
        t1 = std::min(t1, tt1);
        if (t0 > t1) return false;
    }
    if (tNear) *tNear = t0;
    if (tFar) *tFar = t1;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Builder: accelerators/bvh.cc:139-237
// ---------------------------------------------------------------------------------------------
struct Bounds {
    double lo[3], hi[3];
    Bounds() {  // core/bounds3d_detail.h:13-21
        for (int k = 0; k < 3; k++) { lo[k] = DBL_MAX; hi[k] = -DBL_MAX; }
    }
    static Bounds merge(const Bounds& a, const Bounds& b) {  // :82-86
        Bounds r;
        for (int k = 0; k < 3; k++) { r.lo[k] = std::min(a.lo[k], b.lo[k]); r.hi[k] = std::max(a.hi[k], b.hi[k]); }
        return r;
    }
    void mergePoint(const double p[3]) {  // :95-98 (minimum(p, posMin_))
        for (int k = 0; k < 3; k++) { lo[k] = std::min(p[k], lo[k]); hi[k] = std::max(p[k], hi[k]); }
    }
    int maximumExtent() const {  // :55-61
        const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx >= dy && dx >= dz) return 0;
        if (dy >= dz) return 1;
        return 2;
    }
    double area() const {  // :107-114
        const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        const double xy = std::abs(dx * dy), yz = std::abs(dy * dz), zx = std::abs(dz * dx);
        return 2.0 * (xy + yz + zx);
    }
};

struct PrimInfo {  // accelerators/bvh.h:14-26
    int primIdx;
    double centroid[3];
    Bounds bounds;
};

struct Builder {
    std::vector<PrimInfo> data;
    so_node* nodes;
    int64_t count = 0;

    int32_t rec(int start, int end) {
        if (start == end) return -1;
        const int32_t me = (int32_t)count++;
        Bounds bounds;
        for (int i = start; i < end; i++) bounds = Bounds::merge(bounds, data[i].bounds);
        const int nprims = end - start;
        so_node nd;
        for (int k = 0; k < 3; k++) { nd.lo[k] = bounds.lo[k]; nd.hi[k] = bounds.hi[k]; }
        if (nprims == 1) {
            nd.left = nd.right = -1; nd.prim = data[start].primIdx; nd.axis = 0;
            nodes[me] = nd;
            return me;
        }
        Bounds cb;
        for (int i = start; i < end; i++) cb.mergePoint(data[i].centroid);
        const int axis = cb.maximumExtent();
        int mid = (start + end) / 2;
        if (nprims <= 8) {
            std::nth_element(data.begin() + start, data.begin() + mid, data.begin() + end,
                             [axis](const PrimInfo& a, const PrimInfo& b) { return a.centroid[axis] < b.centroid[axis]; });
        } else {
            const int nBuckets = 16;
            int cnt[nBuckets] = {0};
            Bounds bb[nBuckets];
            const double cmin = cb.lo[axis], cmax = cb.hi[axis];
            const double idenom = 1.0 / (std::abs(cmax - cmin) + kEps);
            for (int i = start; i < end; i++) {
                const double numer = data[i].centroid[axis] - cb.lo[axis];
                int b = static_cast<int>(nBuckets * std::abs(numer) * idenom);
                if (b == nBuckets) b = nBuckets - 1;
                cnt[b]++;
                bb[b] = Bounds::merge(bb[b], data[i].bounds);
            }
            double cost[nBuckets - 1] = {0};
            for (int i = 0; i < nBuckets - 1; i++) {
                Bounds b0, b1;
                int c0 = 0, c1 = 0;
                for (int j = 0; j <= i; j++) { b0 = Bounds::merge(b0, bb[j]); c0 += cnt[j]; }
                for (int j = i + 1; j < nBuckets; j++) { b1 = Bounds::merge(b1, bb[j]); c1 += cnt[j]; }
                cost[i] += 0.125 + (c0 * b0.area() + c1 * b1.area()) / bounds.area();
            }
            double minCost = cost[0];
            int minSplit = 0;
            for (int i = 1; i < nBuckets - 1; i++) {
                if (minCost > cost[i]) { minCost = cost[i]; minSplit = i; }
            }
            if (minCost < nprims) {
                // CompareToBucket, accelerators/bvh.cc:72-95
                auto it = std::partition(data.begin() + start, data.begin() + end,
                    [&](const PrimInfo& p) {
                        const double lo = cb.lo[axis], hi = cb.hi[axis];
                        const double inv = (1.0) / (std::abs(hi - lo) + kEps);
                        const double diff = std::abs(p.centroid[axis] - lo);
                        int b = static_cast<int>(nBuckets * diff * inv);
                        if (b >= nBuckets) b = nBuckets - 1;
                        return b <= minSplit;
                    });
                mid = (int)(it - data.begin());
            }
        }
        const int32_t l = rec(start, mid);
        const int32_t r = rec(mid, end);
        nd.left = l; nd.right = r; nd.prim = -1; nd.axis = axis;
        nodes[me] = nd;
        return me;
    }
};

struct RayReader {
    const void* rays;
    int f64;
    bool get(int64_t i, Ray* r) const {
        double o[3], d[3], tmax;
        if (f64) {
            const double* p = (const double*)rays + i * 8;
            for (int k = 0; k < 3; k++) { o[k] = p[k]; d[k] = p[3 + k]; }
            tmax = p[7];
        } else {
            const float* p = (const float*)rays + i * 8;
            for (int k = 0; k < 3; k++) { o[k] = (double)p[k]; d[k] = (double)p[3 + k]; }
            tmax = (double)p[7];
        }
        return rayInit(o, d, tmax, r);
    }
};

template <class F>
void parallelFor(int64_t n, int threads, F&& f) {
    if (threads <= 1 || n < 1024) { f(0, n, 0); return; }
    std::atomic<int64_t> cursor(0);
    const int64_t chunk = 2048;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([&, t]() {
            for (;;) {
                int64_t b = cursor.fetch_add(chunk);
                if (b >= n) break;
                f(b, std::min(n, b + chunk), t);
            }
        });
    }
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

int so_ray_init(const double o[3], const double d[3], double dir_out[3], double invdir_out[3]) {
    Ray r;
    if (!rayInit(o, d, kInfty, &r)) return 0;
    for (int k = 0; k < 3; k++) { dir_out[k] = r.d[k]; invdir_out[k] = r.inv[k]; }
    return 1;
}

int so_triangle_intersect(const double tri[9], const double org[3], const double dir[3],
                          double max_dist, double* t, double* u, double* v) {
    Ray r;
    for (int k = 0; k < 3; k++) { r.o[k] = org[k]; r.d[k] = dir[k]; r.inv[k] = 0; }
    r.maxDist = max_dist;
    return triIntersect(tri, r, t, u, v) ? 1 : 0;
}

int so_bounds_intersect(const double lo[3], const double hi[3], const double org[3],
                        const double invdir[3], double max_dist, double* t_near, double* t_far) {
    Ray r;
    for (int k = 0; k < 3; k++) { r.o[k] = org[k]; r.d[k] = 0; r.inv[k] = invdir[k]; }
    r.maxDist = max_dist;
    return boxIntersect(lo, hi, r, t_near, t_far) ? 1 : 0;
}

int64_t so_bvh_build(const double* tris, int64_t n, so_node* nodes) {
    if (n <= 0) return 0;
    Builder b;
    b.nodes = nodes;
    b.data.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        const double* p = tris + i * 9;
        PrimInfo& pi = b.data[(size_t)i];
        pi.primIdx = (int)i;
        // core/triangle.cc:199-203
        for (int k = 0; k < 3; k++) {
            pi.bounds.lo[k] = std::min(p[k], std::min(p[3 + k], p[6 + k]));
            pi.bounds.hi[k] = std::max(p[k], std::max(p[3 + k], p[6 + k]));
            pi.centroid[k] = (pi.bounds.hi[k] + pi.bounds.lo[k]) * 0.5;  // bvh.h:24
        }
    }
    b.rec(0, (int)n);
    return b.count;
}

void so_trace_closest(const so_node* nodes, int32_t root, const double* tris, const void* rays,
                      int ray_f64, int64_t n, int32_t* prim, double* t, double* u, double* v,
                      int threads) {
    RayReader rr{rays, ray_f64};
    parallelFor(n, threads, [&](int64_t b, int64_t e, int) {
        std::vector<int32_t> stack;
        for (int64_t i = b; i < e; i++) {
            Ray r;
            prim[i] = -1; t[i] = 0.0;
            if (u) u[i] = 0.0;
            if (v) v[i] = 0.0;
            if (root < 0 || !rr.get(i, &r)) continue;
            stack.clear();
            stack.push_back(root);
            while (!stack.empty()) {
                const so_node& nd = nodes[stack.back()];
                stack.pop_back();
                if (nd.prim >= 0) {
                    double th, uu, vv;
                    if (triIntersect(tris + (int64_t)nd.prim * 9, r, &th, &uu, &vv)) {
                        r.maxDist = th;  // core/primitive.cc:52
                        prim[i] = nd.prim; t[i] = th;
                        if (u) u[i] = uu;
                        if (v) v[i] = vv;
                    }
                } else if (boxIntersect(nd.lo, nd.hi, r, nullptr, nullptr)) {
                    if (nd.left >= 0) stack.push_back(nd.left);
                    if (nd.right >= 0) stack.push_back(nd.right);
                }
            }
        }
    });
}

void so_trace_any(const so_node* nodes, int32_t root, const double* tris, const void* rays,
                  int ray_f64, int64_t n, uint8_t* occluded, int threads) {
    RayReader rr{rays, ray_f64};
    parallelFor(n, threads, [&](int64_t b, int64_t e, int) {
        std::vector<int32_t> stack;
        for (int64_t i = b; i < e; i++) {
            Ray r;
            occluded[i] = 0;
            if (root < 0 || !rr.get(i, &r)) continue;
            stack.clear();
            stack.push_back(root);
            while (!stack.empty()) {
                const so_node& nd = nodes[stack.back()];
                stack.pop_back();
                if (nd.prim >= 0) {
                    double th;
                    if (triIntersect(tris + (int64_t)nd.prim * 9, r, &th, nullptr, nullptr)) { occluded[i] = 1; break; }
                } else if (boxIntersect(nd.lo, nd.hi, r, nullptr, nullptr)) {
                    if (nd.left >= 0) stack.push_back(nd.left);
                    if (nd.right >= 0) stack.push_back(nd.right);
                }
            }
        }
    });
}

void so_trace_bruteforce(const double* tris, int64_t n_tris, const void* rays, int ray_f64,
                         int64_t n, int32_t* prim, double* t, int threads) {
    RayReader rr{rays, ray_f64};
    parallelFor(n, threads, [&](int64_t b, int64_t e, int) {
        for (int64_t i = b; i < e; i++) {
            Ray r;
            prim[i] = -1; t[i] = 0.0;
            if (!rr.get(i, &r)) continue;
            for (int64_t k = 0; k < n_tris; k++) {
                double th;
                if (triIntersect(tris + k * 9, r, &th, nullptr, nullptr)) { r.maxDist = th; prim[i] = (int32_t)k; t[i] = th; }
            }
        }
    });
}

// Ordered, tmax-pruned traversal used ONLY to define "algorithmic bytes per ray" (SURVEY 8d):
// an inner visit fetches the two child boxes and tests both over [0, tmax]; the nearer hit child
// is entered first, the farther is stacked with its entry distance and dropped when popped if
// that entry distance is already beyond the shrunken tmax.  A leaf visit tests one triangle.
void so_count_ordered_visits(const so_node* nodes, int32_t root, const double* tris,
                             const void* rays, int ray_f64, int64_t n, int64_t* sum_inner,
                             int64_t* sum_leaf, int threads) {
    RayReader rr{rays, ray_f64};
    std::atomic<int64_t> accInner(0), accLeaf(0);
    parallelFor(n, threads, [&](int64_t b, int64_t e, int) {
        struct Item { int32_t node; double tNear; };
        std::vector<Item> stack;
        int64_t nin = 0, nlf = 0;
        for (int64_t i = b; i < e; i++) {
            Ray r;
            if (root < 0 || !rr.get(i, &r)) continue;
            double tn, tf;
            if (nodes[root].prim < 0 && !boxIntersect(nodes[root].lo, nodes[root].hi, r, &tn, &tf)) continue;
            stack.clear();
            stack.push_back({root, 0.0});
            while (!stack.empty()) {
                Item it = stack.back();
                stack.pop_back();
                if (it.tNear > r.maxDist) continue;
                const so_node& nd = nodes[it.node];
                if (nd.prim >= 0) {
                    nlf++;
                    double th;
                    if (triIntersect(tris + (int64_t)nd.prim * 9, r, &th, nullptr, nullptr)) r.maxDist = th;
                    continue;
                }
                nin++;
                double tl = 0, tr = 0, tmp;
                bool hl = false, hr = false;
                if (nd.left >= 0) {
                    const so_node& c = nodes[nd.left];
                    hl = boxIntersect(c.lo, c.hi, r, &tl, &tmp);
                }
                if (nd.right >= 0) {
                    const so_node& c = nodes[nd.right];
                    hr = boxIntersect(c.lo, c.hi, r, &tr, &tmp);
                }
                if (hl && hr) {
                    if (tl <= tr) { stack.push_back({nd.right, tr}); stack.push_back({nd.left, tl}); }
                    else          { stack.push_back({nd.left, tl}); stack.push_back({nd.right, tr}); }
                } else if (hl) stack.push_back({nd.left, tl});
                else if (hr) stack.push_back({nd.right, tr});
            }
        }
        accInner += nin; accLeaf += nlf;
    });
    *sum_inner = accInner.load();
    *sum_leaf = accLeaf.load();
}

}  // extern "C"
