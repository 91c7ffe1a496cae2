#include "bvh_host.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <future>
#include <queue>
#include <thread>

namespace spb {
namespace {

struct Box {
    double lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; k++) { lo[k] = DBL_MAX; hi[k] = -DBL_MAX; } }
    void grow(const Box& b) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    void grow(const double p[3]) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
    double area() const {
        const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0) return 0.0;
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
};

inline Box triBox(const double* v) {
    Box b;
    for (int k = 0; k < 3; k++) {
        b.lo[k] = std::min(v[k], std::min(v[3 + k], v[6 + k]));
        b.hi[k] = std::max(v[k], std::max(v[3 + k], v[6 + k]));
    }
    return b;
}

struct PrimRef { Box b; double c[3]; int32_t id; };   // kept physically in leaf order while the tree is built: every pass below is sequential

struct SahBuilder {
    std::vector<PrimRef>& prims;           // partitioned in place (position i <-> order[i])
    std::vector<int32_t>& order;
    std::vector<BinNode>& nodes;
    std::atomic<int32_t> nextNode{0};
    int bins;

    SahBuilder(std::vector<PrimRef>& p, std::vector<int32_t>& o, std::vector<BinNode>& n, int b)
        : prims(p), order(o), nodes(n), bins(b) {}

    // The few nodes near the root hold most of the primitives while only one or two threads are busy with
    // them: their bounds and bin passes run in slices on all cores (min / max / counts merge exactly, so the
    // tree is the one the serial passes give).
    static constexpr int32_t kSliceMin = 1 << 17;
    int slicesFor(int32_t count, int depth) const {
        if (count < 2 * kSliceMin || depth > 4) return 1;
        const int hw = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        return std::max(1, std::min(hw >> depth, (int)(count / kSliceMin)));
    }
    template <class F>
    static void forSlices(int slices, int32_t first, int32_t count, F&& fn) {      // fn(slice, first, count)
        if (slices <= 1) { fn(0, first, count); return; }
        std::vector<std::future<void>> futs;
        for (int s = 1; s < slices; s++) {
            const int32_t a = first + (int32_t)((int64_t)count * s / slices), b = first + (int32_t)((int64_t)count * (s + 1) / slices);
            futs.push_back(std::async(std::launch::async, [&fn, s, a, b]() { fn(s, a, b - a); }));
        }
        fn(0, first, (int32_t)((int64_t)count / slices));
        for (auto& f : futs) f.get();
    }

    int32_t build(int32_t first, int32_t count, int depth) {
        const int32_t me = nextNode.fetch_add(1);
        Box bounds, cb;
        bounds.reset(); cb.reset();
        const int slices = slicesFor(count, depth);
        {
            std::vector<Box> sb((size_t)slices), sc((size_t)slices);
            forSlices(slices, first, count, [&](int sl, int32_t a, int32_t n) {
                Box b1, c1; b1.reset(); c1.reset();
                for (int32_t i = a; i < a + n; i++) { const PrimRef& p = prims[i]; b1.grow(p.b); c1.grow(p.c); }
                sb[sl] = b1; sc[sl] = c1;
            });
            for (int sl = 0; sl < slices; sl++) { bounds.grow(sb[sl]); cb.grow(sc[sl]); }
        }
        BinNode nd;
        for (int k = 0; k < 3; k++) { nd.lo[k] = bounds.lo[k]; nd.hi[k] = bounds.hi[k]; }
        nd.first = first; nd.count = count; nd.left = nd.right = -1;
        if (count == 1) { nodes[me] = nd; return me; }

        // binned SAH over the three axes
        const int B = bins;
        double bestCost = DBL_MAX;
        int bestAxis = -1, bestSplit = -1;
        // per-thread scratch: the bins are dead before the recursion below starts, and a million small nodes
        // would otherwise pay three heap allocations each
        thread_local std::vector<Box> bb, rightAcc;
        thread_local std::vector<int32_t> cnt;
        if ((int)bb.size() < B) { bb.resize(B); rightAcc.resize(B); cnt.resize(B); }
        for (int axis = 0; axis < 3; axis++) {
            const double cmin = cb.lo[axis], cmax = cb.hi[axis];
            if (!(cmax > cmin)) continue;
            const double scale = B / (cmax - cmin);
            if (count * 2 <= B) {
                // Few primitives: only the occupied bins are built and swept (at most `count` of them, in
                // ascending order).  The full sweep below gives every split position inside a run of empty
                // bins the cost of the occupied bin before it and keeps the first strict minimum, i.e. that
                // occupied bin: same cost, same split, same tree -- without touching 3 x B bins for each of
                // the million nodes at the bottom of the tree.
                int ob[128]; Box obox[128], racc[128]; int32_t ocnt[128]; int m = 0;
                for (int32_t i = first; i < first + count; i++) {
                    const PrimRef& p = prims[i];
                    int b = (int)((p.c[axis] - cmin) * scale);
                    b = b < 0 ? 0 : (b >= B ? B - 1 : b);
                    int k = 0;
                    while (k < m && ob[k] < b) k++;
                    if (k == m || ob[k] != b) {
                        for (int j = m; j > k; j--) { ob[j] = ob[j - 1]; obox[j] = obox[j - 1]; ocnt[j] = ocnt[j - 1]; }
                        ob[k] = b; obox[k].reset(); ocnt[k] = 0; m++;
                    }
                    ocnt[k]++;
                    obox[k].grow(p.b);
                }
                Box acc; acc.reset();
                for (int k = m - 1; k > 0; k--) { acc.grow(obox[k]); racc[k] = acc; }
                acc.reset();
                int32_t nl = 0;
                for (int k = 0; k + 1 < m; k++) {
                    acc.grow(obox[k]);
                    nl += ocnt[k];
                    const double cost = acc.area() * nl + racc[k + 1].area() * (count - nl);
                    if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestSplit = ob[k]; }
                }
                continue;
            }
            for (int b = 0; b < B; b++) { bb[b].reset(); cnt[b] = 0; }
            if (slices <= 1) {
                for (int32_t i = first; i < first + count; i++) {
                    const PrimRef& p = prims[i];
                    int b = (int)((p.c[axis] - cmin) * scale);
                    b = b < 0 ? 0 : (b >= B ? B - 1 : b);
                    cnt[b]++;
                    bb[b].grow(p.b);
                }
            } else {
                std::vector<Box> sbb((size_t)slices * B);
                std::vector<int32_t> scnt((size_t)slices * B, 0);
                for (auto& x : sbb) x.reset();
                forSlices(slices, first, count, [&](int sl, int32_t a, int32_t n) {
                    Box* mb = &sbb[(size_t)sl * B]; int32_t* mc = &scnt[(size_t)sl * B];
                    for (int32_t i = a; i < a + n; i++) {
                        const PrimRef& p = prims[i];
                        int b = (int)((p.c[axis] - cmin) * scale);
                        b = b < 0 ? 0 : (b >= B ? B - 1 : b);
                        mc[b]++;
                        mb[b].grow(p.b);
                    }
                });
                for (int sl = 0; sl < slices; sl++)
                    for (int b = 0; b < B; b++) { cnt[b] += scnt[(size_t)sl * B + b]; if (scnt[(size_t)sl * B + b]) bb[b].grow(sbb[(size_t)sl * B + b]); }
            }
            Box acc; acc.reset();
            for (int b = B - 1; b > 0; b--) { acc.grow(bb[b]); rightAcc[b] = acc; }
            acc.reset();
            int32_t nl = 0;
            for (int b = 0; b < B - 1; b++) {
                acc.grow(bb[b]);
                nl += cnt[b];
                const int32_t nr = count - nl;
                if (nl == 0 || nr == 0) continue;
                const double cost = acc.area() * nl + rightAcc[b + 1].area() * nr;
                if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestSplit = b; }
            }
        }

        int32_t mid;
        if (bestAxis < 0) {
            // all centroids coincide (e.g. the two triangles of a quad whose diagonal spans its bounding box): split by
            // index -- in ascending primitive order, so that the tree does not depend on the order earlier partitions left
            std::sort(prims.begin() + first, prims.begin() + first + count, [](const PrimRef& a, const PrimRef& b) { return a.id < b.id; });
            mid = first + count / 2;
        } else {
            const double cmin = cb.lo[bestAxis];
            const double scale = B / (cb.hi[bestAxis] - cmin);
            auto it = std::partition(prims.begin() + first, prims.begin() + first + count, [&](const PrimRef& p) {
                int b = (int)((p.c[bestAxis] - cmin) * scale);
                b = b < 0 ? 0 : (b >= B ? B - 1 : b);
                return b <= bestSplit;
            });
            mid = (int32_t)(it - prims.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }

        int32_t l, r;
        if (count > 32768 && depth < 6) {
            auto fut = std::async(std::launch::async, [&, first, mid, depth]() { return build(first, mid - first, depth + 1); });
            r = build(mid, first + count - mid, depth + 1);
            l = fut.get();
        } else {
            l = build(first, mid - first, depth + 1);
            r = build(mid, first + count - mid, depth + 1);
        }
        nd.left = l; nd.right = r;
        nodes[me] = nd;
        return me;
    }
};

inline float floatDown(double x) {
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -INFINITY);
    return f;
}

inline bool isFloatExact(double x) { return (double)(float)x == x; }

}  // namespace

void build_binary_sah(const double* verts, int64_t n, int bins, BinaryBVH* out) {
    out->nodes.clear(); out->order.clear(); out->root = -1; out->imported = false;
    if (n <= 0) return;
    if (bins < 4) bins = 4;
    if (bins > 256) bins = 256;
    std::vector<PrimRef> prims((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        prims[i].b = triBox(verts + i * 9);
        for (int k = 0; k < 3; k++) prims[i].c[k] = 0.5 * (prims[i].b.lo[k] + prims[i].b.hi[k]);
        prims[i].id = (int32_t)i;
    }
    out->order.resize((size_t)n);
    out->nodes.resize((size_t)(2 * n - 1));
    SahBuilder b(prims, out->order, out->nodes, bins);
    out->root = b.build(0, (int32_t)n, 0);
    out->nodes.resize((size_t)b.nextNode.load());
    for (int64_t i = 0; i < n; i++) out->order[i] = prims[i].id;
}

bool import_binary(const spb_import_node* in, int64_t n_nodes, int32_t root, const double* verts,
                   int64_t n_tris, BinaryBVH* out, std::string* err) {
    out->nodes.clear(); out->order.clear(); out->root = -1; out->imported = true;
    if (n_nodes <= 0 || root < 0 || root >= n_nodes) { *err = "import: empty tree or bad root"; return false; }
    out->nodes.resize((size_t)n_nodes);
    std::vector<uint8_t> seenPrim((size_t)n_tris, 0);
    // iterative post-order: state 0 = enter, 1 = children done
    struct Item { int32_t node; int state; };
    std::vector<Item> st;
    std::vector<uint8_t> visited((size_t)n_nodes, 0);
    st.push_back({root, 0});
    while (!st.empty()) {
        Item it = st.back(); st.pop_back();
        const spb_import_node& s = in[it.node];
        BinNode& d = out->nodes[it.node];
        if (it.state == 0) {
            if (visited[it.node]) { *err = "import: node reachable twice (not a tree)"; return false; }
            visited[it.node] = 1;
            if (s.prim >= 0) {
                if (s.prim >= n_tris) { *err = "import: primitive index out of range"; return false; }
                if (seenPrim[s.prim]) { *err = "import: primitive referenced by two leaves"; return false; }
                seenPrim[s.prim] = 1;
                const Box b = triBox(verts + (int64_t)s.prim * 9);
                for (int k = 0; k < 3; k++) { d.lo[k] = b.lo[k]; d.hi[k] = b.hi[k]; }
                d.left = d.right = -1;
                d.first = (int32_t)out->order.size(); d.count = 1;
                out->order.push_back(s.prim);
            } else {
                if ((s.left < 0 && s.right < 0) || s.left >= n_nodes || s.right >= n_nodes) { *err = "import: bad child index"; return false; }
                d.first = (int32_t)out->order.size();   // leaves are appended left to right
                st.push_back({it.node, 1});
                if (s.right >= 0) st.push_back({s.right, 0});
                if (s.left >= 0) st.push_back({s.left, 0});   // popped first
            }
        } else {
            Box b; b.reset();
            d.left = s.left; d.right = s.right;
            for (int32_t c : {s.left, s.right}) {
                if (c < 0) continue;
                Box cbx; for (int k = 0; k < 3; k++) { cbx.lo[k] = out->nodes[c].lo[k]; cbx.hi[k] = out->nodes[c].hi[k]; }
                b.grow(cbx);
            }
            for (int k = 0; k < 3; k++) { d.lo[k] = b.lo[k]; d.hi[k] = b.hi[k]; }
            d.count = (int32_t)out->order.size() - d.first;
        }
    }
    if ((int64_t)out->order.size() != n_tris) { *err = "import: tree does not cover every primitive exactly once"; return false; }
    out->root = root;
    return true;
}

bool encode_wide(const BinaryBVH& bin, const double* verts, int64_t n_tris, int max_leaf, HostBVH* out,
                 std::string* err) {
    out->nodes.clear(); out->tris.clear();
    out->n_tris = n_tris; out->max_depth = 0; out->sah_cost = 0.0;
    out->n_wide = 0; out->tri_bytes_device = 0;
    out->n_binary_nodes = (int64_t)bin.nodes.size();
    if (max_leaf < 1) max_leaf = 1;
    if (max_leaf > 3) max_leaf = 3;
    if (n_tris <= 0 || bin.root < 0) return true;

    // vertex format + world box + inflation
    bool f32ok = true;
    Box world; world.reset();
    for (int64_t i = 0; i < n_tris; i++) {
        const double* v = verts + i * 9;
        for (int k = 0; k < 9 && f32ok; k++) f32ok = isFloatExact(v[k]);
        world.grow(triBox(v));
    }
    out->tri_format = f32ok ? 0 : 1;
    double mag = 0.0;
    for (int k = 0; k < 3; k++) {
        out->wlo[k] = world.lo[k]; out->whi[k] = world.hi[k];
        mag = std::max(mag, std::max(std::abs(world.lo[k]), std::abs(world.hi[k])));
        mag = std::max(mag, world.hi[k] - world.lo[k]);
    }
    if (!(mag < 1e30)) { *err = "scene coordinates too large for the float32 culling grid"; return false; }
    // fp32 culling slack: covers rounding of the culling ray (origin, direction) and of the slab
    // arithmetic; see DESIGN.md "conservative culling".
    const double inflate = std::max(mag * std::ldexp(1.0, -19), 1e-30);
    out->inflate = inflate;

    // rank for ties: imported tree -> position in left-to-right leaf order; own tree -> prim id
    std::vector<int32_t> rank((size_t)n_tris);
    if (bin.imported) { for (int64_t i = 0; i < n_tris; i++) rank[bin.order[i]] = (int32_t)i; }
    else { for (int64_t i = 0; i < n_tris; i++) rank[i] = (int32_t)i; }

    const size_t triSize = f32ok ? sizeof(TriF32) : sizeof(TriF64);
    out->tris.resize((size_t)n_tris * triSize);
    int64_t triCursor = 0;
    auto emitTri = [&](int32_t prim) {
        const double* v = verts + (int64_t)prim * 9;
        if (f32ok) {
            TriF32 t;
            for (int k = 0; k < 3; k++) { t.v0[k] = (float)v[k]; t.v1[k] = (float)v[3 + k]; t.v2[k] = (float)v[6 + k]; }
            t.id = prim; t.rank = rank[prim]; t.pad = 0;
            std::memcpy(&out->tris[(size_t)triCursor * triSize], &t, sizeof(t));
        } else {
            TriF64 t;
            for (int k = 0; k < 9; k++) t.v[k] = v[k];
            t.id = prim; t.rank = rank[prim];
            std::memcpy(&out->tris[(size_t)triCursor * triSize], &t, sizeof(t));
        }
        triCursor++;
    };

    auto leafLike = [&](int32_t b) {
        const BinNode& n = bin.nodes[b];
        return n.count <= max_leaf || (n.left < 0 && n.right < 0);
    };
    auto nodeArea = [&](int32_t b) {
        Box bx; for (int k = 0; k < 3; k++) { bx.lo[k] = bin.nodes[b].lo[k]; bx.hi[k] = bin.nodes[b].hi[k]; }
        return bx.area();
    };
    const double rootArea = std::max(nodeArea(bin.root), 1e-300);

    // ---- cost-optimal collapse (dynamic programme over the binary tree, after Ylitie, Karras, Laine,
    // "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs", HPG 2017, sec. 3.1):
    // C(n, i) = least SAH cost of turning the subtree of n into at most i children of one wide node,
    //   C(n, 1) = min(leaf: A_n * count * c_tri [count <= max_leaf], inner: A_n * c_node + D(n, 8))
    //   C(n, i) = min(C(n, i-1), D(n, i)),   D(n, j) = min_k C(left, k) + C(right, j - k).
    // The greedy expansion below fills 57 % of the child slots of the C2 tree; the optimal cut packs
    // the bottom of the tree better.  SPICA_BVH_COLLAPSE=0 keeps the greedy rule (also used above 12 M
    // triangles, where the tables would take more than 1 GB of host memory).
    bool dp = n_tris <= (int64_t)12 << 20;
    if (const char* e = std::getenv("SPICA_BVH_COLLAPSE")) dp = std::atoi(e) != 0 && dp;
    struct DpRow { float c[8]; uint8_t k[8]; uint8_t k8; uint8_t leaf; };   // c[i-1] = C(n, i), k[i-1] = left share (0: same as i-1)
    std::vector<DpRow> row;
    auto skipSingles = [&](int32_t b) {      // an inner node with one child stands for that child
        for (;;) {
            const BinNode& n = bin.nodes[b];
            if (n.left >= 0 && n.right >= 0) return b;
            if (n.left < 0 && n.right < 0) return b;
            b = n.left >= 0 ? n.left : n.right;
        }
    };
    if (dp) {
        float cTri = 0.7f;
        if (const char* e = std::getenv("SPICA_BVH_CTRI")) cTri = (float)std::atof(e);
        const float cNode = 1.0f, inf = 3.0e38f;
        row.resize(bin.nodes.size());
        auto computeRow = [&](int32_t b) {          // the rows of b's children are final
            const BinNode& n = bin.nodes[b];
            DpRow& R = row[b];
            const float A = (float)(nodeArea(b) / rootArea);
            const float cLeaf = (n.count <= max_leaf || (n.left < 0 && n.right < 0)) ? A * (float)n.count * cTri : inf;
            if (n.left < 0 && n.right < 0) {
                for (int i = 0; i < 8; i++) { R.c[i] = cLeaf; R.k[i] = 0; }
                R.k8 = 0; R.leaf = 1;
                return;
            }
            if (n.left < 0 || n.right < 0) { R = row[n.left >= 0 ? n.left : n.right]; return; }   // pass-through
            const DpRow& L = row[n.left]; const DpRow& Rr = row[n.right];
            float D[9]; uint8_t Dk[9];
            for (int j = 2; j <= 8; j++) {
                float best = inf; uint8_t bk = 1;
                for (int k = 1; k < j; k++) {
                    const float v = L.c[k - 1] + Rr.c[j - k - 1];
                    if (v < best) { best = v; bk = (uint8_t)k; }
                }
                D[j] = best; Dk[j] = bk;
            }
            const float cInner = A * cNode + D[8];
            R.k8 = Dk[8];
            R.leaf = cLeaf <= cInner ? 1 : 0;
            R.c[0] = R.leaf ? cLeaf : cInner; R.k[0] = 0;
            for (int i = 2; i <= 7; i++) {
                if (D[i] < R.c[i - 2]) { R.c[i - 1] = D[i]; R.k[i - 1] = Dk[i]; }
                else { R.c[i - 1] = R.c[i - 2]; R.k[i - 1] = 0; }
            }
            R.c[7] = R.c[6]; R.k[7] = 0;
        };
        // children before parents: the reverse of a pre-order walk of the subtree
        auto solveSerial = [&](int32_t root) {
            std::vector<int32_t> stackv, orderv;
            stackv.push_back(root);
            while (!stackv.empty()) {
                const int32_t b = stackv.back(); stackv.pop_back();
                orderv.push_back(b);
                if (bin.nodes[b].left >= 0) stackv.push_back(bin.nodes[b].left);
                if (bin.nodes[b].right >= 0) stackv.push_back(bin.nodes[b].right);
            }
            for (size_t oi = orderv.size(); oi-- > 0;) computeRow(orderv[oi]);
        };
        // the big subtrees near the root on their own threads (rows of different subtrees never alias)
        std::function<void(int32_t, int)> solve = [&](int32_t b, int depth) {
            const BinNode& n = bin.nodes[b];
            if (n.count > 65536 && depth < 5 && n.left >= 0 && n.right >= 0) {
                auto fut = std::async(std::launch::async, [&solve, &n, depth]() { solve(n.left, depth + 1); });
                solve(n.right, depth + 1);
                fut.get();
                computeRow(b);
            } else {
                solveSerial(b);
            }
        };
        solve(bin.root, 0);
    }
    std::vector<uint8_t> dpLeaf;           // per binary node: chosen as a leaf child by the optimal cut
    if (dp) dpLeaf.assign(bin.nodes.size(), 0);
    // children of the forest C(b, i), appended to ch[]
    struct Frame { int32_t b; int i; };
    auto collectDp = [&](int32_t b0, int i0, int32_t* ch, int& nch) {
        Frame fs[64]; int nf = 0;
        fs[nf++] = {b0, i0};
        while (nf > 0) {
            Frame f = fs[--nf];
            const int32_t b = skipSingles(f.b);
            const BinNode& n = bin.nodes[b];
            int i = f.i;
            if (n.left < 0 && n.right < 0) { dpLeaf[b] = 1; ch[nch++] = b; continue; }
            while (i > 1 && row[b].k[i - 1] == 0) i--;
            if (i == 1) { dpLeaf[b] = row[b].leaf; ch[nch++] = b; continue; }
            const int k = row[b].k[i - 1];
            fs[nf++] = {n.right, i - k};       // (popped after the left part: children stay in left-to-right order)
            fs[nf++] = {n.left, k};
        }
    };

    struct Work { int32_t bnode; int32_t depth; };
    std::vector<Work> work;               // work[i] describes wide node i (BFS order)
    work.push_back({bin.root, 1});
    out->nodes.reserve((size_t)(n_tris / 2 + 16));

    for (size_t wi = 0; wi < work.size(); wi++) {
        const Work w = work[wi];
        out->max_depth = std::max(out->max_depth, w.depth);
        // ---- gather up to 8 children by greedy surface-area expansion
        int32_t ch[8]; int nch = 0;
        const BinNode& bn = bin.nodes[w.bnode];
        if (wi == 0 && leafLike(w.bnode)) {
            ch[nch++] = w.bnode;          // degenerate root: a single leaf child
            if (dp) dpLeaf[w.bnode] = 1;
        } else if (dp) {
            const int32_t b = skipSingles(w.bnode);
            const BinNode& n = bin.nodes[b];
            if (n.left < 0 && n.right < 0) { dpLeaf[b] = 1; ch[nch++] = b; }
            else { const int k = row[b].k8; collectDp(n.left, k, ch, nch); collectDp(n.right, 8 - k, ch, nch); }
        } else {
            if (bn.left >= 0) ch[nch++] = bn.left;
            if (bn.right >= 0) ch[nch++] = bn.right;
            for (;;) {
                // pass-through of single-child inner nodes, and expansion
                int best = -1; double bestA = -1.0;
                for (int i = 0; i < nch; i++) {
                    if (leafLike(ch[i])) continue;
                    const BinNode& c = bin.nodes[ch[i]];
                    const int kids = (c.left >= 0) + (c.right >= 0);
                    if (kids == 1) { best = i; bestA = DBL_MAX; break; }
                    if (nch >= 8) continue;
                    const double a = nodeArea(ch[i]);
                    if (a > bestA) { bestA = a; best = i; }
                }
                if (best < 0) break;
                const BinNode& c = bin.nodes[ch[best]];
                if (c.left >= 0 && c.right >= 0) {
                    if (nch >= 8) break;
                    ch[best] = c.left; ch[nch++] = c.right;
                } else {
                    ch[best] = c.left >= 0 ? c.left : c.right;
                }
            }
        }

        // ---- slot assignment: child whose offset from the node centre aligns best with the
        // slot diagonal D_s = (+-1,+-1,+-1) (bit set => +) ; greedy on the largest dot product.
        double ctr[3];
        for (int k = 0; k < 3; k++) ctr[k] = 0.5 * (bn.lo[k] + bn.hi[k]);
        int slotOf[8]; bool slotUsed[8] = {false}; bool chDone[8] = {false};
        for (int it = 0; it < nch; it++) {
            double bestV = -DBL_MAX; int bc = -1, bs = -1;
            for (int c = 0; c < nch; c++) {
                if (chDone[c]) continue;
                const BinNode& cn = bin.nodes[ch[c]];
                double off[3];
                for (int k = 0; k < 3; k++) off[k] = 0.5 * (cn.lo[k] + cn.hi[k]) - ctr[k];
                for (int s = 0; s < 8; s++) {
                    if (slotUsed[s]) continue;
                    const double v = ((s & 1) ? off[0] : -off[0]) + ((s & 2) ? off[1] : -off[1]) + ((s & 4) ? off[2] : -off[2]);
                    if (v > bestV) { bestV = v; bc = c; bs = s; }
                }
            }
            chDone[bc] = true; slotUsed[bs] = true; slotOf[bc] = bs;
        }
        int32_t slotChild[8];
        for (int s = 0; s < 8; s++) slotChild[s] = -1;
        for (int c = 0; c < nch; c++) slotChild[slotOf[c]] = ch[c];

        // ---- quantisation grid
        WideNode wn;
        std::memset(&wn, 0, sizeof(wn));
        double plo[3], ext[3];
        for (int k = 0; k < 3; k++) {
            double mn = DBL_MAX, mx = -DBL_MAX;
            for (int c = 0; c < nch; c++) {
                mn = std::min(mn, bin.nodes[ch[c]].lo[k] - inflate);
                mx = std::max(mx, bin.nodes[ch[c]].hi[k] + inflate);
            }
            wn.p[k] = floatDown(mn);
            plo[k] = (double)wn.p[k];
            ext[k] = mx - plo[k];
        }
        int ex[3];
        for (int k = 0; k < 3; k++) {
            // smallest e with 2^e >= ext / 255, from the exponent of the quotient (exact; the device encoder uses the same rule)
            int e;
            { int x; const double m = std::frexp(std::max(ext[k], 1e-300) / 255.0, &x); e = (m == 0.5) ? x - 1 : x; }
            while (std::ldexp(255.0, e) < ext[k]) e++;
            if (e < -100) e = -100;
            if (e > 100) { *err = "quantisation exponent out of range"; return false; }
            ex[k] = e;
        }
        // quantise, bumping the exponent if a plane would not fit in 8 bits
        for (int k = 0; k < 3; k++) {
            for (;;) {
                bool ok = true;
                const double step = std::ldexp(1.0, ex[k]);
                for (int s = 0; s < 8 && ok; s++) {
                    if (slotChild[s] < 0) continue;
                    const double lo = bin.nodes[slotChild[s]].lo[k] - inflate, hi = bin.nodes[slotChild[s]].hi[k] + inflate;
                    double ql = std::floor((lo - plo[k]) / step);
                    double qh = std::ceil((hi - plo[k]) / step);
                    while (plo[k] + ql * step > lo) ql -= 1.0;
                    while (plo[k] + qh * step < hi) qh += 1.0;
                    if (ql < 0.0) ql = 0.0;
                    if (qh > 255.0) { ok = false; break; }
                    wn.qlo[k][s] = (uint8_t)ql;
                    wn.qhi[k][s] = (uint8_t)qh;
                }
                if (ok) break;
                ex[k]++;
            }
            wn.e[k] = (uint8_t)(ex[k] + 127);
        }

        // ---- children: inner ones get contiguous wide indices in slot order; leaves get triangles
        wn.child_base = (uint32_t)work.size();
        wn.tri_base = (uint32_t)triCursor;
        int triOff = 0;
        for (int s = 0; s < 8; s++) {
            const int32_t c = slotChild[s];
            if (c < 0) { wn.meta[s] = 0; continue; }
            const double relA = nodeArea(c) / rootArea;
            if (dp ? dpLeaf[c] != 0 : leafLike(c)) {
                const BinNode& cn = bin.nodes[c];
                const int n = cn.count;     // 1..3
                const uint8_t unary = (uint8_t)((1u << n) - 1u);
                wn.meta[s] = (uint8_t)((unary << 5) | (uint8_t)triOff);
                // Within a leaf the triangles go in ascending primitive index: the builders agree on the SET of a leaf, not on
                // the order their partition passes leave it in, and the records should not depend on that.  (An imported
                // tree keeps its leaf order: there the position is the tie-break rank.)
                int32_t ids[3];
                for (int i = 0; i < n; i++) ids[i] = bin.order[cn.first + i];
                if (!bin.imported) std::sort(ids, ids + n);
                for (int i = 0; i < n; i++) emitTri(ids[i]);
                triOff += n;
                out->sah_cost += relA * n;
            } else {
                wn.imask |= (uint8_t)(1u << s);
                wn.meta[s] = (uint8_t)((1u << 5) | (24 + s));
                work.push_back({c, w.depth + 1});
                out->sah_cost += relA;
            }
        }
        out->nodes.push_back(wn);
    }
    if (triCursor != n_tris) { *err = "internal: collapse did not emit every triangle"; return false; }
    out->n_wide = (int64_t)out->nodes.size();
    out->tri_bytes_device = (int64_t)out->tris.size();
    if (out->max_depth > kStackCapacity - 2) {
        *err = "wide tree deeper than the traversal stack (" + std::to_string(out->max_depth) + ")";
        return false;
    }
    return true;
}

}  // namespace spb
