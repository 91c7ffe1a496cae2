// K6: GPU LBVH build (builder = 1 of spb_bvh_build). Replaces the reference's single-threaded
// top-down builder BVHAccel::constructRec (accelerators/bvh.cc:150-237) with
//   (1) per-triangle bounds + 63-bit Morton codes of the centroids      mortonKernel
//   (2) radix sort of (code, primitive) pairs                            cub::DeviceRadixSort (library sort)
//   (3) Karras' binary radix tree: one thread per inner node             hierarchyKernel
//   (4) bottom-up AABB refit with one atomic flag per inner node         refitKernel
// The result is a BinaryBVH (same structure the host SAH builder produces: 1 primitive per leaf,
// primitives in left-to-right leaf order), which the common 8-wide collapse / quantisation then
// encodes. Bounds are exact doubles of the triangle vertices, so culling stays conservative and
// query results are identical to any other valid tree (tests/test_trace_gpu.py).
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cfloat>

#include "context.h"

namespace spb {

struct DevBinNode {          // mirrors BinNode (bvh_host.h)
    double lo[3], hi[3];
    int32_t left, right, first, count;
};
static_assert(sizeof(DevBinNode) == sizeof(BinNode), "BinNode layout");

__device__ __forceinline__ unsigned long long expandBits21(unsigned long long v) {
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void mortonKernel(const double* __restrict__ verts, int64_t n, double3 lo, double3 inv, unsigned long long* __restrict__ keys,
                             int32_t* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* v = verts + i * 9;
    double c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double mn = fmin(v[k], fmin(v[3 + k], v[6 + k])), mx = fmax(v[k], fmax(v[3 + k], v[6 + k]));
        c[k] = 0.5 * (mn + mx);
    }
    const double s = 2097151.0;
    const unsigned long long x = (unsigned long long)fmin(fmax((c[0] - lo.x) * inv.x * s, 0.0), s);
    const unsigned long long y = (unsigned long long)fmin(fmax((c[1] - lo.y) * inv.y * s, 0.0), s);
    const unsigned long long z = (unsigned long long)fmin(fmax((c[2] - lo.z) * inv.z * s, 0.0), s);
    keys[i] = (expandBits21(x) << 2) | (expandBits21(y) << 1) | expandBits21(z);
    vals[i] = (int32_t)i;
}

// common-prefix length of keys i and j (ties broken by the index), -1 outside the array
__device__ __forceinline__ int delta(const unsigned long long* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

// Karras 2012, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees".
// Inner nodes 0..n-2, leaves n-1..2n-2 (leaf k = sorted position k). parent[] feeds the refit.
__global__ void hierarchyKernel(const unsigned long long* __restrict__ keys, int n, DevBinNode* __restrict__ nodes, int32_t* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const int left = (first == gamma) ? (n - 1 + gamma) : gamma;
    const int right = (last == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    DevBinNode& nd = nodes[i];
    nd.left = left; nd.right = right; nd.first = first; nd.count = last - first + 1;
    parent[left] = i; parent[right] = i;
    if (i == 0) parent[0] = -1;
}

__global__ void refitKernel(const double* __restrict__ verts, const int32_t* __restrict__ order, int n, DevBinNode* __restrict__ nodes,
                            const int32_t* __restrict__ parent, unsigned int* __restrict__ flags) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double* v = verts + (int64_t)order[k] * 9;
    DevBinNode& leaf = nodes[n - 1 + k];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        leaf.lo[a] = fmin(v[a], fmin(v[3 + a], v[6 + a]));
        leaf.hi[a] = fmax(v[a], fmax(v[3 + a], v[6 + a]));
    }
    leaf.left = -1; leaf.right = -1; leaf.first = k; leaf.count = 1;
    __threadfence();
    int p = parent[n - 1 + k];
    while (p >= 0) {
        if (atomicAdd(&flags[p], 1u) == 0u) return;     // the second child to arrive carries on
        __threadfence();
        DevBinNode& nd = nodes[p];
        const volatile DevBinNode& L = nodes[nd.left];
        const volatile DevBinNode& R = nodes[nd.right];
#pragma unroll
        for (int a = 0; a < 3; a++) { nd.lo[a] = fmin(L.lo[a], R.lo[a]); nd.hi[a] = fmax(L.hi[a], R.hi[a]); }
        __threadfence();
        p = parent[p];
    }
}

// Builds the binary LBVH of ctx->verts on the context's GPU into *out. Returns SPB_OK or an error.
int buildLbvhDevice(spb_ctx* ctx, BinaryBVH* out) {
    const int64_t n64 = ctx->n_tris;
    out->nodes.clear(); out->order.clear(); out->root = -1; out->imported = false;
    if (n64 <= 0) return SPB_OK;
    if (n64 == 1) {
        BinNode nd;
        const double* v = ctx->geo->verts.data();
        for (int a = 0; a < 3; a++) { nd.lo[a] = std::min(v[a], std::min(v[3 + a], v[6 + a])); nd.hi[a] = std::max(v[a], std::max(v[3 + a], v[6 + a])); }
        nd.left = nd.right = -1; nd.first = 0; nd.count = 1;
        out->nodes.push_back(nd); out->order.push_back(0); out->root = 0;
        return SPB_OK;
    }
    const int n = (int)n64;
    // centroid bounds on the host (one pass over data that is already host-resident)
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int64_t i = 0; i < n64; i++) {
        const double* v = ctx->geo->verts.data() + i * 9;
        for (int a = 0; a < 3; a++) {
            const double c = 0.5 * (std::min(v[a], std::min(v[3 + a], v[6 + a])) + std::max(v[a], std::max(v[3 + a], v[6 + a])));
            lo[a] = std::min(lo[a], c); hi[a] = std::max(hi[a], c);
        }
    }
    double3 dlo = make_double3(lo[0], lo[1], lo[2]);
    double3 inv = make_double3(hi[0] > lo[0] ? 1.0 / (hi[0] - lo[0]) : 0.0, hi[1] > lo[1] ? 1.0 / (hi[1] - lo[1]) : 0.0,
                               hi[2] > lo[2] ? 1.0 / (hi[2] - lo[2]) : 0.0);
    cudaStream_t st = ctx->stream;
    double* d_verts = nullptr; unsigned long long *d_k0 = nullptr, *d_k1 = nullptr; int32_t *d_v0 = nullptr, *d_v1 = nullptr, *d_parent = nullptr;
    DevBinNode* d_nodes = nullptr; unsigned int* d_flags = nullptr; void* d_tmp = nullptr;
    auto freeAll = [&]() {
        cudaFree(d_verts); cudaFree(d_k0); cudaFree(d_k1); cudaFree(d_v0); cudaFree(d_v1); cudaFree(d_parent);
        cudaFree(d_nodes); cudaFree(d_flags); cudaFree(d_tmp);
    };
#define LB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { freeAll(); cudaOk(ctx, e_, #call); return e_ == cudaErrorMemoryAllocation ? SPB_ERR_OOM : SPB_ERR_CUDA; } } while (0)
    LB(cudaMalloc(&d_verts, (size_t)n * 9 * sizeof(double)));
    LB(cudaMalloc(&d_k0, (size_t)n * 8)); LB(cudaMalloc(&d_k1, (size_t)n * 8));
    LB(cudaMalloc(&d_v0, (size_t)n * 4)); LB(cudaMalloc(&d_v1, (size_t)n * 4));
    LB(cudaMalloc(&d_parent, (size_t)(2 * n - 1) * 4));
    LB(cudaMalloc(&d_nodes, (size_t)(2 * n - 1) * sizeof(DevBinNode)));
    LB(cudaMalloc(&d_flags, (size_t)n * 4));
    LB(cudaMemcpyAsync(d_verts, ctx->geo->verts.data(), (size_t)n * 9 * sizeof(double), cudaMemcpyHostToDevice, st));
    const int B = 256;
    mortonKernel<<<(n + B - 1) / B, B, 0, st>>>(d_verts, n, dlo, inv, d_k0, d_v0);
    size_t tmpBytes = 0;
    LB(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, d_k0, d_k1, d_v0, d_v1, n, 0, 63, st));
    LB(cudaMalloc(&d_tmp, tmpBytes));
    LB(cub::DeviceRadixSort::SortPairs(d_tmp, tmpBytes, d_k0, d_k1, d_v0, d_v1, n, 0, 63, st));
    LB(cudaMemsetAsync(d_flags, 0, (size_t)n * 4, st));
    hierarchyKernel<<<(n - 1 + B - 1) / B, B, 0, st>>>(d_k1, n, d_nodes, d_parent);
    refitKernel<<<(n + B - 1) / B, B, 0, st>>>(d_verts, d_v1, n, d_nodes, d_parent, d_flags);
    LB(cudaGetLastError());
    ctx->kernel_launches += 5;
    out->nodes.resize((size_t)(2 * n - 1));
    out->order.resize((size_t)n);
    LB(cudaMemcpyAsync(out->nodes.data(), d_nodes, (size_t)(2 * n - 1) * sizeof(DevBinNode), cudaMemcpyDeviceToHost, st));
    LB(cudaMemcpyAsync(out->order.data(), d_v1, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    LB(cudaStreamSynchronize(st));
#undef LB
    freeAll();
    out->root = 0;
    return SPB_OK;
}

}  // namespace spb
