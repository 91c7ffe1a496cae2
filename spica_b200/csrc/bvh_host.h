// Host side of the acceleration structure: binary SAH build (or import of a reference-built
// binary tree), greedy collapse to 8-wide, octant slot assignment, conservative quantisation.
// Replaces BVHAccel::construct / constructRec (reference accelerators/bvh.cc:139-237) and takes
// the idea - not the code - of its 4-wide collapse (accelerators/bvh.cc:248-313).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/spica_b200.h"
#include "bvh_layout.h"

namespace spb {

struct BinNode {
    double lo[3], hi[3];
    int32_t left, right;    // -1 for leaves
    int32_t first, count;   // range in BinaryBVH::order covered by this subtree
};

struct BinaryBVH {
    std::vector<BinNode> nodes;
    std::vector<int32_t> order;   // primitive indices, left-to-right leaf order
    int32_t root = -1;
    bool imported = false;
};

struct HostBVH {
    std::vector<WideNode> nodes;
    std::vector<uint8_t>  tris;       // TriF32 or TriF64 records in leaf order
    int32_t tri_format = 0;
    int64_t n_tris = 0;
    double  wlo[3] = {0, 0, 0}, whi[3] = {0, 0, 0};   // exact world bound
    double  inflate = 0.0;
    int32_t max_depth = 0;
    double  sah_cost = 0.0;
    int64_t n_binary_nodes = 0;
    int64_t n_wide = 0;              // wide nodes / triangle bytes on the device (the device builder leaves `nodes` and `tris` empty)
    int64_t tri_bytes_device = 0;
};

// Top-down binned SAH over all three axes, 1 primitive per leaf (the collapse decides leaf sizes).
void build_binary_sah(const double* verts, int64_t n, int bins, BinaryBVH* out);

// Adopt a reference-built topology (a dump of the reference BVHAccel nodes). Bounds are recomputed from the
// triangles so that culling never depends on imported numbers.
bool import_binary(const spb_import_node* nodes, int64_t n_nodes, int32_t root, const double* verts,
                   int64_t n_tris, BinaryBVH* out, std::string* err);

bool encode_wide(const BinaryBVH& bin, const double* verts, int64_t n_tris, int max_leaf_tris,
                 HostBVH* out, std::string* err);

}  // namespace spb
