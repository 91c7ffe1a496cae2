// Device-resident acceleration-structure layout (shared by the host encoder, the CUDA kernels and
// the CPU emulation used by the unit tests).
//
// 8-wide compressed BVH: 80-byte nodes read as five 128-bit loads.  Child boxes are quantised to
// 8 bits per plane on a per-node power-of-two grid and are CONSERVATIVE (they contain the exact
// fp64 child bounds inflated by `inflate`), because they only cull: the winner of a query is
// always decided by the fp64 Moeller-Trumbore test, which restates core/triangle.cc:98-117 of the
// reference bit for bit.
#pragma once
#include <stdint.h>

namespace spb {

struct alignas(16) WideNode {           // 80 B
    float    p[3];                      // grid origin (<= every child lo)
    uint8_t  e[3];                      // biased exponents: plane = p + q * 2^(e-127)
    uint8_t  imask;                     // bit s set: slot s holds an inner node
    uint32_t child_base;                // index of the first inner child (children are contiguous, slot order)
    uint32_t tri_base;                  // index of the first triangle of this node's leaves
    uint8_t  meta[8];                   // per slot: 0 = empty; inner: (1<<5)|(24+slot); leaf: (unary(n)<<5)|offset
    uint8_t  qlo[3][8];                 // [axis][slot]
    uint8_t  qhi[3][8];
};
static_assert(sizeof(WideNode) == 80, "WideNode must be 80 bytes");

// Triangle records, stored in leaf order.  id = the caller's primitive index; rank = tie-break key
// (smaller wins on exactly equal t).
struct alignas(16) TriF32 {             // 48 B: vertices exactly representable in float32
    float v0[3]; int32_t id;
    float v1[3]; int32_t rank;
    float v2[3]; int32_t pad;
};
struct alignas(16) TriF64 {             // 80 B
    double v[9];
    int32_t id, rank;
};
static_assert(sizeof(TriF32) == 48 && sizeof(TriF64) == 80, "triangle record sizes");

enum { kStackCapacity = 40 };

struct SceneParams {                    // small POD passed to kernels by value
    const WideNode* nodes;
    const void*     tris;               // TriF32* or TriF64*
    double wlo[3], whi[3];              // inflated world box
    float  inflate;
    float  max_coord;                   // max |coordinate| of the inflated world box (error bounds of the float32 pre-test)
    int32_t tri_format;                 // 0 = TriF32, 1 = TriF64
    int32_t n_tris;
    int32_t empty;                      // 1: no geometry
    int32_t max_depth;                  // depth of the wide tree = the most stack entries a ray can hold
    uint32_t f32_one;                   // 0x3f800000, as a parameter-block operand of the node test's PRMT (trace_core.h, byteMant)
    // What the float32 pre-test reads (trace_core.h, triPretestMayHit): for TriF32 scenes the records themselves and
    // pre_round = 0; for TriF64 scenes a TriF32 copy whose vertices are the doubles rounded to nearest float32, and
    // pre_round = 2^-23 * max_coord >= twice the largest rounding step, which the pre-test adds to its error bounds.
    const TriF32* pre_tris;
    float    pre_round;
};

}  // namespace spb
