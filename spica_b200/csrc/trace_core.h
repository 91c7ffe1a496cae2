// Per-ray traversal of the 8-wide compressed BVH.  __host__ __device__ so that the very same code
// is (a) inlined into the sm_100a kernels (trace.cu, integrator.cu) and (b) compiled by g++ into
// the unit-test emulation (tests/emul), where the node decoding, stack handling, ordering and the
// exact triangle test are checked against the oracle without a GPU.  The emulation is test
// infrastructure; the product never runs this on the CPU.
//
// Two arithmetic domains:
//   culling  : float32, conservative (quantised boxes inflated by SceneParams::inflate, ray clipped
//              to the world box so every float32 magnitude is bounded by the scene size);
//   deciding : float64 Moeller-Trumbore in the reference's operation order with no FMA
//              contraction (core/triangle.cc:98-117), on the reference's own ray representation
//              (core/ray.cc:11-19).  Only this domain determines (prim, t, u, v).
#pragma once
#include <math.h>
#include <stdint.h>

#include "bvh_layout.h"

#ifndef SPB_NODE_I2F
#define SPB_NODE_I2F 0     // 1: the r01 node test (integer-to-float conversions on the XU pipe), kept for A/B measurement builds
#endif

#if defined(__CUDACC__)
#define SPB_HD __host__ __device__ __forceinline__
#else
#define SPB_HD inline
#endif

namespace spb {

// ---- exact double arithmetic: never contracted into FMA --------------------------------------
#if defined(__CUDA_ARCH__)
SPB_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
SPB_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
SPB_HD double dsub(double a, double b) { return __dadd_rn(a, -b); }
SPB_HD double drcp(double a) { return __drcp_rn(a); }            // == 1.0 / a, correctly rounded
SPB_HD double dsqrt(double a) { return __dsqrt_rn(a); }
SPB_HD float  ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
SPB_HD float  frcpApprox(float a) { return __fdividef(1.0f, a); }   // MUFU.RCP, <= 2 ulp for |a| < 2^126
SPB_HD int    clz32(uint32_t x) { return __clz((int)x); }
SPB_HD int    ctz32(uint32_t x) { return __ffs((int)x) - 1; }
SPB_HD int    popc32(uint32_t x) { return __popc(x); }
#else
// host emulation: compiled with -ffp-contract=off on x86-64 (no FMA unless asked)
SPB_HD double dmul(double a, double b) { return a * b; }
SPB_HD double dadd(double a, double b) { return a + b; }
SPB_HD double dsub(double a, double b) { return a - b; }
SPB_HD double drcp(double a) { return 1.0 / a; }
SPB_HD double dsqrt(double a) { return sqrt(a); }
SPB_HD float  ffma(float a, float b, float c) { return fmaf(a, b, c); }
SPB_HD float  frcpApprox(float a) { return 1.0f / a; }
SPB_HD int    clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
SPB_HD int    ctz32(uint32_t x) { return x ? __builtin_ctz(x) : -1; }
SPB_HD int    popc32(uint32_t x) { return __builtin_popcount(x); }
#endif

struct U4 { uint32_t x, y, z, w; };
struct U2 { uint32_t x, y; };

// x << (s mod 32): one SHF.L.W on the device, no masking of the shift amount
#if defined(__CUDA_ARCH__)
SPB_HD uint32_t shlWrap(uint32_t x, uint32_t s) { return __funnelshift_l(0u, x, s); }
#else
SPB_HD uint32_t shlWrap(uint32_t x, uint32_t s) { return x << (s & 31u); }
#endif

SPB_HD float asFloat(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; __builtin_memcpy(&f, &u, 4); return f;
#endif
}
SPB_HD uint32_t byteOf(uint32_t w, int i) { return (w >> (8 * i)) & 0xffu; }

#if defined(__CUDA_ARCH__)
SPB_HD U4 ldg4(const void* p) { const uint4 v = __ldg((const uint4*)p); return U4{v.x, v.y, v.z, v.w}; }
#else
SPB_HD U4 ldg4(const void* p) { U4 v; __builtin_memcpy(&v, p, 16); return v; }
#endif

const double kRefEps = 1.0e-12;     // reference core/common.h:55
const double kRefInfty = 1.0e32;    // reference core/common.h:54

struct TraceCounters { unsigned long long nodes, tris, exact; };   // node visits, triangle candidates, exact (double) tests

struct RayState {
    // deciding domain
    double ox, oy, oz, dx, dy, dz;
    double best_t;                  // the reference's shrinking ray.maxDist (core/primitive.cc:52)
    double t_shift;                 // culling origin = o + t_shift * d
    // culling domain
    float cox, coy, coz, idx, idy, idz, ctmax;
    float fdx, fdy, fdz;            // float32 direction (pre-test only)
    uint32_t oct_inv;
    // result
    int32_t best_prim, best_rank;
    float best_u, best_v;
};

SPB_HD float cullTmax(const RayState& r, double t) {
    // distance along the culling ray, rounded up with slack
    return (float)((t - r.t_shift) * 1.000002) + 1e-30f;
}

// Applies the reference's Ray constructor (core/ray.cc:11-19: dir = d * (1.0/|d|)) and derives the
// float32 culling ray.  Returns false when the ray cannot hit anything (zero direction, or it
// misses the inflated world box / the box lies beyond tmax).
SPB_HD bool rayBegin(const SceneParams& sp, double ox, double oy, double oz, double dx, double dy,
                     double dz, double tmax, RayState& r) {
    r.best_prim = -1; r.best_rank = 0x7fffffff; r.best_u = 0.f; r.best_v = 0.f;
    r.best_t = tmax;
    r.ox = ox; r.oy = oy; r.oz = oz;
    const double sq = dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
    const double nrm = dsqrt(sq);
    if (!(nrm > 0.0) || sp.empty) { r.dx = r.dy = r.dz = 0.0; return false; }
    const double s = drcp(nrm);
    r.dx = dmul(dx, s); r.dy = dmul(dy, s); r.dz = dmul(dz, s);

    // clip against the inflated world box in double (not part of the parity arithmetic)
    double t0 = 0.0, t1 = tmax;
    const double o[3] = {ox, oy, oz}, d[3] = {r.dx, r.dy, r.dz};
    for (int k = 0; k < 3; k++) {
        if (d[k] == 0.0) {
            if (o[k] < sp.wlo[k] || o[k] > sp.whi[k]) return false;
        } else if (fabs(d[k]) > 1e-30) {                 // below that the slab gives no usable bound: skip it
            // float32 reciprocal (relative error < 2^-21 incl. the rounding of d): an IEEE double
            // division costs ~15 instructions per axis and this test only prunes, so it is widened instead
            const double inv = (double)frcpApprox((float)d[k]);
            double ta = (sp.wlo[k] - o[k]) * inv, tb = (sp.whi[k] - o[k]) * inv;
            if (ta > tb) { const double tt = ta; ta = tb; tb = tt; }
            ta -= fabs(ta) * 1e-6; tb += fabs(tb) * 1e-6;
            if (ta > t0) t0 = ta;
            if (tb < t1) t1 = tb;
        }
    }
    if (!(t0 <= t1)) return false;
    r.t_shift = t0 > 0.0 ? t0 * 0.999999 : 0.0;
    r.cox = (float)(ox + r.t_shift * r.dx);
    r.coy = (float)(oy + r.t_shift * r.dy);
    r.coz = (float)(oz + r.t_shift * r.dz);
    const float fdx = (float)r.dx, fdy = (float)r.dy, fdz = (float)r.dz;
    r.fdx = fdx; r.fdy = fdy; r.fdz = fdz;
    const float tiny = 8.6736174e-19f;  // 2^-60: keeps 1/d finite; far below the culling slack
    r.idx = 1.0f / (fabsf(fdx) < tiny ? (fdx < 0.f ? -tiny : tiny) : fdx);
    r.idy = 1.0f / (fabsf(fdy) < tiny ? (fdy < 0.f ? -tiny : tiny) : fdy);
    r.idz = 1.0f / (fabsf(fdz) < tiny ? (fdz < 0.f ? -tiny : tiny) : fdz);
    // bit set where the direction is non-negative (slot ^ oct_inv == 7 for the nearest corner)
    r.oct_inv = (fdx < 0.f ? 0u : 1u) | (fdy < 0.f ? 0u : 2u) | (fdz < 0.f ? 0u : 4u);
    r.ctmax = cullTmax(r, t1 < tmax ? t1 : tmax);
    return true;
}

// fp64 Moeller-Trumbore, same operations in the same order as core/triangle.cc:100-117.
SPB_HD bool triTestExact(const RayState& r, const double p0[3], const double p1[3], const double p2[3],
                         double* tOut, double* uOut, double* vOut) {
    const double e1x = dsub(p1[0], p0[0]), e1y = dsub(p1[1], p0[1]), e1z = dsub(p1[2], p0[2]);
    const double e2x = dsub(p2[0], p0[0]), e2y = dsub(p2[1], p0[1]), e2z = dsub(p2[2], p0[2]);
    const double px = dsub(dmul(r.dy, e2z), dmul(r.dz, e2y));
    const double py = dsub(dmul(r.dz, e2x), dmul(r.dx, e2z));
    const double pz = dsub(dmul(r.dx, e2y), dmul(r.dy, e2x));
    const double det = dadd(dadd(dmul(e1x, px), dmul(e1y, py)), dmul(e1z, pz));
    if (det > -kRefEps && det < kRefEps) return false;
    const double invdet = drcp(det);
    const double tx = dsub(r.ox, p0[0]), ty = dsub(r.oy, p0[1]), tz = dsub(r.oz, p0[2]);
    const double u = dmul(dadd(dadd(dmul(tx, px), dmul(ty, py)), dmul(tz, pz)), invdet);
    if (u < 0.0 || u > 1.0) return false;
    const double qx = dsub(dmul(ty, e1z), dmul(tz, e1y));
    const double qy = dsub(dmul(tz, e1x), dmul(tx, e1z));
    const double qz = dsub(dmul(tx, e1y), dmul(ty, e1x));
    const double v = dmul(dadd(dadd(dmul(r.dx, qx), dmul(r.dy, qy)), dmul(r.dz, qz)), invdet);
    if (v < 0.0 || dadd(u, v) > 1.0) return false;
    const double t = dmul(dadd(dadd(dmul(e2x, qx), dmul(e2y, qy)), dmul(e2z, qz)), invdet);
    if (t <= kRefEps || t > r.best_t) return false;
    *tOut = t; *uOut = u; *vOut = v;
    return true;
}


// Conservative float32 pre-test for float32-exact triangles (TriF32). It may only say "certainly
// rejected by the exact test"; everything else goes on to triTestExact, which alone decides.
// Moeller-Trumbore on the culling ray (origin c = fl32(o + t_shift d), direction fl32(d)) with
// forward error bounds. With E1 = max|e1_k|, E2 = max|e2_k|, TV = max|tv_k|, M = max scene
// coordinate and u = 2^-24, the float32 results differ from the real-number values by at most
//   |d det| <=  72 u E1 E2            -> 2^-17 E1 E2
//   |d U|   <=  90 u E2 (M + TV)      -> 2^-17 E2 (M + TV)
//   |d V|   <=  72 u E1 (M + TV)      -> 2^-17 E1 (M + TV)
//   |d T|   <=  72 u E1 E2 (M + TV)   -> 2^-17 E1 E2 (M + TV)
// (input roundings: |d c_k| <= u (M + TV), |d dir_k| <= u, edges relative u; then the running
// error of each 3-term product sum; derivation in DESIGN.md "float32 pre-test"). Every comparison is
// written so that NaN / Inf fall through to the exact test.
struct CullRay { float cox, coy, coz, fdx, fdy, fdz, ctmax; };   // what the pre-test reads of a ray (7 floats: cheap to hand to another lane)

//
// Rounded triangles (TriF64 scenes): a, b, c are the double vertices rounded to float32, each coordinate off by at
// most u M.  With rnd = 2^-23 M (twice that, which also swallows the second-order terms), exact arithmetic on the
// rounded triangle differs from exact arithmetic on the true one by at most (|d_k| <= 1, first-order perturbation
// of each triple product: edges move by <= rnd, tv by <= rnd / 2)
//   |d det| <= 6 rnd (E1 + E2)            |d U| <= rnd (3 E2 + 6 TV)
//   |d V|   <= rnd (3 E1 + 6 TV)          |d T| <= rnd (6 TV E1 + 3 E1 E2 + 6 E2 TV)
// and the pre-test adds twice these to the bounds above.  rnd = 0 (a literal at the TriF32 call sites) folds them away.
SPB_HD bool triPretestMayHit(const CullRay& r, float maxCoord, const U4& a, const U4& b, const U4& c, const float rnd = 0.f) {
    const float p0x = asFloat(a.x), p0y = asFloat(a.y), p0z = asFloat(a.z);
    const float e1x = asFloat(b.x) - p0x, e1y = asFloat(b.y) - p0y, e1z = asFloat(b.z) - p0z;
    const float e2x = asFloat(c.x) - p0x, e2y = asFloat(c.y) - p0y, e2z = asFloat(c.z) - p0z;
    const float px = r.fdy * e2z - r.fdz * e2y, py = r.fdz * e2x - r.fdx * e2z, pz = r.fdx * e2y - r.fdy * e2x;
    const float det = e1x * px + e1y * py + e1z * pz;
    const float tx = r.cox - p0x, ty = r.coy - p0y, tz = r.coz - p0z;
    const float U = tx * px + ty * py + tz * pz;
    const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
    const float V = r.fdx * qx + r.fdy * qy + r.fdz * qz;
    const float T = e2x * qx + e2y * qy + e2z * qz;
    const float E1 = fmaxf(fmaxf(fabsf(e1x), fabsf(e1y)), fabsf(e1z));
    const float E2 = fmaxf(fmaxf(fabsf(e2x), fabsf(e2y)), fabsf(e2z));
    const float TV = fmaxf(fmaxf(fabsf(tx), fabsf(ty)), fabsf(tz));
    const float k = 7.62939453125e-06f;              // 2^-17
    const float mt = k * (maxCoord + TV);
    float eDet = k * E1 * E2, eU = mt * E2, eV = mt * E1, eT = mt * E1 * E2;
    if (rnd != 0.f) {
        const float rt = 12.f * rnd * TV;
        eDet += 12.f * rnd * (E1 + E2);
        eU += 6.f * rnd * E2 + rt;
        eV += 6.f * rnd * E1 + rt;
        eT += rt * (E1 + E2) + 6.f * rnd * E1 * E2;
    }
    const float D = fabsf(det);
    if (!(D > eDet)) return true;                    // sign of det uncertain (or NaN): exact test decides
    const float sgn = det < 0.f ? -1.f : 1.f;
    const float Us = sgn * U, Vs = sgn * V, Ts = sgn * T, Dhi = D + eDet;
    if (Us + eU < 0.f) return false;                 // u < 0
    if (Us - eU > Dhi) return false;                 // u > 1
    if (Vs + eV < 0.f) return false;                 // v < 0
    if ((Us + Vs) - (eU + eV) > Dhi) return false;   // u + v > 1
    if (Ts + eT < 0.f) return false;                 // t below the culling origin (outside the world box)
    if (Ts - eT > r.ctmax * Dhi) return false;       // t > current maxDist
    return true;
}

// The same pre-test with a third answer (kernel variant 6): 0 = certainly rejected, 1 = undecided
// (the exact test decides), 2 = CERTAIN hit: every inequality of the exact test holds with the
// error bound on the other side, so the exact test accepts this triangle at some distance in
// [*tlo, *thi] along the culling ray (before its t <= maxDist comparison, which the caller makes
// certain separately).  A certain hit lets the ray shrink its culling interval at once and postpone
// the double-precision test until the lane has nothing else to do.
#if defined(__CUDA_ARCH__)
SPB_HD float fdivApprox(float a, float b) { return __fdividef(a, b); }   // <= 2 ulp for |b| < 2^126
#else
SPB_HD float fdivApprox(float a, float b) { return a / b; }
#endif

SPB_HD int triPretestClassify(const CullRay& r, float maxCoord, const U4& a, const U4& b, const U4& c, float* tlo,
                              float* thi) {
    const float p0x = asFloat(a.x), p0y = asFloat(a.y), p0z = asFloat(a.z);
    const float e1x = asFloat(b.x) - p0x, e1y = asFloat(b.y) - p0y, e1z = asFloat(b.z) - p0z;
    const float e2x = asFloat(c.x) - p0x, e2y = asFloat(c.y) - p0y, e2z = asFloat(c.z) - p0z;
    const float px = r.fdy * e2z - r.fdz * e2y, py = r.fdz * e2x - r.fdx * e2z, pz = r.fdx * e2y - r.fdy * e2x;
    const float det = e1x * px + e1y * py + e1z * pz;
    const float tx = r.cox - p0x, ty = r.coy - p0y, tz = r.coz - p0z;
    const float U = tx * px + ty * py + tz * pz;
    const float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
    const float V = r.fdx * qx + r.fdy * qy + r.fdz * qz;
    const float T = e2x * qx + e2y * qy + e2z * qz;
    const float E1 = fmaxf(fmaxf(fabsf(e1x), fabsf(e1y)), fabsf(e1z));
    const float E2 = fmaxf(fmaxf(fabsf(e2x), fabsf(e2y)), fabsf(e2z));
    const float TV = fmaxf(fmaxf(fabsf(tx), fabsf(ty)), fabsf(tz));
    const float k = 7.62939453125e-06f;              // 2^-17
    const float mt = k * (maxCoord + TV);
    const float eDet = k * E1 * E2, eU = mt * E2, eV = mt * E1, eT = mt * E1 * E2;
    const float D = fabsf(det);
    if (!(D > eDet)) return 1;
    const float sgn = det < 0.f ? -1.f : 1.f;
    const float Us = sgn * U, Vs = sgn * V, Ts = sgn * T, Dhi = D + eDet, Dlo = D - eDet;
    if (Us + eU < 0.f) return 0;
    if (Us - eU > Dhi) return 0;
    if (Vs + eV < 0.f) return 0;
    if ((Us + Vs) - (eU + eV) > Dhi) return 0;
    if (Ts + eT < 0.f) return 0;
    if (Ts - eT > r.ctmax * Dhi) return 0;
    // certain? (every comparison is false for NaN, which leaves the triangle undecided)
    const float Tlo = Ts - eT, Thi = Ts + eT;
    if (!(Dlo > 1e-10f && Us - eU > 0.f && Us + eU < Dlo && Vs - eV > 0.f && (Us + Vs) + (eU + eV) < Dlo && Tlo > 0.f))
        return 1;
    const float lo = fdivApprox(Tlo, Dhi) * 0.999999f, hi = fdivApprox(Thi, Dlo) * 1.000001f;
    if (!(lo > 1e-9f && hi < 1e30f)) return 1;
    *tlo = lo; *thi = hi;
    return 2;
}

SPB_HD bool triPretestMayHit(const RayState& r, float maxCoord, const U4& a, const U4& b, const U4& c, const float rnd = 0.f) {
    const CullRay cr = {r.cox, r.coy, r.coz, r.fdx, r.fdy, r.fdz, r.ctmax};
    return triPretestMayHit(cr, maxCoord, a, b, c, rnd);
}

// PRETEST = false: the candidate already passed the pre-test (cooperative kernel), go straight to the exact test
template <int TRI_FMT, bool PRETEST = true>
SPB_HD bool triTestRecord(const SceneParams& sp, const RayState& r, uint32_t index, double* t, double* u,
                          double* v, int32_t* id, int32_t* rank, TraceCounters* ctr = nullptr) {
    double p0[3], p1[3], p2[3];
    if (TRI_FMT == 0) {
        const TriF32* tp = (const TriF32*)sp.tris + index;
        const U4 a = ldg4(&tp->v0[0]), b = ldg4(&tp->v1[0]), c = ldg4(&tp->v2[0]);
        if (PRETEST && !triPretestMayHit(r, sp.max_coord, a, b, c)) return false;
        p0[0] = (double)asFloat(a.x); p0[1] = (double)asFloat(a.y); p0[2] = (double)asFloat(a.z);
        p1[0] = (double)asFloat(b.x); p1[1] = (double)asFloat(b.y); p1[2] = (double)asFloat(b.z);
        p2[0] = (double)asFloat(c.x); p2[1] = (double)asFloat(c.y); p2[2] = (double)asFloat(c.z);
        *id = (int32_t)a.w; *rank = (int32_t)b.w;
    } else {
        if (PRETEST) {
            const TriF32* pp = sp.pre_tris + index;
            const U4 pa = ldg4(&pp->v0[0]), pb = ldg4(&pp->v1[0]), pc = ldg4(&pp->v2[0]);
            if (!triPretestMayHit(r, sp.max_coord, pa, pb, pc, sp.pre_round)) return false;
        }
        const TriF64* tp = (const TriF64*)sp.tris + index;
        const U4 a = ldg4(&tp->v[0]), b = ldg4(&tp->v[2]), c = ldg4(&tp->v[4]), d = ldg4(&tp->v[6]),
                 e = ldg4(&tp->v[8]);
        auto mk = [](uint32_t lo, uint32_t hi) {
            const unsigned long long bits = ((unsigned long long)hi << 32) | lo;
            double x;
#if defined(__CUDA_ARCH__)
            x = __longlong_as_double((long long)bits);
#else
            __builtin_memcpy(&x, &bits, 8);
#endif
            return x;
        };
        p0[0] = mk(a.x, a.y); p0[1] = mk(a.z, a.w); p0[2] = mk(b.x, b.y);
        p1[0] = mk(b.z, b.w); p1[1] = mk(c.x, c.y); p1[2] = mk(c.z, c.w);
        p2[0] = mk(d.x, d.y); p2[1] = mk(d.z, d.w); p2[2] = mk(e.x, e.y);
        *id = (int32_t)e.z; *rank = (int32_t)e.w;
    }
    if (ctr) ctr->exact++;
    return triTestExact(r, p0, p1, p2, t, u, v);
}

// byte j of `w` as a float32 WITHOUT an integer-to-float conversion: the byte is dropped into mantissa
// bits 15..8 of 1.0f, which gives exactly 1 + q * 2^-15.  One PRMT on the ALU pipe instead of one I2F on
// the XU pipe: ncu on the r01 kernel showed the XU pipe (16 lanes per clock and SM) as its busiest unit
// (58 % over the whole kernel, saturated during node visits: 48 conversions = 384 XU cycles per
// warp-visit against ~270 issue slots; profiles/r02a_trace_xu_pipe.md).
// `one` = the bits of 1.0f, read from the kernel's parameter block (SceneParams::f32_one): PRMT takes ONE
// immediate, and with a literal 0x3f800000 ptxas spends it on the constant and loads the four selectors into
// registers with an extra IMAD.MOV each (+36 instructions per node visit); a constant-bank operand leaves
// the immediate to the selector.
#if defined(__CUDA_ARCH__)
SPB_HD float byteMant(uint32_t w, int j, uint32_t one) { return __uint_as_float(__byte_perm(w, one, 0x7604u | ((uint32_t)j << 4))); }
#else
SPB_HD float byteMant(uint32_t w, int j, uint32_t one) { return asFloat(one | (byteOf(w, j) << 8)); }
#endif

// Tests the 8 quantised child boxes of one node; returns the hit mask in traversal layout:
// bits 24..31 inner children at priority (slot ^ oct_inv), bits 0..23 triangles of hit leaves.
//
// Plane distance along the culling ray: t = q * ad + ao with ad = 2^e / d, ao = (p - o) / d.  With
// f = 1 + q * 2^-15 (byteMant) the same value is f * AD + AO, AD = ad * 2^15 (exact: a power of two),
// AO = ao - AD.  AO is rounded once (|error| <= |AO| * 2^-24); the near planes use AO - |AO| * 2^-23
// and the far planes AO + |AO| * 2^-23, so the interval only ever grows: by at most 1/256 of a
// quantisation step (plus 2^-23 of the distance to the node), which no traversal statistic notices.
SPB_HD uint32_t nodeHitMask(const U4& n0, const U4& n1, const U4& n2, const U4& n3, const U4& n4,
                            const RayState& r, uint32_t one) {
#if SPB_NODE_I2F
    const float adx = r.idx * asFloat(byteOf(n0.w, 0) << 23);
    const float ady = r.idy * asFloat(byteOf(n0.w, 1) << 23);
    const float adz = r.idz * asFloat(byteOf(n0.w, 2) << 23);
    const float aox = (asFloat(n0.x) - r.cox) * r.idx;
    const float aoy = (asFloat(n0.y) - r.coy) * r.idy;
    const float aoz = (asFloat(n0.z) - r.coz) * r.idz;
    const float anx = aox, any_ = aoy, anz = aoz, afx = aox, afy = aoy, afz = aoz;
(void)one;
#define SPB_Q(w, j) ((float)byteOf(w, j))
#else
    // 2^(e - 127 + 15); the builder keeps e - 127 <= 100, so the exponent field cannot overflow
    const float adx = r.idx * asFloat((byteOf(n0.w, 0) + 15u) << 23);
    const float ady = r.idy * asFloat((byteOf(n0.w, 1) + 15u) << 23);
    const float adz = r.idz * asFloat((byteOf(n0.w, 2) + 15u) << 23);
    const float aox = ffma(asFloat(n0.x) - r.cox, r.idx, -adx);
    const float aoy = ffma(asFloat(n0.y) - r.coy, r.idy, -ady);
    const float aoz = ffma(asFloat(n0.z) - r.coz, r.idz, -adz);
    const float kUlp2 = 1.1920929e-07f;      // 2^-23
    const float mx = fabsf(aox) * kUlp2, my = fabsf(aoy) * kUlp2, mz = fabsf(aoz) * kUlp2;
    const float anx = aox - mx, any_ = aoy - my, anz = aoz - mz, afx = aox + mx, afy = aoy + my, afz = aoz + mz;
#define SPB_Q(w, j) byteMant(w, j, one)
#endif
    const bool nx = r.idx < 0.f, ny = r.idy < 0.f, nz = r.idz < 0.f;
    const uint32_t oct4 = r.oct_inv * 0x01010101u;
    uint32_t hitmask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int h = 0; h < 2; h++) {
        // meta bytes of four slots at once: an inner child (position 24 + slot: bits 3 and 4 set) moves to
        // its place in the ray's front-to-back order, 24 + (slot ^ oct_inv)
        uint32_t meta4 = h ? n1.w : n1.z;
        meta4 ^= oct4 & (((meta4 >> 3) & (meta4 >> 4) & 0x01010101u) * 7u);
        const uint32_t lox = h ? n2.y : n2.x, loy = h ? n2.w : n2.z, loz = h ? n3.y : n3.x;
        const uint32_t hix = h ? n3.w : n3.z, hiy = h ? n4.y : n4.x, hiz = h ? n4.w : n4.z;
        const uint32_t nearx = nx ? hix : lox, farx = nx ? lox : hix;
        const uint32_t neary = ny ? hiy : loy, fary = ny ? loy : hiy;
        const uint32_t nearz = nz ? hiz : loz, farz = nz ? loz : hiz;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 4; j++) {
            const float tnx = ffma(SPB_Q(nearx, j), adx, anx);
            const float tny = ffma(SPB_Q(neary, j), ady, any_);
            const float tnz = ffma(SPB_Q(nearz, j), adz, anz);
            const float tfx = ffma(SPB_Q(farx, j), adx, afx);
            const float tfy = ffma(SPB_Q(fary, j), ady, afy);
            const float tfz = ffma(SPB_Q(farz, j), adz, afz);
            const float tmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
            const float tmax = fminf(fminf(tfx, tfy), fminf(tfz, r.ctmax));
            // (unary count) << position; an empty slot has meta 0 and contributes nothing
            if (tmin <= tmax) hitmask |= shlWrap((meta4 >> (8 * j + 5)) & 7u, meta4 >> (8 * j));
        }
    }
#undef SPB_Q
    return hitmask;
}

// The traversal state machine.  One step consumes one node group entry: nodePhase (visit one inner
// node, test its 8 child boxes) -> triPhase (test the triangles of the leaves that were hit) ->
// popPhase (next node group).  The cooperative kernel runs its own warp-wide pre-test between
// nodePhase and triPhase<false>.
template <int TRI_FMT, bool ANY_HIT>
struct Traverser {
    U2 stack[kStackCapacity];
    int sp_;
    U2 ngroup;       // x: child base index, y: [31..24] pending inner hits, [7..0] imask
    bool finished;

    SPB_HD void begin(bool valid) {
        sp_ = 0;
        ngroup.x = 0u; ngroup.y = 0x80000000u;   // the root as a one-node group
        finished = !valid;
    }

    SPB_HD U2 nodePhase(const SceneParams& sp, const RayState& r, TraceCounters* ctr) {
        const uint32_t hits = ngroup.y;
        const int bit = 31 - clz32(hits);
        ngroup.y &= ~(1u << bit);
        if (ngroup.y & 0xff000000u) stack[sp_++] = ngroup;
        const uint32_t slot = (uint32_t)(bit - 24) ^ r.oct_inv;
        const uint32_t rel = (uint32_t)popc32(hits & ~(0xffffffffu << slot) & 0xffu);
        const WideNode* np = sp.nodes + (ngroup.x + rel);
        const U4 n0 = ldg4((const char*)np), n1 = ldg4((const char*)np + 16), n2 = ldg4((const char*)np + 32),
                 n3 = ldg4((const char*)np + 48), n4 = ldg4((const char*)np + 64);
        const uint32_t hm = nodeHitMask(n0, n1, n2, n3, n4, r, sp.f32_one);
        if (ctr) ctr->nodes++;
        ngroup.x = n1.x;
        ngroup.y = (hm & 0xff000000u) | (n0.w >> 24);
        U2 tgroup;
        tgroup.x = n1.y;
        tgroup.y = hm & 0x00ffffffu;
        return tgroup;
    }

    template <bool PRETEST>
    SPB_HD void triPhase(const SceneParams& sp, RayState& r, U2 tgroup, TraceCounters* ctr) {
        while (tgroup.y) {
            const int i = ctz32(tgroup.y);
            tgroup.y &= tgroup.y - 1u;
            double t, u, v; int32_t id, rank;
            if (ctr && PRETEST) ctr->tris++;
            if (triTestRecord<TRI_FMT, PRETEST>(sp, r, tgroup.x + (uint32_t)i, &t, &u, &v, &id, &rank, ctr)) {
                if (ANY_HIT) { r.best_prim = id; r.best_t = t; finished = true; return; }
                // t <= best_t here.  Exact tie: the smaller rank wins (see spica_b200.h,
                // spb_bvh_import_binary); a first hit at t == tmax is accepted.
                if (t < r.best_t || r.best_prim < 0 || rank < r.best_rank) {
                    r.best_t = t; r.best_prim = id; r.best_rank = rank;
                    r.best_u = (float)u; r.best_v = (float)v;
                    r.ctmax = fminf(r.ctmax, cullTmax(r, t));
                }
            }
        }
    }

    SPB_HD void popPhase() {
        if (finished) return;
        if (!(ngroup.y & 0xff000000u)) {
            if (sp_ == 0) { finished = true; return; }
            ngroup = stack[--sp_];
        }
    }

    SPB_HD void step(const SceneParams& sp, RayState& r, TraceCounters* ctr) {
        const U2 tgroup = nodePhase(sp, r, ctr);
        triPhase<true>(sp, r, tgroup, ctr);
        popPhase();
    }

    // ---- early-select order (kernel variants 3, 4) -------------------------------------------------
    // The same walk with the phases reordered: visit -> select the NEXT node -> triangles.  The
    // loop-carried state is one node index (`cur`), known before the triangle phase starts, so the
    // kernel can start fetching that node while the warp is busy with the triangle tests.  A closer
    // hit found in the triangle phase cannot un-select the node; it is visited with the smaller
    // ctmax and culls its children (conservative, never changes the result).
    uint32_t cur;

    SPB_HD void begin2(bool valid) { sp_ = 0; cur = 0u; finished = !valid; }

    SPB_HD U2 visitLoaded(const U4& n0, const U4& n1, const U4& n2, const U4& n3, const U4& n4, const RayState& r,
                          U2* cgroup, TraceCounters* ctr, uint32_t one) {
        const uint32_t hm = nodeHitMask(n0, n1, n2, n3, n4, r, one);
        if (ctr) ctr->nodes++;
        cgroup->x = n1.x;
        cgroup->y = (hm & 0xff000000u) | (n0.w >> 24);
        U2 tgroup;
        tgroup.x = n1.y;
        tgroup.y = hm & 0x00ffffffu;
        return tgroup;
    }

    SPB_HD U2 visitPhase(const SceneParams& sp, const RayState& r, U2* cgroup, TraceCounters* ctr) {
        const char* np = (const char*)(sp.nodes + cur);
        const U4 n0 = ldg4(np), n1 = ldg4(np + 16), n2 = ldg4(np + 32), n3 = ldg4(np + 48), n4 = ldg4(np + 64);
        return visitLoaded(n0, n1, n2, n3, n4, r, cgroup, ctr, sp.f32_one);
    }

    // g: the child group of the node just visited.  Takes the nearest pending inner child of g, or of
    // the group on top of the stack when g has none; the rest of that group goes (back) on the stack.
    SPB_HD void selectPhase(const RayState& r, U2 g) { selectPhaseOn(r, g, LocalStack{stack}); }

    // The same with the stack held elsewhere (kernel variant 4: shared memory, entry i of a thread
    // at s[i * stride], which no two lanes of a warp ever share a bank for).
    struct LocalStack {
        U2* s;
        SPB_HD U2 get(int i) const { return s[i]; }
        SPB_HD void set(int i, U2 v) const { s[i] = v; }
    };
    template <class Stack>
    SPB_HD void selectPhaseOn(const RayState& r, U2 g, const Stack& st) {
        if (!(g.y & 0xff000000u)) {
            if (sp_ == 0) { finished = true; return; }
            g = st.get(--sp_);
        }
        const uint32_t hits = g.y;
        const int bit = 31 - clz32(hits);
        g.y &= ~(1u << bit);
        if (g.y & 0xff000000u) st.set(sp_++, g);
        const uint32_t slot = (uint32_t)(bit - 24) ^ r.oct_inv;
        cur = g.x + (uint32_t)popc32(hits & ~(0xffffffffu << slot) & 0xffu);
    }

    // ---- deferred exact test (kernel variant 6, float32-exact triangles) ----------------------------
    // `pend` is a triangle record known to be a certain hit closer than r.best_t (see
    // triPretestClassify) whose exact test has not been run yet; r.ctmax has already been shrunk to
    // its upper bound.  It may be replaced by a certain hit that is certainly nearer, and it is
    // resolved by the ordinary exact test when the ray has finished its walk - in the kernel all
    // lanes waiting for new rays do that together instead of one or two lanes after each node.
    // The answer is the reference's: every dropped candidate is certainly farther than a certain
    // hit, and whatever is not certain goes through the exact test at once, as in triPhase.
    uint32_t pend;
    float pend_lo;

    SPB_HD void acceptExact(const SceneParams& sp, RayState& r, uint32_t index, TraceCounters* ctr) {
        double t, u, v; int32_t id, rank;
        if (triTestRecord<TRI_FMT, false>(sp, r, index, &t, &u, &v, &id, &rank, ctr)) {
            if (ANY_HIT) { r.best_prim = id; r.best_t = t; finished = true; return; }
            if (t < r.best_t || r.best_prim < 0 || rank < r.best_rank) {
                r.best_t = t; r.best_prim = id; r.best_rank = rank;
                r.best_u = (float)u; r.best_v = (float)v;
                r.ctmax = fminf(r.ctmax, cullTmax(r, t));
            }
        }
    }

    SPB_HD void resolvePending(const SceneParams& sp, RayState& r, TraceCounters* ctr) {
        if (pend != 0xffffffffu) { const uint32_t p = pend; pend = 0xffffffffu; acceptExact(sp, r, p, ctr); }
    }

    // One node's candidates after the pre-test: `maybe` = undecided triangles, `cert` = certain hits
    // that are not certainly farther than the nearest certain one, whose distance bounds are [tlo, thi]
    // (tlo is only meaningful when `cert` has a single bit).
    SPB_HD void mergePhase(const SceneParams& sp, RayState& r, uint32_t base, uint32_t maybe, uint32_t cert, float tlo,
                           float thi, TraceCounters* ctr) {
        uint32_t now = maybe;
        if (cert) {
            const bool single = !(cert & (cert - 1u));
            // certainly inside the reference's t <= maxDist (the exact best so far, or the caller's tmax)
            const bool closer = (double)thi * 1.000002 + r.t_shift < r.best_t;
            if (single && closer && (pend == 0xffffffffu || thi < pend_lo)) {
                const uint32_t index = base + (uint32_t)ctz32(cert);
                if (ANY_HIT) {
                    const TriF32* tp = (const TriF32*)sp.tris + index;
                    r.best_prim = (int32_t)ldg4(&tp->v0[0]).w; finished = true; return;
                }
                pend = index; pend_lo = tlo;
                r.ctmax = fminf(r.ctmax, thi * 1.000002f + 1e-30f);
            } else {
                now |= cert;
            }
        }
        while (now) {
            const int i = ctz32(now);
            now &= now - 1u;
            acceptExact(sp, r, base + (uint32_t)i, ctr);
            if (ANY_HIT && r.best_prim >= 0) return;    // (`finished` may already be set by selectPhase)
        }
    }

    // scalar statement of the kernel's pooled protocol for one lane (unit-test emulation)
    SPB_HD void deferredTriPhase(const SceneParams& sp, RayState& r, U2 tgroup, TraceCounters* ctr) {
        uint32_t maybe = 0u, certAll = 0u, cert = 0u;
        float lo[24], hi[24], cmin = 3.0e38f, clo = 0.f;
        const CullRay cr = {r.cox, r.coy, r.coz, r.fdx, r.fdy, r.fdz, r.ctmax};
        for (uint32_t m = tgroup.y; m; m &= m - 1u) {
            const int i = ctz32(m);
            if (ctr) ctr->tris++;
            const TriF32* tp = (const TriF32*)sp.tris + (tgroup.x + (uint32_t)i);
            const U4 a = ldg4(&tp->v0[0]), b = ldg4(&tp->v1[0]), c = ldg4(&tp->v2[0]);
            const int cls = triPretestClassify(cr, sp.max_coord, a, b, c, &lo[i], &hi[i]);
            if (cls == 1) maybe |= 1u << i;
            if (cls == 2) { certAll |= 1u << i; if (hi[i] < cmin) cmin = hi[i]; }
        }
        for (uint32_t m = certAll; m; m &= m - 1u) {
            const int i = ctz32(m);
            if (lo[i] <= cmin) { cert |= 1u << i; if (hi[i] == cmin) clo = lo[i]; }
        }
        if (maybe | cert) mergePhase(sp, r, tgroup.x, maybe, cert, clo, cmin, ctr);
    }

    SPB_HD void step2(const SceneParams& sp, RayState& r, TraceCounters* ctr) {
        U2 g;
        const U2 tgroup = visitPhase(sp, r, &g, ctr);
        selectPhase(r, g);
        triPhase<true>(sp, r, tgroup, ctr);
    }
};



template <int TRI_FMT, bool ANY_HIT>
SPB_HD void traceRay(const SceneParams& sp, RayState& r, bool valid, TraceCounters* ctr) {
    Traverser<TRI_FMT, ANY_HIT> tr;
    tr.begin(valid);
    while (!tr.finished) tr.step(sp, r, ctr);
}

// early-select order + deferred exact test (kernel variant 6); the result is identical to traceRay's
template <bool ANY_HIT>
SPB_HD void traceRayDeferred(const SceneParams& sp, RayState& r, bool valid, TraceCounters* ctr) {
    Traverser<0, ANY_HIT> tr;
    tr.begin2(valid);
    tr.pend = 0xffffffffu; tr.pend_lo = 0.f;
    if (tr.finished) return;
    for (;;) {
        U2 g;
        const U2 tgroup = tr.visitPhase(sp, r, &g, ctr);
        tr.selectPhase(r, g);
        const bool last = tr.finished;
        if (tgroup.y) tr.deferredTriPhase(sp, r, tgroup, ctr);
        if (last || tr.finished) break;
    }
    if (!ANY_HIT || r.best_prim < 0) tr.resolvePending(sp, r, ctr);
}

// K node visits per triangle phase (kernel variant 5 uses K = 3); the result is identical to traceRay's
template <int TRI_FMT, bool ANY_HIT, int K>
SPB_HD void traceRayKVisits(const SceneParams& sp, RayState& r, bool valid, TraceCounters* ctr) {
    Traverser<TRI_FMT, ANY_HIT> tr;
    tr.begin2(valid);
    if (tr.finished) return;
    for (;;) {
        U2 tg[K];
        for (int q = 0; q < K; q++) {
            tg[q].x = 0u; tg[q].y = 0u;
            if (!tr.finished) { U2 g; tg[q] = tr.visitPhase(sp, r, &g, ctr); tr.selectPhase(r, g); }
        }
        const bool last = tr.finished;
        for (int q = 0; q < K; q++)
            if (!(ANY_HIT && r.best_prim >= 0)) tr.template triPhase<true>(sp, r, tg[q], ctr);
        if (last || tr.finished) break;
    }
}

// early-select order (Traverser::step2); the result is identical to traceRay's
template <int TRI_FMT, bool ANY_HIT>
SPB_HD void traceRayEarlySelect(const SceneParams& sp, RayState& r, bool valid, TraceCounters* ctr) {
    Traverser<TRI_FMT, ANY_HIT> tr;
    tr.begin2(valid);
    if (tr.finished) return;
    for (;;) {
        U2 g;
        const U2 tgroup = tr.visitPhase(sp, r, &g, ctr);
        tr.selectPhase(r, g);
        const bool last = tr.finished;          // nothing left after this node's triangles
        tr.template triPhase<true>(sp, r, tgroup, ctr);
        if (last || tr.finished) break;
    }
}

}  // namespace spb
