// Device PODs and functions of the shading side of the path tracer: vectors, the counter-based
// sampler, surface interaction, BSDF lobes (Lambertian, Fresnel-specular dielectric, microfacet
// reflection / transmission with Beckmann and GGX, specular conductor), area and environment
// lights.  Every function cites the reference code it restates (paths relative to
// /root/reference/sources).  The reference computes in double; this path shades in float32 (the
// image contract is statistical: relMSE against the reference, BASELINE.json) while ray casting
// keeps the exact double triangle test of trace_core.h.
//
// SPB_HD functions compile for the device and, in tests/emul, for the host.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/spica_b200.h"

#if defined(__CUDACC__)
#define SPB_SHD __host__ __device__ __forceinline__
#else
#define SPB_SHD inline
#endif

namespace spb {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kOffsetEps = 1.0e-3f;       // core/interaction.cc:17
constexpr float kDeltaEps = 1.0e-4f;        // core/bxdf.cc:19
constexpr float kRayInf = 1.0e32f;          // core/common.h:54

struct V3 { float x, y, z; };
SPB_SHD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
SPB_SHD V3 v3(float a) { return v3(a, a, a); }
SPB_SHD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
SPB_SHD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
SPB_SHD V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
SPB_SHD V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
SPB_SHD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
SPB_SHD V3 operator*(float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
SPB_SHD V3 operator/(V3 a, float s) { const float r = 1.0f / s; return v3(a.x * r, a.y * r, a.z * r); }
SPB_SHD V3 operator/(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
SPB_SHD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SPB_SHD float absDot(V3 a, V3 b) { return fabsf(dot(a, b)); }
SPB_SHD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
SPB_SHD float length(V3 a) { return sqrtf(dot(a, a)); }
SPB_SHD V3 normalize(V3 a) { const float n = length(a); return n > 0.f ? a * (1.0f / n) : a; }
SPB_SHD bool isBlack(V3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }          // core/spectrum.cc:649
SPB_SHD float gray(V3 a) { return 0.2126f * a.x + 0.7152f * a.y + 0.0722f * a.z; }       // core/spectrum.cc:677
SPB_SHD V3 vsqrt(V3 a) { return v3(sqrtf(fmaxf(a.x, 0.f)), sqrtf(fmaxf(a.y, 0.f)), sqrtf(fmaxf(a.z, 0.f))); }
SPB_SHD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// ---- counter-based sampler ------------------------------------------------------------------------
// Replaces the per-thread MT19937 `Independent` sampler reseeded from time(0) (core/integrator.cc:
// 51,69-73; samplers/independent.cc:17-23): a sample value is a pure function of
// (seed, pixel, sample index, dimension), so an image does not depend on how samples are
// partitioned over GPUs or waves.
SPB_SHD uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// The key of a path is 64 bits -- one word from (seed, pixel), one from (seed, sample) -- so that a
// 1024 x 1024 x 1024-spp image (2^30 paths) does not run into the birthday bound of a 32-bit key.
struct Key2 { uint32_t x, y; };
SPB_SHD Key2 samplerKey2(uint64_t seed, uint32_t pixel, uint32_t sample) {
    const uint32_t s0 = hash32((uint32_t)seed ^ 0x9e3779b9u), s1 = hash32((uint32_t)(seed >> 32) ^ 0x85ebca6bu);
    Key2 k;
    k.x = hash32(hash32(pixel ^ s0) + s1);
    k.y = hash32(hash32(sample + 0x27d4eb2fu + s1) ^ s0);
    return k;
}
// uniform in [0,1) with a 24-bit mantissa (the reference draws genrand_int32 / 2^32, core/random.cc:161)
SPB_SHD float sample1D(uint32_t k0, uint32_t k1, uint32_t dim) {
    const uint32_t h = hash32(k0 + dim * 0x9e3779b1u) ^ hash32(k1 ^ (dim * 0x68bc21ebu + 0x02e5be93u));
    return (float)(hash32(h) >> 8) * (1.0f / 16777216.0f);
}
#if defined(__CUDACC__)
SPB_SHD uint2 samplerKey(uint64_t seed, uint32_t pixel, uint32_t sample) { const Key2 k = samplerKey2(seed, pixel, sample); return make_uint2(k.x, k.y); }
SPB_SHD float sample1D(uint2 key, uint32_t dim) { return sample1D(key.x, key.y, dim); }
#endif

// sampler dimensions: 0,1 film, 2,3 lens, then kDimsPerBounce per path vertex
enum { kDimFilm = 0, kDimLens = 2, kDimBounce0 = 4, kDimsPerBounce = 10,
       kDLightPick = 0, kDLight = 1, kDShade = 3, kDBsdf = 5, kDRoulette = 7, kDLobeMis = 8, kDLobePath = 9 };

// ---- sampling warps ---------------------------------------------------------------------------------
// core/sampling.cc:151-164
SPB_SHD void concentricDisk(float u0, float u1, float* dx, float* dy) {
    const float ox = 2.0f * u0 - 1.0f, oy = 2.0f * u1 - 1.0f;
    if (ox == 0.f && oy == 0.f) { *dx = 0.f; *dy = 0.f; return; }
    float theta, r;
    if (fabsf(ox) > fabsf(oy)) { r = ox; theta = kPi * (oy / ox) * 0.25f; }
    else { r = oy; theta = 0.5f * kPi - kPi * (ox / oy) * 0.25f; }
    float s, c;
#if defined(__CUDA_ARCH__)
    sincosf(theta, &s, &c);
#else
    s = sinf(theta); c = cosf(theta);
#endif
    *dx = r * c; *dy = r * s;
}
// core/sampling.cc:175-179
SPB_SHD V3 cosineHemisphere(float u0, float u1) {
    float dx, dy;
    concentricDisk(u0, u1, &dx, &dy);
    return v3(dx, dy, sqrtf(fmaxf(0.f, 1.0f - dx * dx - dy * dy)));
}
// core/vect_math.h:115-123
SPB_SHD void coordinateSystem(V3 w, V3* u, V3* v) {
    if (fabsf(w.x) > fabsf(w.y)) *u = normalize(v3(-w.z, 0.f, w.x));
    else *u = normalize(v3(0.f, w.z, -w.y));
    *v = normalize(cross(w, *u));
}

// ---- geometry records ----------------------------------------------------------------------------
// Per-triangle shading record (indexed by the caller's primitive id), 112 B = 7 x 128-bit loads.
struct alignas(16) ShadeTri {
    float p0[3], e1x;          // e1 = p1 - p0, e2 = p2 - p0 (core/triangle.cc:99-100)
    float e1y, e1z, e2[2];
    float e2z, ng[3];          // ng = normalize(dpdu x dpdv): Interaction::normal (core/interaction.cc:116)
    float fn[3], area;         // Triangle::faceNormal_ (core/triangle.cc:26-30,44-50), Triangle::area (:212-216)
    float ss[3]; int32_t material;   // normalized dpdu, dpdv: the BSDF frame when ns == fn (core/bsdf.cc:17-21)
    float ts[3]; int32_t light;      // index into the light list or -1 (core/primitive.cc:66-68)
    int32_t has_normals, pad0, pad1, pad2;
};
static_assert(sizeof(ShadeTri) == 112, "ShadeTri layout");

struct TriGeom { V3 p0, e1, e2, ng, fn, ss, ts; float area; int material, light, has_normals; };

// a surface point as the integrator needs it: core/interaction.h (pos, normal, shading frame)
struct SurfacePoint {
    V3 p, ng;       // position, geometric normal
    V3 ss, ts, ns;  // BSDF frame: tangent, binormal, normal (each normalized, core/bsdf.cc:17-21)
};

// core/interaction.cc:16-30; the nextFloatUp/Down step is below float32 resolution and is omitted.
SPB_SHD V3 offsetRayOrigin(V3 p, V3 n, V3 w) {
    V3 off = n * kOffsetEps;
    if (dot(w, n) < 0.f) off = -off;
    return p + off;
}

// Moeller-Trumbore in float32 against ONE known triangle (a light), core/triangle.cc:98-117; used
// for the solid-angle pdf of Shape::pdf (core/shape.cc:39-48), never for scene visibility.
SPB_SHD bool triHitF32(const TriGeom& t, V3 o, V3 d, float* tHit) {
    const V3 pv = cross(d, t.e2);
    const float det = dot(t.e1, pv);
    if (det > -1e-12f && det < 1e-12f) return false;
    const float inv = 1.0f / det;
    const V3 tv = o - t.p0;
    const float u = dot(tv, pv) * inv;
    if (u < 0.f || u > 1.f) return false;
    const V3 qv = cross(tv, t.e1);
    const float v = dot(d, qv) * inv;
    if (v < 0.f || u + v > 1.f) return false;
    const float tt = dot(t.e2, qv) * inv;
    if (tt <= 1e-12f) return false;
    *tHit = tt;
    return true;
}

// Shape::pdf(pObj, wi) (core/shape.cc:39-48) for a triangle light, from the surface point `sp`.
SPB_SHD float trianglePdfSolidAngle(const TriGeom& lt, const SurfacePoint& sp, V3 wi) {
    const V3 o = offsetRayOrigin(sp.p, sp.ng, wi);
    float t;
    if (!triHitF32(lt, o, wi, &t)) return 0.f;
    const V3 hp = o + wi * t;
    const V3 dd = sp.p - hp;
    const float ret = dot(dd, dd) / (absDot(lt.ng, -wi) * lt.area);
    return isinf(ret) ? 0.f : ret;
}

// core/mis.cc:138-141
SPB_SHD float powerHeuristic(float f, float g) { return (f * f) / (f * f + g * g); }

// ---- Fresnel --------------------------------------------------------------------------------------
// core/fresnel.cc:65-85
SPB_SHD float frDielectric(float cosI, float etaI, float etaT) {
    cosI = clampf(cosI, -1.f, 1.f);
    if (!(cosI > 0.f)) { const float t = etaI; etaI = etaT; etaT = t; cosI = fabsf(cosI); }
    const float sinI = sqrtf(fmaxf(0.f, 1.f - cosI * cosI));
    const float sinT = etaI / etaT * sinI;
    if (sinT >= 1.f) return 1.f;
    const float cosT = sqrtf(fmaxf(0.f, 1.f - sinT * sinT));
    const float rpa = ((etaT * cosI) - (etaI * cosT)) / ((etaT * cosI) + (etaI * cosT));
    const float rpe = ((etaI * cosI) - (etaT * cosT)) / ((etaI * cosI) + (etaT * cosT));
    return 0.5f * (rpa * rpa + rpe * rpe);
}
// core/fresnel.cc:40-63 with etaI = 1 (bsdfs/roughconductor.cc:58)
SPB_SHD V3 frConductor(float cosI, V3 eta, V3 k) {
    cosI = clampf(cosI, -1.f, 1.f);
    const float c2 = cosI * cosI, s2 = 1.f - c2;
    const V3 eta2 = eta * eta, k2 = k * k;
    const V3 t0 = eta2 - k2 - v3(s2);
    const V3 a2b2 = vsqrt(t0 * t0 + 4.0f * eta2 * k2);
    const V3 t1 = a2b2 + v3(c2);
    const V3 a = vsqrt(0.5f * (a2b2 + t0));
    const V3 t2 = 2.0f * cosI * a;
    const V3 rs = (t1 - t2) / (t1 + t2);
    const V3 t3 = c2 * a2b2 + v3(s2 * s2);
    const V3 t4 = t2 * s2;
    const V3 rp = rs * (t3 - t4) / (t3 + t4 + v3(1e-12f));
    return 0.5f * (rp + rs);
}
// core/vect_math.h:144-153
SPB_SHD bool refractDir(V3 wi, V3 n, float eta, V3* wt) {
    const float cosI = dot(n, wi);
    const float sin2I = fmaxf(0.f, 1.f - cosI * cosI);
    const float sin2T = eta * eta * sin2I;
    if (sin2T >= 1.f) return false;
    const float cosT = sqrtf(1.f - sin2T);
    *wt = eta * (-wi) + (eta * cosI - cosT) * n;
    return true;
}

// ---- microfacet distributions (core/microfacet.cc) ---------------------------------------------------
SPB_SHD float cos2Theta(V3 w) { return w.z * w.z; }
SPB_SHD float sin2Theta(V3 w) { return fmaxf(0.f, 1.f - w.z * w.z); }
SPB_SHD float sinTheta(V3 w) { return sqrtf(sin2Theta(w)); }
SPB_SHD float tanTheta(V3 w) { return sinTheta(w) / w.z; }
SPB_SHD float tan2Theta(V3 w) { return sin2Theta(w) / cos2Theta(w); }
SPB_SHD float cosPhi(V3 w) { const float s = sinTheta(w); return s == 0.f ? 1.f : clampf(w.x / s, -1.f, 1.f); }
SPB_SHD float sinPhi(V3 w) { const float s = sinTheta(w); return s == 0.f ? 0.f : clampf(w.y / s, -1.f, 1.f); }

struct Microfacet { float ax, ay; int ggx; };

// core/math.h:14-42 (float polynomial), :44-65
SPB_SHD float refErfinv(float x) {
    x = clampf(x, -0.99999f, 0.99999f);
    float w = -logf((1.f - x) * (1.f + x)), p;
    if (w < 5.f) {
        w -= 2.5f;
        p = 2.81022636e-08f; p = 3.43273939e-07f + p * w; p = -3.5233877e-06f + p * w;
        p = -4.39150654e-06f + p * w; p = 0.00021858087f + p * w; p = -0.00125372503f + p * w;
        p = -0.00417768164f + p * w; p = 0.246640727f + p * w; p = 1.50140941f + p * w;
    } else {
        w = sqrtf(w) - 3.f;
        p = -0.000200214257f; p = 0.000100950558f + p * w; p = 0.00134934322f + p * w;
        p = -0.00367342844f + p * w; p = 0.00573950773f + p * w; p = -0.0076224613f + p * w;
        p = 0.00943887047f + p * w; p = 1.00167406f + p * w; p = 2.83297682f + p * w;
    }
    return p * x;
}
SPB_SHD float refErf(float x) {
    const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f,
                p = 0.3275911f;
    const float sign = x < 0.f ? -1.f : 1.f;
    x = fabsf(x);
    const float t = 1.f / (1.f + p * x);
    const float y = 1.f - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * expf(-x * x);
    return sign * y;
}

// D: core/microfacet.cc:193-201 (GGX), :267-278 (Beckmann)
SPB_SHD float mfD(const Microfacet& m, V3 wh) {
    const float t2 = tan2Theta(wh);
    if (isinf(t2) || isnan(t2)) return 0.f;
    const float c4 = cos2Theta(wh) * cos2Theta(wh);
    const float cp = cosPhi(wh), sp = sinPhi(wh);
    const float e = (cp * cp / (m.ax * m.ax) + sp * sp / (m.ay * m.ay)) * t2;
    if (m.ggx) return 1.f / (kPi * m.ax * m.ay * c4 * (1.f + e) * (1.f + e));
    return expf(-e) / (kPi * m.ax * m.ay * c4);
}
// lambda: core/microfacet.cc:243-252 (GGX), :331-345 (Beckmann)
SPB_SHD float mfLambda(const Microfacet& m, V3 w) {
    const float at = fabsf(tanTheta(w));
    if (isinf(at) || isnan(at)) return 0.f;
    const float cp = cosPhi(w), sp = sinPhi(w);
    const float alpha = sqrtf(cp * cp * m.ax * m.ax + sp * sp * m.ay * m.ay);
    if (m.ggx) {
        const float ai = alpha * at;
        return (-1.f + sqrtf(1.f + ai * ai)) * 0.5f;
    }
    const float a = 1.f / (alpha * at);
    if (a >= 1.6f) return 0.f;
    return (1.f - 1.259f * a + 0.396f * a * a) / (3.535f * a + 2.181f * a * a);
}
// core/microfacet.cc:154-172
SPB_SHD float mfG1(const Microfacet& m, V3 w, V3 wh) {
    if (w.z * dot(w, wh) <= 0.f) return 0.f;
    return 1.f / (1.f + mfLambda(m, w));
}
SPB_SHD float mfG(const Microfacet& m, V3 wo, V3 wi, V3 wh) {
    if (wi.z * dot(wi, wh) <= 0.f) return 0.f;
    if (wo.z * dot(wo, wh) <= 0.f) return 0.f;
    return 1.f / (1.f + mfLambda(m, wo) + mfLambda(m, wi));
}
// visible-normal pdf (sampleVisibleArea defaults to true), core/microfacet.cc:174-180
SPB_SHD float mfPdf(const Microfacet& m, V3 wo, V3 wh) {
    return mfD(m, wh) * mfG1(m, wo, wh) * absDot(wo, wh) / fabsf(wo.z);
}

// core/microfacet.cc:16-61
SPB_SHD void ggxSample11(float cosT, float u0, float u1, float* sx, float* sy) {
    if (cosT > 0.9999f) {
        const float r = sqrtf(u0 / (1.f - u0)), phi = 2.f * kPi * u1;
        *sx = r * cosf(phi); *sy = r * sinf(phi);
        return;
    }
    const float sinT = sqrtf(fmaxf(0.f, 1.f - cosT * cosT));
    const float tanT = sinT / cosT;
    const float a = 1.f / tanT;
    const float G1 = 2.f / (1.f + sqrtf(1.f + 1.f / (a * a)));
    const float A = 2.f * u0 / G1 - 1.f;
    float tmp = 1.f / (A * A - 1.f);
    if (tmp > 1e10f) tmp = 1e10f;
    const float B = tanT;
    const float D = sqrtf(fmaxf(0.f, B * B * tmp * tmp - (A * A - B * B) * tmp));
    const float s1 = B * tmp - D, s2 = B * tmp + D;
    *sx = (A < 0.f || s2 > 1.f / tanT) ? s1 : s2;
    float S, U;
    if (u1 > 0.5f) { S = 1.f; U = 2.f * (u1 - 0.5f); } else { S = -1.f; U = 2.f * (0.5f - u1); }
    const float z = (U * (U * (U * 0.27385f - 0.73369f) + 0.46341f)) /
                    (U * (U * (U * 0.093073f + 0.309420f) - 1.f) + 0.597999f);
    *sy = S * z * sqrtf(1.f + (*sx) * (*sx));
}
// core/microfacet.cc:87-133
SPB_SHD void beckmannSample11(float cosI, float u0, float u1, float* sx, float* sy) {
    if (cosI > 0.9999f) {
        const float r = sqrtf(-logf(1.f - u0));
        *sx = r * cosf(2.f * kPi * u1); *sy = r * sinf(2.f * kPi * u1);
        return;
    }
    const float sinI = sqrtf(fmaxf(0.f, 1.f - cosI * cosI));
    const float tanI = sinI / cosI, cotI = 1.f / tanI;
    float a = -1.f, c = refErf(cotI);
    const float sampleX = fmaxf(u0, 1e-6f);
    const float thetaI = acosf(cosI);
    const float fit = 1.f + thetaI * (-0.876f + thetaI * (0.4265f - 0.0594f * thetaI));
    float b = c - (1.f + c) * powf(1.f - sampleX, fit);
    const float sqrtPiInv = 0.5641895835477563f;
    const float norm = 1.f / (1.f + c + sqrtPiInv * tanI * expf(-cotI * cotI));
    for (int it = 0; it < 16; it++) {
        if (b < a || c < b) b = 0.5f * (a + c);
        const float xm = refErfinv(b);
        const float value = norm * (1.f + b + sqrtPiInv * tanI * expf(-xm * xm)) - sampleX;
        const float deriv = norm * (1.f - xm * tanI);
        if (fabsf(value) < 1e-6f) break;
        if (value > 0.f) c = b; else a = b;
        b -= value / deriv;
    }
    *sx = refErfinv(b);
    *sy = refErfinv(2.f * fmaxf(u1, 1e-6f) - 1.f);
}
// visible-normal sampling: core/microfacet.cc:63-85 (GGX), :135-157 (Beckmann), :203-233, :280-321
SPB_SHD V3 mfSample(const Microfacet& m, V3 wo, float u0, float u1) {
    const bool flip = wo.z < 0.f;
    const V3 wi = flip ? -wo : wo;
    const V3 ws = normalize(v3(m.ax * wi.x, m.ay * wi.y, wi.z));
    float sx, sy;
    if (m.ggx) ggxSample11(ws.z, u0, u1, &sx, &sy); else beckmannSample11(ws.z, u0, u1, &sx, &sy);
    const float cp = cosPhi(ws), sp = sinPhi(ws);
    const float tmp = cp * sx - sp * sy;
    sy = sp * sx + cp * sy;
    sx = tmp;
    sx *= m.ax; sy *= m.ay;
    V3 wh = normalize(v3(-sx, -sy, 1.f));
    return flip ? -wh : wh;
}

// ---- BSDF (one lobe per material, as the in-scope material plugins build them) --------------------
enum { kBxReflection = 1, kBxTransmission = 2, kBxDiffuse = 4, kBxGlossy = 8, kBxSpecular = 16,
       kBxAll = 31, kBxNonSpecular = 15 };

struct Bsdf {
    uint32_t allowed;    // bit t set: lobe type t can occur in this scene.  The shade kernel is instantiated per
                         // set of lobe types (integrator.cu, shadeKernel<.., TYPES>); with `allowed` a compile-time
                         // constant every `lobeLive` test below folds and the code of the other lobes is never emitted.
    int type;            // SPB_MAT_*; SPB_MAT_NONE = no lobes
    int flags;           // BxDFType of the single lobe (core/bxdf.h)
    V3 kr, kt, eta, k;
    Microfacet mf;
    float ior;           // etaB of the dielectric lobes; etaA = 1 (bsdfs/dielectric.cc:41)
};

// what the material plugins' setScatterFuncs build (bsdfs/diffuse.cc:23-32, dielectric.cc:30-42,
// roughconductor.cc:38-75, roughdielectric.cc:41-83)
SPB_SHD bool lobeLive(uint32_t allowed, int t) { return (allowed >> t) & 1u; }

SPB_SHD Bsdf makeBsdf(const spb_material& m, uint32_t allowed = 0xffffffffu) {
    Bsdf b;
    b.allowed = allowed;
    b.type = m.type; b.flags = 0;
    b.kr = v3(fmaxf(m.kr[0], 0.f), fmaxf(m.kr[1], 0.f), fmaxf(m.kr[2], 0.f));   // Spectrum::clamp
    b.kt = v3(fmaxf(m.kt[0], 0.f), fmaxf(m.kt[1], 0.f), fmaxf(m.kt[2], 0.f));
    b.eta = v3(m.eta[0], m.eta[1], m.eta[2]);
    b.k = v3(m.k[0], m.k[1], m.k[2]);
    b.mf.ax = m.alpha_u; b.mf.ay = m.alpha_v; b.mf.ggx = m.distribution == SPB_DISTR_GGX;
    b.ior = m.eta[0];
    switch (m.type) {
    case SPB_MAT_DIFFUSE: if (!lobeLive(allowed, SPB_MAT_DIFFUSE)) { b.type = SPB_MAT_NONE; break; }
        if (!isBlack(b.kr)) b.flags = kBxReflection | kBxDiffuse; else b.type = SPB_MAT_NONE;
        break;
    case SPB_MAT_DIELECTRIC: if (!lobeLive(allowed, SPB_MAT_DIELECTRIC)) { b.type = SPB_MAT_NONE; break; }
        if (isBlack(b.kr) && isBlack(b.kt)) b.type = SPB_MAT_NONE;
        else b.flags = kBxReflection | kBxTransmission | kBxSpecular;
        break;
    case SPB_MAT_ROUGHCONDUCTOR: if (!lobeLive(allowed, SPB_MAT_ROUGHCONDUCTOR)) { b.type = SPB_MAT_NONE; break; }
        b.kr = v3(1.f);
        if (m.alpha_u == 0.f && m.alpha_v == 0.f) { b.type = SPB_MAT_CONDUCTOR; b.flags = kBxReflection | kBxSpecular; }
        else b.flags = kBxReflection | kBxGlossy;
        break;
    case SPB_MAT_CONDUCTOR: if (!lobeLive(allowed, SPB_MAT_CONDUCTOR)) { b.type = SPB_MAT_NONE; break; }
        b.flags = kBxReflection | kBxSpecular;
        break;
    case SPB_MAT_ROUGHDIELECTRIC: if (!lobeLive(allowed, SPB_MAT_ROUGHDIELECTRIC)) { b.type = SPB_MAT_NONE; break; }
        if (isBlack(b.kr) && isBlack(b.kt)) b.type = SPB_MAT_NONE;
        else if (m.alpha_u == 0.f && m.alpha_v == 0.f) { b.type = SPB_MAT_DIELECTRIC; b.flags = kBxReflection | kBxTransmission | kBxSpecular; }
        else b.flags = kBxReflection | kBxTransmission | kBxGlossy;
        break;
    case SPB_MAT_PLASTIC: if (!lobeLive(allowed, SPB_MAT_PLASTIC)) { b.type = SPB_MAT_NONE; break; }                                                         // bsdfs/plastic.cc:22-33,118-128: kr = Ks, kt = Kd
        b.flags = kBxReflection | kBxSpecular | kBxDiffuse;
        break;
    case SPB_MAT_ROUGHPLASTIC: if (!lobeLive(allowed, SPB_MAT_ROUGHPLASTIC)) { b.type = SPB_MAT_NONE; break; } {                                                  // bsdfs/roughplastic.cc:22-33,130-165
        const float rough = fmaxf(0.001f, m.alpha_u);
        b.mf.ax = b.mf.ay = rough;
        if (isBlack(b.kr)) { b.type = SPB_MAT_DIFFUSE; b.kr = b.kt; b.flags = kBxReflection | kBxDiffuse; }   // LambertianReflection(kd), added even when kd is black
        else b.flags = kBxReflection | kBxDiffuse | kBxGlossy;
        break;
    }
    default: b.type = SPB_MAT_NONE; break;
    }
    return b;
}

// core/fresnel.cc:87-95
SPB_SHD float frDiffuseReflectance(float eta) {
    if (eta >= 1.f) return -1.4399f / (eta * eta) + 0.7099f / eta + 0.6681f + 0.0636f * eta;
    return -0.4399f + 0.7099f / eta - 0.3319f / (eta * eta) + 0.0636f / (eta * eta * eta);
}
// PlasticBRDF / RoughPlasticBRDF (bsdfs/plastic.cc:20-88, bsdfs/roughplastic.cc:20-98): kr = Ks, kt = Kd, etaA = 1, etaB = ior.
// Their f / pdf / sample are kept as the reference wrote them, including what looks unintended: the
// "specular" test of PlasticBRDF::f/pdf mirrors wi about the normal and compares it with wi itself
// (true only for wi exactly along the normal), and the diffuse branches sample the +z hemisphere
// whatever side wo is on.
SPB_SHD float plasticProbSpecular(const Bsdf& b, float Fo) {
    const float w = gray(b.kr) / (gray(b.kt) + gray(b.kr));
    return (Fo * w) / (Fo * w + (1.f - Fo) * (1.f - w));
}
SPB_SHD V3 plasticDiffuse(const Bsdf& b, float Fo, float cosI) {
    const float invEta = 1.f / b.ior;
    const float Fi = frDielectric(cosI, 1.f, b.ior);
    return b.kt * ((1.f - Fo) * (1.f - Fi) * invEta * invEta * kInvPi / (1.f - frDiffuseReflectance(b.ior)));
}
SPB_SHD bool plasticAlongNormal(V3 wi) { return wi.z * wi.z - wi.x * wi.x - wi.y * wi.y > 1.f; }   // > 1 - 1e-12 in float

// BSDF::numComponents (core/bsdf.cc:28-36): a lobe counts only if ALL its flags are inside `type`
SPB_SHD int bsdfNumComponents(const Bsdf& b, int type) { return (b.flags != 0 && (b.flags & type) == b.flags) ? 1 : 0; }

// MicrofacetReflection::f (core/bxdf.cc:267-282)
SPB_SHD V3 mfReflF(const Bsdf& b, V3 wo, V3 wi) {
    const float co = fabsf(wo.z), ci = fabsf(wi.z);
    V3 wh = wi + wo;
    if (ci == 0.f || co == 0.f) return v3(0.f);
    if (wh.x == 0.f && wh.y == 0.f && wh.z == 0.f) return v3(0.f);
    wh = normalize(wh);
    const V3 F = frConductor(fabsf(dot(wi, wh)), b.eta, b.k);
    return b.kr * F * (mfD(b.mf, wh) * mfG(b.mf, wo, wi, wh) / (4.f * ci * co));
}
SPB_SHD float mfReflPdf(const Bsdf& b, V3 wo, V3 wi) {      // core/bxdf.cc:297-301
    if (!(wo.z * wi.z > 0.f)) return 0.f;
    const V3 wh = normalize(wo + wi);
    return mfPdf(b.mf, wo, wh) / (4.f * dot(wo, wh));
}
// MicrofacetTransmission::f / pdf with an explicit half vector (core/bxdf.cc:336-359, 408-424)
SPB_SHD V3 mfTransF(const Bsdf& b, V3 wo, V3 wi, V3 wh) {
    const float co = wo.z, ci = wi.z;
    if (co == 0.f || ci == 0.f) return v3(0.f);
    if (wo.z * dot(wo, wh) <= 0.f) return v3(0.f);
    if (wi.z * dot(wi, wh) <= 0.f) return v3(0.f);
    const float F = frDielectric(dot(wo, wh), 1.f, b.ior);
    if (wo.z * wi.z > 0.f) return b.kr * (F * mfD(b.mf, wh) * mfG(b.mf, wo, wi, wh) / (4.f * co * ci));
    const float eta = dot(wo, wh) > 0.f ? b.ior : 1.f / b.ior;
    const float factor = 1.f / eta;
    const float sd = dot(wo, wh) + eta * dot(wi, wh);
    return b.kt * ((1.f - F) * fabsf(mfD(b.mf, wh) * mfG(b.mf, wo, wi, wh) * eta * eta * absDot(wi, wh) * absDot(wo, wh) *
                                    factor * factor / (ci * co * sd * sd)));
}
SPB_SHD float mfTransPdf(const Bsdf& b, V3 wo, V3 wi, V3 wh) {
    if (wo.z * dot(wo, wh) <= 0.f) return 0.f;
    if (wi.z * dot(wi, wh) <= 0.f) return 0.f;
    const float F = frDielectric(dot(wo, wh), 1.f, b.ior);
    if (wo.z * wi.z > 0.f) return F * mfPdf(b.mf, wo, wh) / (4.f * absDot(wo, wh));
    const float eta = dot(wo, wh) > 0.f ? b.ior : 1.f / b.ior;
    const float sd = dot(wo, wh) + eta * dot(wi, wh);
    const float dwh = fabsf((eta * eta * dot(wi, wh)) / (sd * sd));
    return (1.f - F) * mfPdf(b.mf, wo, wh) * dwh;
}
SPB_SHD V3 mfTransHalf(const Bsdf& b, V3 wo, V3 wi) {        // core/bxdf.cc:321-334, 395-406
    V3 wh;
    if (wo.z * wi.z > 0.f) wh = normalize(wo + wi);
    else { const float eta = wo.z > 0.f ? b.ior : 1.f / b.ior; wh = normalize(wo + wi * eta); }
    return wh.z >= 0.f ? wh : -wh;
}

// the lobe's f(wo, wi) in the local frame
SPB_SHD V3 lobeF(const Bsdf& b, V3 wo, V3 wi) {
    switch (b.type) {
    case SPB_MAT_DIFFUSE: if (!lobeLive(b.allowed, SPB_MAT_DIFFUSE)) break; return b.kr * kInvPi;                                   // core/bxdf.cc:56-58
    case SPB_MAT_ROUGHCONDUCTOR: if (!lobeLive(b.allowed, SPB_MAT_ROUGHCONDUCTOR)) break; return mfReflF(b, wo, wi);
    case SPB_MAT_ROUGHDIELECTRIC: if (!lobeLive(b.allowed, SPB_MAT_ROUGHDIELECTRIC)) break; return mfTransF(b, wo, wi, mfTransHalf(b, wo, wi));
    case SPB_MAT_CONDUCTOR: if (!lobeLive(b.allowed, SPB_MAT_CONDUCTOR)) break;                                                       // core/bxdf.cc:86-91
        if (dot(v3(-wo.x, -wo.y, wo.z), wi) > 1.f - kDeltaEps) return b.kr / fabsf(wi.z);
        return v3(0.f);
    case SPB_MAT_DIELECTRIC: if (!lobeLive(b.allowed, SPB_MAT_DIELECTRIC)) break; {                                                    // core/bxdf.cc:170-191
        const float F = frDielectric(wo.z, 1.f, b.ior);
        if (wo.z * wi.z > 0.f) {
            if (dot(v3(-wo.x, -wo.y, wo.z), wi) > 1.f - kDeltaEps) return b.kr * (F / fabsf(wi.z));
        } else {
            const bool entering = wo.z > 0.f;
            const float etaI = entering ? 1.f : b.ior, etaT = entering ? b.ior : 1.f;
            V3 wt;
            if (!refractDir(wo, v3(0.f, 0.f, wo.z < 0.f ? -1.f : 1.f), etaI / etaT, &wt)) return v3(0.f);
            if (dot(wi, wt) > 1.f - kDeltaEps) return b.kt * ((1.f - F) / fabsf(wi.z));
        }
        return v3(0.f);
    }
    case SPB_MAT_PLASTIC: if (!lobeLive(b.allowed, SPB_MAT_PLASTIC)) break; {                                                       // bsdfs/plastic.cc:32-46
        const float Fo = frDielectric(wo.z, 1.f, b.ior);
        if (plasticAlongNormal(wi)) return b.kr * (Fo / fabsf(wi.z));
        return plasticDiffuse(b, Fo, wi.z);
    }
    case SPB_MAT_ROUGHPLASTIC: if (!lobeLive(b.allowed, SPB_MAT_ROUGHPLASTIC)) break; {                                                  // bsdfs/roughplastic.cc:35-47
        if (wo.z == 0.f || wi.z == 0.f) return v3(0.f);
        if (!(wo.z * wi.z > 0.f)) return v3(0.f);
        const V3 wh = normalize(wi + wo);
        const float F = frDielectric(dot(wo, wh), 1.f, b.ior);
        return b.kr * (F * mfD(b.mf, wh) * mfG(b.mf, wo, wi, wh) / (4.f * wo.z * wi.z)) + b.kt * ((1.f - F) * kInvPi);
    }
    default: return v3(0.f);
    }
    return v3(0.f);
}
SPB_SHD float lobePdf(const Bsdf& b, V3 wo, V3 wi) {
    switch (b.type) {
    case SPB_MAT_DIFFUSE: if (!lobeLive(b.allowed, SPB_MAT_DIFFUSE)) break; return wo.z * wi.z > 0.f ? fabsf(wi.z) * kInvPi : 0.f;  // core/bxdf.cc:43-45
    case SPB_MAT_ROUGHCONDUCTOR: if (!lobeLive(b.allowed, SPB_MAT_ROUGHCONDUCTOR)) break; return mfReflPdf(b, wo, wi);
    case SPB_MAT_ROUGHDIELECTRIC: if (!lobeLive(b.allowed, SPB_MAT_ROUGHDIELECTRIC)) break; return mfTransPdf(b, wo, wi, mfTransHalf(b, wo, wi));
    case SPB_MAT_CONDUCTOR: if (!lobeLive(b.allowed, SPB_MAT_CONDUCTOR)) break; return dot(v3(-wo.x, -wo.y, wo.z), wi) > 1.f - kDeltaEps ? 1.f : 0.f;   // :103-108
    case SPB_MAT_DIELECTRIC: if (!lobeLive(b.allowed, SPB_MAT_DIELECTRIC)) break; {                                                    // core/bxdf.cc:228-248
        const float F = frDielectric(wo.z, 1.f, b.ior);
        if (wo.z * wi.z > 0.f) {
            if (dot(v3(-wo.x, -wo.y, wo.z), wi) > 1.f - kDeltaEps) return F;
        } else {
            const bool entering = wo.z > 0.f;
            const float etaI = entering ? 1.f : b.ior, etaT = entering ? b.ior : 1.f;
            V3 wt;
            if (!refractDir(wo, v3(0.f, 0.f, wo.z < 0.f ? -1.f : 1.f), etaI / etaT, &wt)) return 0.f;
            if (dot(wi, wt) > 1.f - kDeltaEps) return 1.f - F;
        }
        return 0.f;
    }
    case SPB_MAT_PLASTIC: if (!lobeLive(b.allowed, SPB_MAT_PLASTIC)) break; {                                                       // bsdfs/plastic.cc:70-82
        const float ps = plasticProbSpecular(b, frDielectric(wo.z, 1.f, b.ior));
        if (plasticAlongNormal(wi)) return ps;
        return (1.f - ps) * fabsf(wi.z) * kInvPi;
    }
    case SPB_MAT_ROUGHPLASTIC: if (!lobeLive(b.allowed, SPB_MAT_ROUGHPLASTIC)) break; {                                                  // bsdfs/roughplastic.cc:88-96
        if (!(wo.z * wi.z > 0.f)) return 0.f;
        const V3 wh = normalize(wi + wo);
        const float F = frDielectric(dot(wo, wh), 1.f, b.ior);
        return F * mfPdf(b.mf, wo, wh) / (4.f * absDot(wo, wh)) + (1.f - F) * fabsf(wi.z) * kInvPi;
    }
    default: return 0.f;
    }
    return 0.f;
}
// the lobe's sample(); u2 replaces the hidden thread_local Random of MicrofacetTransmission::sample
// (core/bxdf.cc:18,364). Returns f; *pdf = 0 means "no sample".
SPB_SHD V3 lobeSample(const Bsdf& b, V3 wo, float u0, float u1, float u2, V3* wi, float* pdf, int* sampled) {
    *pdf = 0.f; *sampled = b.flags;
    switch (b.type) {
    case SPB_MAT_DIFFUSE: if (!lobeLive(b.allowed, SPB_MAT_DIFFUSE)) break; {                                                       // core/bxdf.cc:35-41
        *wi = cosineHemisphere(u0, u1);
        if (wo.z < 0.f) wi->z = -wi->z;
        *pdf = lobePdf(b, wo, *wi);
        return b.kr * kInvPi;
    }
    case SPB_MAT_CONDUCTOR: if (!lobeLive(b.allowed, SPB_MAT_CONDUCTOR)) break; {                                                     // core/bxdf.cc:93-101
        *wi = v3(-wo.x, -wo.y, wo.z);
        *pdf = 1.f;
        return frConductor(fabsf(wi->z), b.eta, b.k) * b.kr / fabsf(wi->z);
    }
    case SPB_MAT_DIELECTRIC: if (!lobeLive(b.allowed, SPB_MAT_DIELECTRIC)) break; {                                                    // core/bxdf.cc:193-226
        const float F = frDielectric(wo.z, 1.f, b.ior);
        if (u0 < F) {
            *wi = v3(-wo.x, -wo.y, wo.z);
            *sampled = kBxSpecular | kBxReflection;
            *pdf = F;
            return b.kr * (F / fabsf(wi->z));
        }
        const bool entering = wo.z > 0.f;
        const float etaI = entering ? 1.f : b.ior, etaT = entering ? b.ior : 1.f;
        if (!refractDir(wo, v3(0.f, 0.f, wo.z < 0.f ? -1.f : 1.f), etaI / etaT, wi)) return v3(0.f);
        V3 ft = b.kt * (1.f - F);
        ft = ft * ((etaI * etaI) / (etaT * etaT));
        *sampled = kBxSpecular | kBxTransmission;
        *pdf = 1.f - F;
        return ft / fabsf(wi->z);
    }
    case SPB_MAT_ROUGHCONDUCTOR: if (!lobeLive(b.allowed, SPB_MAT_ROUGHCONDUCTOR)) break; {                                                // core/bxdf.cc:284-295
        if (wo.z == 0.f) return v3(0.f);
        const V3 wh = mfSample(b.mf, wo, u0, u1);
        *wi = -wo + wh * (2.f * dot(wh, wo));                                      // vect::reflect
        if (!(wo.z * wi->z > 0.f)) return v3(0.f);
        *pdf = mfPdf(b.mf, wo, wh) / (4.f * dot(wo, wh));
        return mfReflF(b, wo, *wi);
    }
    case SPB_MAT_ROUGHDIELECTRIC: if (!lobeLive(b.allowed, SPB_MAT_ROUGHDIELECTRIC)) break; {                                               // core/bxdf.cc:361-393
        if (wo.z == 0.f) return v3(0.f);
        const V3 wh = mfSample(b.mf, wo, u0, u1);
        const float co = dot(wo, wh.z > 0.f ? wh : -wh);
        const float F = frDielectric(co, 1.f, b.ior);
        if (u2 < F) {
            *wi = -wo + wh * (2.f * dot(wh, wo));
            if (!(wo.z * wi->z > 0.f)) return v3(0.f);
            *sampled = kBxReflection | kBxGlossy;
        } else {
            const float eta = co > 0.f ? 1.f / b.ior : b.ior;
            if (!refractDir(wo, wh, eta, wi)) return v3(0.f);
            if (wo.z * wi->z > 0.f) return v3(0.f);
            *sampled = kBxTransmission | kBxGlossy;
        }
        const V3 whp = wh.z >= 0.f ? wh : -wh;
        *pdf = mfTransPdf(b, wo, *wi, whp);
        return mfTransF(b, wo, *wi, whp);
    }
    case SPB_MAT_PLASTIC: if (!lobeLive(b.allowed, SPB_MAT_PLASTIC)) break; {                                                       // bsdfs/plastic.cc:48-68 (u2: its thread_local Random)
        const float Fo = frDielectric(wo.z, 1.f, b.ior);
        const float ps = plasticProbSpecular(b, Fo);
        if (u2 < ps) {
            *wi = v3(-wo.x, -wo.y, wo.z);
            *pdf = ps;
            return b.kr * (Fo / fabsf(wi->z));
        }
        *wi = cosineHemisphere(u0, u1);
        *pdf = (1.f - ps) * fabsf(wi->z) * kInvPi;
        return plasticDiffuse(b, Fo, wi->z);
    }
    case SPB_MAT_ROUGHPLASTIC: if (!lobeLive(b.allowed, SPB_MAT_ROUGHPLASTIC)) break; {                                                  // bsdfs/roughplastic.cc:49-86
        if (wo.z == 0.f) return v3(0.f);
        const V3 wh = mfSample(b.mf, wo, u0, u1);
        const float Fo = frDielectric(dot(wo, wh), 1.f, b.ior);
        const float ps = plasticProbSpecular(b, Fo);
        if (u2 < ps) {
            *wi = -wo + wh * (2.f * dot(wh, wo));
            if (!(wo.z * wi->z > 0.f)) return v3(0.f);
            *pdf = ps * mfPdf(b.mf, wo, wh) / (4.f * absDot(wo, wh));
            return b.kr * (Fo * mfD(b.mf, wh) * mfG(b.mf, wo, *wi, wh) / (4.f * wo.z * wi->z));
        }
        *wi = cosineHemisphere(u0, u1);
        *pdf = (1.f - ps) * fabsf(wi->z) * kInvPi;
        return plasticDiffuse(b, Fo, wi->z);
    }
    default: return v3(0.f);
    }
    return v3(0.f);
}

SPB_SHD V3 toLocal(const SurfacePoint& s, V3 v) { return v3(dot(s.ss, v), dot(s.ts, v), dot(s.ns, v)); }       // core/bsdf.cc:38-43
SPB_SHD V3 toWorld(const SurfacePoint& s, V3 v) { return v.x * s.ss + v.y * s.ts + v.z * s.ns; }                 // core/bsdf.cc:45-47

// BSDF::f (core/bsdf.cc:49-63)
SPB_SHD V3 bsdfF(const Bsdf& b, const SurfacePoint& s, V3 woW, V3 wiW, int type) {
    if (!(b.flags & type)) return v3(0.f);
    const bool reflect = dot(wiW, s.ns) * dot(woW, s.ns) > 0.f;
    if ((reflect && (b.flags & kBxReflection)) || (!reflect && (b.flags & kBxTransmission)))
        return lobeF(b, toLocal(s, woW), toLocal(s, wiW));
    return v3(0.f);
}
// BSDF::pdf (core/bsdf.cc:124-135)
SPB_SHD float bsdfPdf(const Bsdf& b, const SurfacePoint& s, V3 woW, V3 wiW, int type) {
    if (!(b.flags & type)) return 0.f;
    return lobePdf(b, toLocal(s, woW), toLocal(s, wiW));
}
// BSDF::sample (core/bsdf.cc:65-122) for a one-lobe BSDF
SPB_SHD V3 bsdfSample(const Bsdf& b, const SurfacePoint& s, V3 woW, float u0, float u1, float u2, int type, V3* wiW,
                      float* pdf, int* sampled) {
    *pdf = 0.f; *sampled = 0;
    if (bsdfNumComponents(b, type) == 0) return v3(0.f);
    u0 = fminf(u0, 1.0f - 5.96e-8f);
    const V3 wo = normalize(toLocal(s, woW));
    if (wo.z == 0.f) return v3(0.f);
    V3 wi;
    const V3 f = lobeSample(b, wo, u0, u1, u2, &wi, pdf, sampled);
    if (*pdf == 0.f) { *sampled = 0; return v3(0.f); }
    *wiW = normalize(toWorld(s, wi));
    return f;
}

// ---- environment map (lights/envmap.cc, core/sampling.cc:25-149, core/mipmap.cc:67-115) -------------
struct EnvMap {
    const float4* texels;      // w*h, rgb * scale
    const float*  condFunc;    // h*w        gray * sin(theta)
    const float*  condCdf;     // h*(w+1)
    const float*  condInt;     // h
    const float*  margFunc;    // h   (== condInt)
    const float*  margCdf;     // h+1
    float margInt;
    int w, h;
    float l2w[9], w2l[9];      // lightToWorld_ (= transpose of the XML matrix, envmap.cc:19) and its inverse
    float radius;
    int present;
};

#if defined(__CUDACC__)
SPB_SHD V3 mul3(const float* m, V3 v) {
    return v3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
}
__device__ __forceinline__ V3 envTexel(const EnvMap& e, int s, int t) {
    s = (s % e.w + e.w) % e.w; t = (t % e.h + e.h) % e.h;
    const float4 v = __ldg(e.texels + (size_t)t * e.w + s);
    return v3(v.x, v.y, v.z);
}
// MipMap::lookup(st, 0) -> bilinear on level 0, with the reference's truncating int cast
__device__ __forceinline__ V3 envLookup(const EnvMap& e, float s0, float t0) {
    const float s = s0 * e.w - 0.5f, t = t0 * e.h - 0.5f;
    const int si = (int)s, ti = (int)t;
    const float ds = s - si, dt = t - ti;
    return (1.f - ds) * (1.f - dt) * envTexel(e, si, ti) + ds * (1.f - dt) * envTexel(e, si + 1, ti) +
           (1.f - ds) * dt * envTexel(e, si, ti + 1) + ds * dt * envTexel(e, si + 1, ti + 1);
}
// Distribution1D::findInterval (core/sampling.cc:90-93): upper_bound - 1
__device__ __forceinline__ int findInterval(const float* cdf, int n1, float v) {
    int lo = 0, hi = n1;               // first index with cdf[i] > v
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(cdf + mid) > v) hi = mid; else lo = mid + 1; }
    int r = lo - 1;
    if (r > n1 - 1) r = n1 - 1;
    if (r < 0) r = 0;
    return r;
}
__device__ __forceinline__ float sampleDistr1D(const float* func, const float* cdf, int n, float integral, float u, float* pdf, int* off) {
    int o = findInterval(cdf, n + 1, u);
    if (o > n - 1) o = n - 1;
    *off = o;
    float du = u - __ldg(cdf + o);
    const float w = __ldg(cdf + o + 1) - __ldg(cdf + o);
    if (w > 0.f) du /= w;
    *pdf = integral > 0.f ? __ldg(func + o) / integral : 0.f;
    return (o + du) / n;
}
// Envmap::Le (lights/envmap.cc:130-134)
__device__ __forceinline__ V3 envLe(const EnvMap& e, V3 d) {
    const V3 dir = normalize(mul3(e.w2l, d));
    float phi = atan2f(dir.y, dir.x); if (phi < 0.f) phi += 2.f * kPi;
    const float theta = acosf(clampf(dir.z, -1.f, 1.f));
    return envLookup(e, phi * (0.5f * kInvPi), theta * kInvPi);
}
// Envmap::pdfLi (lights/envmap.cc:81-88)
__device__ __forceinline__ float envPdf(const EnvMap& e, V3 d) {
    const V3 dir = mul3(e.w2l, d);
    float phi = atan2f(dir.y, dir.x); if (phi < 0.f) phi += 2.f * kPi;
    const float theta = acosf(clampf(dir.z, -1.f, 1.f));
    const float st = sinf(theta);
    if (st == 0.f) return 0.f;
    int iu = (int)(phi * (0.5f * kInvPi) * e.w), iv = (int)(theta * kInvPi * e.h);
    iu = iu < 0 ? 0 : (iu > e.w - 1 ? e.w - 1 : iu);
    iv = iv < 0 ? 0 : (iv > e.h - 1 ? e.h - 1 : iv);
    return __ldg(e.condFunc + (size_t)iv * e.w + iu) / e.margInt / (2.f * kPi * kPi * st);
}
// Envmap::sampleLi (lights/envmap.cc:60-79)
__device__ __forceinline__ V3 envSample(const EnvMap& e, float u0, float u1, V3* dir, float* pdf) {
    float p0, p1; int v, uoff;
    const float d1 = sampleDistr1D(e.margFunc, e.margCdf, e.h, e.margInt, u1, &p1, &v);
    const float d0 = sampleDistr1D(e.condFunc + (size_t)v * e.w, e.condCdf + (size_t)v * (e.w + 1), e.w, __ldg(e.condInt + v), u0, &p0, &uoff);
    const float mapPdf = p0 * p1;
    *pdf = 0.f;
    if (mapPdf == 0.f) return v3(0.f);
    const float theta = d1 * kPi, phi = d0 * 2.f * kPi;
    const float ct = cosf(theta), st = sinf(theta), cp = cosf(phi), sp = sinf(phi);
    *dir = mul3(e.l2w, v3(st * cp, st * sp, ct));
    *pdf = st == 0.f ? 0.f : mapPdf / (2.f * kPi * kPi * st);
    return envLookup(e, d0, d1);
}
#endif

}  // namespace spb
