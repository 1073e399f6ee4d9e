// Device restatement of the reference's math layer (src/helperMath.{h,cpp}, src/matrix.hpp).
// The translation unit is compiled with -fmad=false: every float/double expression below is evaluated
// op-for-op like the reference's SSE2 scalar build (no FMA contraction), which is what makes primary-hit
// `t` values bit-identical (SURVEY.md 8a "Numerics contract").  Explicit __f*_rn intrinsics are used in the
// parity-critical routines as a second line of defence.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#define DT_PI 3.14159265358979323846           /* M_PI (double)              */
#define DT_PI_F ((float)DT_PI)                 /* M_PI narrowed to float     */

struct v3 { float x, y, z; };

__device__ __forceinline__ v3 V(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ v3 F3(const float* p) { return V(p[0], p[1], p[2]); }
__device__ __forceinline__ v3 vadd(v3 a, v3 b) { return V(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
__device__ __forceinline__ v3 vsub(v3 a, v3 b) { return V(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ v3 vmul(v3 a, v3 b) { return V(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z)); }
__device__ __forceinline__ v3 vscale(v3 a, float s) { return V(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
// IEEE-exact x / y.  A ZERO numerator (components of axis-aligned normals, black colour channels: nine divisions of an average
// config-5 hit) would send the whole warp through the out-of-line special-case path of the division (7 % of k_shade's issued
// instructions, profiles/r2p_ncu_config5_shade_raw.csv); 0 / y is the product's signed zero whenever y is finite and non-zero.
__device__ __forceinline__ float dt_fdiv(float x, float y) {
    if (x == 0.0f && fabsf(y) > 0.0f && fabsf(y) < CUDART_INF_F) return __fmul_rn(x, y);
    return __fdiv_rn(x, y);
}
__device__ __forceinline__ v3 vdiv(v3 a, float s) { return V(dt_fdiv(a.x, s), dt_fdiv(a.y, s), dt_fdiv(a.z, s)); }
__device__ __forceinline__ v3 vneg(v3 a) { return V(__fmul_rn(a.x, -1.0f), __fmul_rn(a.y, -1.0f), __fmul_rn(a.z, -1.0f)); }   // helperMath.h:41-47
// helperMath.cpp:54-58: a.x*b.x + a.y*b.y + a.z*b.z, left to right
__device__ __forceinline__ float vdot(v3 a, v3 b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
// helperMath.cpp:110-116
__device__ __forceinline__ v3 vcross(v3 a, v3 b) {
    return V(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)),
             __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
             __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
// helperMath.cpp:118-130
__device__ __forceinline__ float vlen(v3 a) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)), __fmul_rn(a.z, a.z)));
}
__device__ __forceinline__ v3 vunit(v3 a) { float l = vlen(a); return V(dt_fdiv(a.x, l), dt_fdiv(a.y, l), dt_fdiv(a.z, l)); }

// helperMath.cpp:59-85
__device__ __forceinline__ void orthonormal_basis(v3 r, v3& u, v3& v) {
    float ax = fabsf(r.x), ay = fabsf(r.y), az = fabsf(r.z);
    v3 rp = r;
    if (ax < ay) { if (ax < az) rp.x = 1.0f; else rp.z = 1.0f; }
    else { if (ay < az) rp.y = 1.0f; else rp.z = 1.0f; }
    u = vunit(vcross(rp, r));
    v = vunit(vcross(r, u));
}

// matrix.hpp:86-121: rows 0..2 of a double 4x4 applied to (v, w); each row summed left to right in double
// without contraction, then rounded to float.
__device__ __forceinline__ float xf_row(const double* t, v3 v, float w) {
    double s = __dmul_rn(t[0], (double)v.x);
    s = __dadd_rn(s, __dmul_rn(t[1], (double)v.y));
    s = __dadd_rn(s, __dmul_rn(t[2], (double)v.z));
    s = __dadd_rn(s, __dmul_rn(t[3], (double)w));
    return (float)s;
}
__device__ __forceinline__ v3 apply_transform(const double* t, v3 v, float w) {
    return V(xf_row(t, v, w), xf_row(t + 4, v, w), xf_row(t + 8, v, w));
}

// helperMath.cpp:154-161
__device__ __forceinline__ double angle_between_unit(v3 a, v3 b) {
    float d = vdot(a, b);
    float c = fminf(1.0f, fmaxf(-1.0f, d));
    return (double)acosf(c) * (180.0f / DT_PI);       // std::acos(float) is the FLOAT overload in the reference; only the scaling is double
}
__device__ __forceinline__ double cos_deg(double a) { return cos(a * (DT_PI / 180.0f)); }

// shape.hpp:78-100 — BoundingBox::doesIntersectWith, exact float semantics (IEEE divide, x-axis compare-swap,
// NaN-dropping fmin/fmax on y/z).
__device__ __forceinline__ bool box_intersect_exact(const float* mn, const float* mx, v3 o, v3 d, float minT) {
    float tx1 = __fdiv_rn(__fsub_rn(mn[0], o.x), d.x);
    float tx2 = __fdiv_rn(__fsub_rn(mx[0], o.x), d.x);
    float tmin = tx1, tmax = tx2;
    if (tx1 > tx2) { tmin = tx2; tmax = tx1; }
    float ty1 = __fdiv_rn(__fsub_rn(mn[1], o.y), d.y);
    float ty2 = __fdiv_rn(__fsub_rn(mx[1], o.y), d.y);
    tmin = fmaxf(tmin, fminf(ty1, ty2));
    tmax = fminf(tmax, fmaxf(ty1, ty2));
    float tz1 = __fdiv_rn(__fsub_rn(mn[2], o.z), d.z);
    float tz2 = __fdiv_rn(__fsub_rn(mx[2], o.z), d.z);
    tmin = fmaxf(tmin, fminf(tz1, tz2));
    tmax = fminf(tmax, fmaxf(tz1, tz2));
    return tmax > 0 && tmax >= tmin && tmin < minT;
}

// Counter-based RNG (the reference's std::mt19937s are unseeded and raced by 8 threads; any i.i.d. U[0,1)
// stream is distributionally equivalent).  PCG-style hash of (key, counter) -> 24-bit uniform float.
__device__ __forceinline__ uint32_t dt_hash(uint32_t a, uint32_t b) {
    uint32_t h = a * 0x9E3779B1u + b * 0x85EBCA77u + 0xC2B2AE3Du;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}
struct DtRng { uint32_t key, ctr; };
__device__ __forceinline__ float rng01(DtRng& r) {
    uint32_t h = dt_hash(r.key, r.ctr++);
    return (float)(h >> 8) * (1.0f / 16777216.0f);      // [0,1)
}
