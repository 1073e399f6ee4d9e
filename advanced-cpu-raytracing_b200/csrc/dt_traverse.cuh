// Closest-hit / any-hit traversal of the two-level wide BVH (device functions).
//
// Replaces: Raytracer::IntersectObjects (raytracer.cpp:625-643), Raytracer::CastShadowRay (:585-623),
// Mesh::Intersect (mesh.cpp:158-188), InstancedMesh::Intersect (instancedMesh.cpp:16-66),
// BVH::IntersectBVH (bvh.cpp:5-31), the hit test of Mesh::IntersectFace (mesh.cpp:201-236) and
// Sphere::Intersect (sphere.cpp:13-72).
//
// Parity design: box tests against the quantised BVH8 are conservative (outward-rounded boxes, FMA with
// slack) and only select candidates; every accept/reject decision that the reference makes in float —
// per-shape slab tests, the double-precision ray transform, Cramer's rule, the sphere quadratic — is
// re-evaluated op-for-op.  Ties: the reference keeps the first hit in scan order (strict `<`), i.e. lowest
// shape index then lowest canonical face index; we replace iff (t, shape, face) is lexicographically smaller.
#pragma once
#include "dt_device.h"
#include "dt_math.cuh"

struct DtHit {
    float t;
    float beta, gamma;
    int shape;
    int face;
};

// helperMath.cpp:132-138 with m = [a b c] as columns (rows x,y,z): exact association, no contraction.
__device__ __forceinline__ float det3(v3 a, v3 b, v3 c) {
    float first = __fmul_rn(a.x, __fsub_rn(__fmul_rn(b.y, c.z), __fmul_rn(c.y, b.z)));
    float second = __fmul_rn(a.y, __fsub_rn(__fmul_rn(c.x, b.z), __fmul_rn(b.x, c.z)));
    float third = __fmul_rn(a.z, __fsub_rn(__fmul_rn(b.x, c.y), __fmul_rn(b.y, c.x)));
    return __fadd_rn(__fadd_rn(first, second), third);
}

__device__ __forceinline__ float dt_rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Mesh::IntersectFace hit test (mesh.cpp:203-236).  e1 = v0-v1, e2 = v0-v2 are stored precomputed (they are
// the same single float subtractions the reference performs per test).  `t_reject_above`: callers pass the
// distance beyond which a hit can no longer matter (best.t).
//
// Every accept decision is made on the reference's exact values (IEEE divisions, same association).  The early
// REJECTS use one approximate reciprocal (relative error < 2^-20): a quotient whose approximation is negative
// by more than a denormal is negative exactly; a barycentric sum or distance that exceeds its bound by a 1e-5
// relative margin exceeds it exactly.  Rejected candidates therefore never pay for the three IEEE divisions.
__device__ __forceinline__ bool tri_test_exact(v3 o, v3 d, v3 v0, v3 e1, v3 e2, float t_reject_above, float& t, float& beta, float& gamma) {
    const float detA = det3(e1, e2, d);
    if (detA == 0) return false;
    const float inv = dt_rcp_approx(detA);
    const v3 s = vsub(v0, o);
    const float detB = det3(s, e2, d);
    const float b_apx = __fmul_rn(detB, inv);
    if (b_apx < -1e-30f || b_apx > 1.00001f) return false;
    const float detG = det3(e1, s, d);
    const float g_apx = __fmul_rn(detG, inv);
    if (g_apx < -1e-30f || __fadd_rn(b_apx, g_apx) > 1.00001f) return false;
    const float detT = det3(e1, e2, s);
    const float t_apx = __fmul_rn(detT, inv);
    if (t_apx < -1e-30f || t_apx > __fmul_rn(t_reject_above, 1.00001f)) return false;
    // clearly inside the triangle (both quotients positive, their sum below 1 by more than the approximation error):
    // beta >= 0, gamma >= 0, beta + gamma <= 1 hold for the exact quotients as well.  Near an edge: the reference's own test.
    beta = __fdiv_rn(detB, detA);
    if (beta < 0) return false;
    gamma = __fdiv_rn(detG, detA);
    if (gamma < 0 || __fadd_rn(gamma, beta) > 1) return false;
    t = __fdiv_rn(detT, detA);
    return true;      // caller applies: t > 0 && t < minT
}

// Sphere::Intersect hit distance (sphere.cpp:13-72) for a ray already in the sphere's local space.
__device__ __forceinline__ bool sphere_test_exact(v3 o, v3 d, v3 center, float radius, float& t) {
    v3 oc = vsub(o, center);
    float c = __fsub_rn(vdot(oc, oc), __fmul_rn(radius, radius));
    float b = __fmul_rn(2.0f, vdot(d, oc));
    float a = vdot(d, d);
    float delta = __fsub_rn(__fmul_rn(b, b), __fmul_rn(__fmul_rn(4.0f, a), c));
    if (delta < 0.0f) return false;
    delta = __fsqrt_rn(delta);
    a = (float)(2.0 * (double)a);
    float t1 = __fdiv_rn(__fadd_rn(-b, delta), a);
    float t2 = __fdiv_rn(__fsub_rn(-b, delta), a);
    t = t1 < t2 ? t1 : t2;
    if (t1 < t2) { if (t1 > 0.0f) t = t1; else t = t2; }
    else if (t2 < t1) { if (t2 > 0.0f) t = t2; else t = t1; }
    return true;      // caller applies: t < minT && t > 0
}

struct DtRayPrep {
    v3 o, d;
    float idx, idy, idz;    // clamped reciprocal direction
    uint32_t oct_inv4;
};

// Reciprocal direction for the CONSERVATIVE tests only (node boxes, certificates): one MUFU, <= 1 ulp off the exact
// quotient, which the 2^-18 slack of those tests covers; every exact decision divides by the direction itself.
__device__ __forceinline__ float dt_safe_rcp(float d) {
    return fabsf(d) < 1e-30f ? copysignf(1e30f, d) : dt_rcp_approx(d);
}
__device__ __forceinline__ void dt_prep(DtRayPrep& r, v3 o, v3 d) {
    r.o = o; r.d = d;
    r.idx = dt_safe_rcp(d.x); r.idy = dt_safe_rcp(d.y); r.idz = dt_safe_rcp(d.z);
    uint32_t oct = (d.x < 0.f ? 0u : 4u) | (d.y < 0.f ? 0u : 2u) | (d.z < 0.f ? 0u : 1u);   // 7 - octant
    r.oct_inv4 = oct * 0x01010101u;
}

__device__ __forceinline__ uint32_t dt_sign_extend_s8x4(uint32_t x) {
    // each byte's top bit replicated over the byte
    return ((x >> 7) & 0x01010101u) * 0xFFu;
}
__device__ __forceinline__ uint32_t dt_byte(uint32_t w, int j) { return (w >> (j * 8)) & 0xFFu; }
// byte j of w as an exact float without the slow I2F pipe: PRMT builds 0x4B0000qq = 2^23 + q, one FADD removes 2^23.
__device__ __forceinline__ float dt_byte_f(uint32_t w, int j) {
    return __fadd_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | (uint32_t)j)), -8388608.0f);
}

#define DT_SLACK_HI 1.0000019f     // 1 + 2^-19: conservative inflation of the slab interval against FMA/rcp rounding
#define DT_SLACK_LO 0.9999981f
#define DT_SLACK_BOTH 1.0000039f   // HI / LO folded onto the far side (tn >= 0, so tn*LO <= tf*HI <=> tn <= tf*HI/LO)

#ifndef DT_NODE_V1
// byte j of w placed in mantissa bits 8..15 of 1.0f: 0x3F80qq00 = 1 + q * 2^-15, ONE PRMT and no int->float conversion.
// The plane distance q*a + b is then evaluated as m*A + B with A = 2^15 * a (exponent bump, exact) and B = b - A
// (one rounding of magnitude <= 2^-24 * 2^15 |a| = 2^-9 quantisation steps, absorbed by the 1/128-step margin the
// flattener adds when it rounds child boxes outward).
// `one` is 0x3F800000 read from the scene constants (DtSceneDev::one_bits) so that neither nvcc nor ptxas can fold it:
// PRMT takes a single immediate, and with a literal the compilers spend it on the constant and re-materialise the
// selector in a register before every one of the 48 PRMTs of a node test.
template <int J>
__device__ __forceinline__ float dt_byte_m(uint32_t w, uint32_t one) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(one), "n"(0x7604 | (J << 4)));
    return __uint_as_float(r);
}
#endif

// Blackwell packed FP32: one FFMA2 evaluates the near and the far plane distance of an axis, (qn, qf) * a + b, with the same
// round-to-nearest as two FFMAs -- the node test is issue-bound, and its 48 plane evaluations become 24 instructions.
__device__ __forceinline__ void dt_fma2(float qn, float qf, float a, float b, float& tn, float& tf) {
    unsigned long long q, aa, bb, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(qn), "f"(qf));
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(q), "l"(aa), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(tn), "=f"(tf) : "l"(r));
}

// Test the 8 quantised child boxes of a node; returns the hit mask (top 8 bits: internal children in
// traversal-priority order, low 24 bits: primitives of hit leaf children).
__device__ __forceinline__ uint32_t dt_node_hits(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4,
                                                 const DtRayPrep& r, float tmax, const uint32_t one) {
    const uint32_t e_imask = n0.w;
#ifdef DT_NODE_V1
    const float ax = __fmul_rn(__uint_as_float((e_imask & 0xFFu) << 23), r.idx);
    const float ay = __fmul_rn(__uint_as_float(((e_imask >> 8) & 0xFFu) << 23), r.idy);
    const float az = __fmul_rn(__uint_as_float(((e_imask >> 16) & 0xFFu) << 23), r.idz);
    const float bx = __fmul_rn(__fsub_rn(__uint_as_float(n0.x), r.o.x), r.idx);
    const float by = __fmul_rn(__fsub_rn(__uint_as_float(n0.y), r.o.y), r.idy);
    const float bz = __fmul_rn(__fsub_rn(__uint_as_float(n0.z), r.o.z), r.idz);
#define DT_QF(w, j) dt_byte_f(w, j)
#else
    const float ax = __fmul_rn(__uint_as_float(((e_imask & 0xFFu) + 15u) << 23), r.idx);
    const float ay = __fmul_rn(__uint_as_float((((e_imask >> 8) & 0xFFu) + 15u) << 23), r.idy);
    const float az = __fmul_rn(__uint_as_float((((e_imask >> 16) & 0xFFu) + 15u) << 23), r.idz);
    const float bx = __fmaf_rn(__fsub_rn(__uint_as_float(n0.x), r.o.x), r.idx, -ax);
    const float by = __fmaf_rn(__fsub_rn(__uint_as_float(n0.y), r.o.y), r.idy, -ay);
    const float bz = __fmaf_rn(__fsub_rn(__uint_as_float(n0.z), r.o.z), r.idz, -az);
#define DT_QF(w, j) dt_byte_m<j>(w, one)
#endif
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = dt_sign_extend_s8x4(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (r.oct_inv4 & inner_mask4)) & 0x1F1F1F1Fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z, qloz = half ? n3.y : n3.x;
        const uint32_t qhix = half ? n3.w : n3.z, qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
        const uint32_t nx = r.idx < 0.f ? qhix : qlox, fx = r.idx < 0.f ? qlox : qhix;
        const uint32_t ny = r.idy < 0.f ? qhiy : qloy, fy = r.idy < 0.f ? qloy : qhiy;
        const uint32_t nz = r.idz < 0.f ? qhiz : qloz, fz = r.idz < 0.f ? qloz : qhiz;
#ifdef DT_NODE_V1
#define DT_CHILD(j) { \
            const float tnx = __fmaf_rn(DT_QF(nx, j), ax, bx), tny = __fmaf_rn(DT_QF(ny, j), ay, by), tnz = __fmaf_rn(DT_QF(nz, j), az, bz); \
            const float tfx = __fmaf_rn(DT_QF(fx, j), ax, bx), tfy = __fmaf_rn(DT_QF(fy, j), ay, by), tfz = __fmaf_rn(DT_QF(fz, j), az, bz); \
            const float tn = __fmul_rn(fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f)), DT_SLACK_LO); \
            const float tf = __fmul_rn(fminf(fminf(tfx, tfy), fminf(tfz, tmax)), DT_SLACK_HI); \
            if (tn <= tf) hitmask |= dt_byte(child_bits4, j) << dt_byte(bit_index4, j); }
#else
        // fminf/fmaxf drop NaNs (0*inf on axis-parallel rays): a NaN constraint is ignored = conservative
#define DT_CHILD(j) { \
            float tnx, tny, tnz, tfx, tfy, tfz; \
            dt_fma2(DT_QF(nx, j), DT_QF(fx, j), ax, bx, tnx, tfx); \
            dt_fma2(DT_QF(ny, j), DT_QF(fy, j), ay, by, tny, tfy); \
            dt_fma2(DT_QF(nz, j), DT_QF(fz, j), az, bz, tnz, tfz); \
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f)); \
            const float tf = __fmul_rn(fminf(fminf(tfx, tfy), fminf(tfz, tmax)), DT_SLACK_BOTH); \
            if (tn <= tf) hitmask |= dt_byte(child_bits4, j) << dt_byte(bit_index4, j); }
#endif
        DT_CHILD(0) DT_CHILD(1) DT_CHILD(2) DT_CHILD(3)
#undef DT_CHILD
    }
#undef DT_QF
    return hitmask;
}

// Certificate for the accept-time leaf-box confirmation: true only when BoundingBox::doesIntersectWith (shape.hpp:78-100)
// is CERTAIN to pass.  The slab distances are formed with the ray's reciprocal direction (<= 1.5 ulp from the reference's
// IEEE quotients); the three conditions tmax > 0, tmax >= tmin, tmin < minT must hold with a 2^-19 relative margin.  Rays
// with a clamped reciprocal (|d| < 1e-30 on some axis), flat boxes and grazing hits fail the certificate and take the
// exact path (box_intersect_exact), so the decision is always the reference's.
__device__ __forceinline__ bool dt_leaf_box_certain(const float4 mn, const float4 mx, const DtRayPrep& r, float minT) {
    if (fabsf(r.idx) >= 1e30f || fabsf(r.idy) >= 1e30f || fabsf(r.idz) >= 1e30f) return false;
    const float x1 = __fmul_rn(__fsub_rn(mn.x, r.o.x), r.idx), x2 = __fmul_rn(__fsub_rn(mx.x, r.o.x), r.idx);
    const float y1 = __fmul_rn(__fsub_rn(mn.y, r.o.y), r.idy), y2 = __fmul_rn(__fsub_rn(mx.y, r.o.y), r.idy);
    const float z1 = __fmul_rn(__fsub_rn(mn.z, r.o.z), r.idz), z2 = __fmul_rn(__fsub_rn(mx.z, r.o.z), r.idz);
    const float tmin = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fminf(z1, z2));
    const float tmax = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fmaxf(z1, z2));
    const float m = __fmul_rn(fmaxf(fabsf(tmin), fabsf(tmax)), 1.9073486e-6f);
    return tmax > 1e-30f && tmax > m && __fsub_rn(tmax, tmin) > __fadd_rn(m, m) && __fadd_rn(tmin, m) < minT;
}

// Three-way certificate for the per-shape slab tests (Mesh::bbox, InstancedMesh::bbox): +1 the reference's test certainly
// passes, -1 it certainly fails, 0 too close to call (the caller then runs box_intersect_exact).  Same error model as above.
__device__ __forceinline__ int dt_box_certificate(const float* mn, const float* mx, v3 o, const DtRayPrep& r, float minT) {
    if (fabsf(r.idx) >= 1e30f || fabsf(r.idy) >= 1e30f || fabsf(r.idz) >= 1e30f) return 0;
    const float x1 = __fmul_rn(__fsub_rn(mn[0], o.x), r.idx), x2 = __fmul_rn(__fsub_rn(mx[0], o.x), r.idx);
    const float y1 = __fmul_rn(__fsub_rn(mn[1], o.y), r.idy), y2 = __fmul_rn(__fsub_rn(mx[1], o.y), r.idy);
    const float z1 = __fmul_rn(__fsub_rn(mn[2], o.z), r.idz), z2 = __fmul_rn(__fsub_rn(mx[2], o.z), r.idz);
    const float tmin = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fminf(z1, z2));
    const float tmax = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fmaxf(z1, z2));
    const float m = __fadd_rn(__fmul_rn(fmaxf(fabsf(tmin), fabsf(tmax)), 1.9073486e-6f), 1e-30f);
    if (!(m < 1e30f)) return 0;                                                        // NaN / overflow: undecided
    if (tmax > m && __fsub_rn(tmax, tmin) > __fadd_rn(m, m) && __fadd_rn(tmin, m) < minT) return 1;
    if (tmax < -m || __fsub_rn(tmin, tmax) > __fadd_rn(m, m) || __fsub_rn(tmin, m) > minT) return -1;
    return 0;
}
__device__ __forceinline__ bool dt_box_test(const float* mn, const float* mx, v3 o, v3 d, const DtRayPrep& r, float minT) {
    const int c = dt_box_certificate(mn, mx, o, r, minT);
    return c != 0 ? c > 0 : box_intersect_exact(mn, mx, o, d, minT);
}

// (t, shape, face) lexicographic "strictly better" — the reference's scan order with strict `<`.
__device__ __forceinline__ bool dt_better(float t, int shape, int face, const DtHit& best) {
    if (t < best.t) return true;
    if (t > best.t || !(t == best.t)) return false;
    if (best.shape < 0) return true;            // best.t == INFINITY and t == INFINITY cannot be a hit; defensive
    if (shape != best.shape) return shape < best.shape;
    return face < best.face;
}

#ifdef DT_TRAV_STATS
__device__ unsigned long long g_dt_stats[8];     // 0 nodes, 1 tri tests, 2 shape visits, 3 blas entries, 4 leaf confirms, 5 steps, 6 rays
#define DT_STAT(i) atomicAdd(&g_dt_stats[i], 1ull)
#else
#define DT_STAT(i)
#endif

// Per-ray traversal state: a resumable state machine so that persistent warps can refill finished lanes
// (dynamic fetch) instead of idling until the slowest ray of the warp is done.
struct DtTrav {
    // The world-space ray and its motion-blur time are NOT kept: the ray equals r while in the TLAS and is re-read from the
    // wave queue (one 32-byte load) when a BLAS is left; the time (queue o.w) is read only by motion-blurred shapes.
    // Seven registers that buy the kernel one more resident block per SM.
    DtRayPrep r;                // ray of the current level (world in the TLAS, local inside a BLAS)
    DtHit best;
    uint2 ng, tg;               // current node group / primitive group
    int sp, blas_sp;
    int cur_shape;              // shape whose BLAS is being traversed, -1 while in the TLAS
};
// The traversal stack is a SEPARATE local array (uint2 stack[DT_STACK_SIZE] in the kernel): with the dynamically indexed
// array inside DtTrav the whole struct lives in local memory; on its own, the scalar state above stays in registers.

template <bool ANY>
__device__ __forceinline__ void dt_trav_init(DtTrav& T, const DtSceneDev& S, v3 wo, v3 wd, float mb_time, float tmax_in) {
    T.best.t = ANY ? tmax_in : CUDART_INF_F;
    T.best.shape = -1; T.best.face = -1; T.best.beta = 0.f; T.best.gamma = 0.f;
    dt_prep(T.r, wo, wd);
    DT_STAT(6);
    T.blas_sp = 0; T.cur_shape = -1; T.sp = 0;
    // root as the single "child" of a virtual group; a scene of a few shapes skips the TLAS node test and starts with the
    // shape list (a primitive group), which is what the reference's linear scan does (raytracer.cpp:625-643)
    T.ng = S.tlas_direct > 0 ? make_uint2(0u, (1u << S.tlas_direct) - 1u) : make_uint2(0u, 0x80000000u);
    T.tg = make_uint2(0u, 0u);
}

// Visit the nearest pending child node of the current group: load 80 B, test 8 boxes, refill ng / tg.
__device__ __forceinline__ void dt_trav_node(DtTrav& T, uint2* __restrict__ stack, const DtSceneDev& S) {
    const uint32_t one = S.one_bits;
    DT_STAT(0);
    const uint32_t hits = T.ng.y;
    const uint32_t imask = T.ng.y & 0xFFu;
    const int child_bit = 31 - __clz(hits);
    T.ng.y &= ~(1u << child_bit);
    if (T.ng.y > 0x00FFFFFFu) { if (T.sp < DT_STACK_SIZE) stack[T.sp++] = T.ng; }
    const uint32_t slot = (uint32_t)(child_bit - 24) ^ (T.r.oct_inv4 & 0xFFu);
    const uint32_t rel = __popc(imask & ~(0xFFFFFFFFu << slot));
    const uint32_t ni = T.ng.x + rel;
    const uint4* np = (T.cur_shape >= 0 ? S.blas_nodes : S.tlas_nodes) + (size_t)ni * 5;
    const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
    const uint32_t hm = dt_node_hits(n0, n1, n2, n3, n4, T.r, T.best.t, one);
    T.ng.x = n1.x;
    // a group without node hits must read as empty (a remnant imask would look like a primitive group)
    T.ng.y = (hm & 0xFF000000u) ? ((hm & 0xFF000000u) | (n0.w >> 24)) : 0u;
    T.tg.x = n1.y;
    T.tg.y = hm & 0x00FFFFFFu;
}

// World -> local ray (mesh.cpp:164-170, sphere.cpp:23-30, instancedMesh.cpp:33-39).  For an identity inverseTransform the
// double-precision product reduces to x*1 + 0 + 0 + 0: the value is unchanged except that -0 becomes +0, which `x + 0.0f`
// reproduces exactly, so the 24 DMUL/DADD are skipped (the common case: untransformed meshes and spheres).
__device__ __forceinline__ void dt_to_local(const DtShapeDev* sh, v3 wo, v3 wd, const float4* __restrict__ ray_o, v3& lo, v3& ld) {
    if (sh->inv_is_identity) {
        lo = V(__fadd_rn(wo.x, 0.0f), __fadd_rn(wo.y, 0.0f), __fadd_rn(wo.z, 0.0f));
        ld = V(__fadd_rn(wd.x, 0.0f), __fadd_rn(wd.y, 0.0f), __fadd_rn(wd.z, 0.0f));
    } else {
        lo = apply_transform(sh->inv, wo, 1.0f);
        ld = apply_transform(sh->inv, wd, 0.0f);
    }
    if (sh->has_motion_blur) lo = vadd(lo, vscale(F3(sh->motion_blur), ray_o->w));       // motionBlurTime of this ray
}

// One primitive of the current primitive group.  Returns true when an ANY query is decided (occluded).
template <bool ANY>
__device__ __forceinline__ bool dt_trav_prim(DtTrav& T, uint2* __restrict__ stack, const DtSceneDev& S, const float4* __restrict__ ray_o, bool& entered_blas) {
    DtHit& best = T.best;
    const int bit = __ffs(T.tg.y) - 1;
    T.tg.y &= ~(1u << bit);
    const uint32_t prim = T.tg.x + (uint32_t)bit;
    if (T.cur_shape >= 0) {
        DT_STAT(1);
        const float4* tp = S.tris + (size_t)prim * 3;
        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        float t, beta, gamma;
        if (tri_test_exact(T.r.o, T.r.d, V(a.x, a.y, a.z), V(a.w, b.x, b.y), V(b.z, b.w, c.x), best.t, t, beta, gamma)) {
            const int face = __float_as_int(c.y);
            const bool cand = ANY ? (t > 0.0f && t < best.t) : (t > 0.0f && dt_better(t, T.cur_shape, face, best));
            if (cand) {
                // The reference only reaches this face if the float slab test of its BVH2 leaf passes (bvh.cpp:7-10);
                // every ancestor box contains the leaf box and the float slab interval is monotone in the box, so the
                // leaf test implies the ancestors'.  Runs only for would-be winners.
                DT_STAT(4);
                const float4 lb0 = __ldg(S.leaf_boxes + (size_t)prim * 2), lb1 = __ldg(S.leaf_boxes + (size_t)prim * 2 + 1);
                // minT the reference would hold when it reaches this face: it scans in (shape, face) order, so a
                // candidate that precedes the current best was tested BEFORE that best existed.
                const bool after_best = best.shape >= 0 && (T.cur_shape > best.shape || (T.cur_shape == best.shape && face > best.face));
                const float ref_min_t = ANY ? __fadd_rn(best.t, 0.01f) : (after_best ? best.t : CUDART_INF_F);     // ANY: minT = lightT + 0.01 (raytracer.cpp:580); best.t is still lightT
                bool ok = dt_leaf_box_certain(lb0, lb1, T.r, ref_min_t);
                if (!ok) {
                    const float lmn[3] = {lb0.x, lb0.y, lb0.z}, lmx[3] = {lb1.x, lb1.y, lb1.z};
                    ok = box_intersect_exact(lmn, lmx, T.r.o, T.r.d, ref_min_t);
                }
                if (ok) {
                    best.t = t; best.beta = beta; best.gamma = gamma; best.shape = T.cur_shape; best.face = face;
                    if (ANY) return true;
                }
            }
        }
        return false;
    }
    DT_STAT(2);
    const int si = __ldg(S.tlas_prims + prim);
    const DtShapeDev* sh = S.shapes + si;
    if (ANY && sh->skip_shadow) return false;
    const int kind = sh->kind;
    if (kind == DT_SHAPE_SPHERE) {
        v3 lo, ld;
        dt_to_local(sh, T.r.o, T.r.d, ray_o, lo, ld);       // TLAS level: T.r is the world ray
        float t;
        if (sphere_test_exact(lo, ld, F3(sh->center), sh->radius, t)) {
            if (ANY) {
                if (t > 0.0f && t < best.t) { best.shape = si; best.face = -1; best.t = t; return true; }
            } else if (t > 0.0f && dt_better(t, si, -1, best)) {
                best.t = t; best.beta = 0.f; best.gamma = 0.f; best.shape = si; best.face = -1;
            }
        }
        return false;
    }
    // Mesh / InstancedMesh: the reference's exact per-shape pre-tests, then descend into the BLAS.
    // ray.hitInfo.minT at the time the reference scans shape si: only hits of lower-index shapes exist.
    const float shape_min_t = ANY ? __fadd_rn(best.t, 0.01f) : ((best.shape >= 0 && si > best.shape) ? best.t : CUDART_INF_F);
    if (kind == DT_SHAPE_INSTANCE) {
        v3 so = T.r.o;
        if (sh->has_motion_blur) so = vadd(so, vscale(F3(sh->motion_blur), ray_o->w));
        if (!dt_box_test(sh->bbox_min, sh->bbox_max, so, T.r.d, T.r, shape_min_t)) return false;       // instancedMesh.cpp:29 (T.r: world ray)
    }
    v3 lo, ld;
    dt_to_local(sh, T.r.o, T.r.d, ray_o, lo, ld);
    const DtMeshDev* m = S.meshes + sh->mesh;
    // mesh.cpp:172 (Mesh::bbox) and the root node of BVH::IntersectBVH (same box) in local space
    DtRayPrep lr;
    dt_prep(lr, lo, ld);
    if (!dt_box_test(m->bbox_min, m->bbox_max, lo, ld, lr, shape_min_t)) return false;
    if (T.ng.y > 0x00FFFFFFu) { if (T.sp < DT_STACK_SIZE) stack[T.sp++] = T.ng; }
    if (T.tg.y != 0u) { if (T.sp < DT_STACK_SIZE) stack[T.sp++] = T.tg; }
    DT_STAT(3);
    T.blas_sp = T.sp;
    T.cur_shape = si;
    T.r = lr;
    T.ng = make_uint2(m->node_root, 0x80000000u);
    T.tg = make_uint2(0u, 0u);
    entered_blas = true;
    return false;
}

// One step of the state machine.  WW = false: one node visit + its primitives ("if-if");
// WW = true: descend nodes until some primitive group is pending, then drain it ("while-while").
// Returns true when the ray is finished (ANY: best.shape >= 0 <=> occluded).
// ray_o / ray_d: where the world-space ray of this traversal can be re-read (queue entry or caller's copy).
template <bool ANY, bool WW>
__device__ __forceinline__ bool dt_trav_step(DtTrav& T, uint2* __restrict__ stack, const DtSceneDev& S, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d) {
    DT_STAT(5);
    if (WW) {
        while (T.ng.y > 0x00FFFFFFu && T.tg.y == 0u) dt_trav_node(T, stack, S);
        if (T.tg.y == 0u && T.ng.y != 0u && T.ng.y <= 0x00FFFFFFu) { T.tg = T.ng; T.ng = make_uint2(0u, 0u); }
    } else {
        if (T.ng.y > 0x00FFFFFFu) dt_trav_node(T, stack, S);
        else { T.tg = T.ng; T.ng = make_uint2(0u, 0u); }
    }
    while (T.tg.y != 0u) {
        bool entered = false;
        if (dt_trav_prim<ANY>(T, stack, S, ray_o, entered)) return true;
        if (entered) break;
    }
    if (T.ng.y <= 0x00FFFFFFu && T.tg.y == 0u) {
        if (T.cur_shape >= 0 && T.sp == T.blas_sp) {
            T.cur_shape = -1;
            const float4 wo = *ray_o, wd = *ray_d;
            dt_prep(T.r, V(wo.x, wo.y, wo.z), V(wd.x, wd.y, wd.z));
        }
        if (T.sp == 0) { if (ANY) T.best.shape = -1; return true; }
        T.ng = stack[--T.sp];
    }
    return false;
}

// ANY = true: occlusion query (CastShadowRay): finishes as soon as any hit with 0 < t < tmax_in exists,
// skipping Emissive mesh shapes; best.shape >= 0 marks "occluded".
template <bool ANY>
__device__ __forceinline__ void dt_trace(const DtSceneDev& S, v3 wo, v3 wd, float mb_time, float tmax_in, DtHit& best) {
    DtTrav T;
    uint2 stack[DT_STACK_SIZE];
    dt_trav_init<ANY>(T, S, wo, wd, mb_time, tmax_in);
    const float4 ro = make_float4(wo.x, wo.y, wo.z, mb_time), rd = make_float4(wd.x, wd.y, wd.z, tmax_in);
    while (!dt_trav_step<ANY, true>(T, stack, S, &ro, &rd)) {}
    best = T.best;
}
