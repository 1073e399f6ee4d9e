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

#define DT_SLACK_BOTH 1.0000039f   // (1 + 2^-19) / (1 - 2^-19): conservative inflation of the slab interval against FMA / rcp rounding,
                                   // folded onto the far side (tn >= 0, so tn*LO <= tf*HI <=> tn <= tf*HI/LO)

// byte J of w placed in mantissa bits 8..15 of 1.0f: 0x3F80qq00 = 1 + q * 2^-15, ONE PRMT and no int->float conversion.
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

// Blackwell packed FP32: one FFMA2 evaluates the near and the far plane distance of an axis, (qn, qf) * a + b, with the same
// round-to-nearest as two FFMAs.
__device__ __forceinline__ void dt_fma2(float qn, float qf, float a, float b, float& tn, float& tf) {
    unsigned long long q, aa, bb, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(qn), "f"(qf));
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(q), "l"(aa), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(tn), "=f"(tf) : "l"(r));
}

// Test the 8 quantised child boxes of a node against the ray: bit s of the result = the box in child slot s is hit within
// [0, tmax] (conservatively).  The ALU pipe (PRMT, min/max, logic) is what this routine saturates, so the result is a plain
// slot mask -- one predicated OR per child -- and the caller derives visiting order and primitive indices from two masks.
__device__ __forceinline__ uint32_t dt_node_hits(const uint4 n0, const uint4 n2, const uint4 n3, const uint4 n4, const DtRayPrep& r, float tmax, const uint32_t one) {
    const uint32_t e_imask = n0.w;
    const float ax = __fmul_rn(__uint_as_float(((e_imask & 0xFFu) + 15u) << 23), r.idx);
    const float ay = __fmul_rn(__uint_as_float((((e_imask >> 8) & 0xFFu) + 15u) << 23), r.idy);
    const float az = __fmul_rn(__uint_as_float((((e_imask >> 16) & 0xFFu) + 15u) << 23), r.idz);
    const float bx = __fmaf_rn(__fsub_rn(__uint_as_float(n0.x), r.o.x), r.idx, -ax);
    const float by = __fmaf_rn(__fsub_rn(__uint_as_float(n0.y), r.o.y), r.idy, -ay);
    const float bz = __fmaf_rn(__fsub_rn(__uint_as_float(n0.z), r.o.z), r.idz, -az);
    uint32_t h = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z, qloz = half ? n3.y : n3.x;
        const uint32_t qhix = half ? n3.w : n3.z, qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
        const uint32_t nx = r.idx < 0.f ? qhix : qlox, fx = r.idx < 0.f ? qlox : qhix;
        const uint32_t ny = r.idy < 0.f ? qhiy : qloy, fy = r.idy < 0.f ? qloy : qhiy;
        const uint32_t nz = r.idz < 0.f ? qhiz : qloz, fz = r.idz < 0.f ? qloz : qhiz;
        // fminf/fmaxf drop NaNs (0*inf on axis-parallel rays): a NaN constraint is ignored = conservative
#define DT_CHILD(j) { \
            float tnx, tny, tnz, tfx, tfy, tfz; \
            dt_fma2(dt_byte_m<j>(nx, one), dt_byte_m<j>(fx, one), ax, bx, tnx, tfx); \
            dt_fma2(dt_byte_m<j>(ny, one), dt_byte_m<j>(fy, one), ay, by, tny, tfy); \
            dt_fma2(dt_byte_m<j>(nz, one), dt_byte_m<j>(fz, one), az, bz, tnz, tfz); \
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f)); \
            const float tf = __fmul_rn(fminf(fminf(tfx, tfy), fminf(tfz, tmax)), DT_SLACK_BOTH); \
            if (tn <= tf) h |= 1u << (half * 4 + j); }
        DT_CHILD(0) DT_CHILD(1) DT_CHILD(2) DT_CHILD(3)
#undef DT_CHILD
    }
    return h;
}

// Visiting order of the inner children: the child in slot s gets priority (s ^ oct) -- slots are assigned along the octant
// corners by the flattener, so this is front-to-back for the ray's direction octant (oct = 7 - octant).  XOR with a constant
// permutes the 8 mask bits: three conditional swaps (adjacent bits, bit pairs, nibbles).
__device__ __forceinline__ uint32_t dt_perm8(uint32_t x, uint32_t oct) {
    if (oct & 1u) x = ((x & 0x55u) << 1) | ((x >> 1) & 0x55u);
    if (oct & 2u) x = ((x & 0x33u) << 2) | ((x >> 2) & 0x33u);
    if (oct & 4u) x = ((x & 0x0Fu) << 4) | ((x >> 4) & 0x0Fu);
    return x;
}

// Certificate for the accept-time leaf-box confirmation: true only when BoundingBox::doesIntersectWith (shape.hpp:78-100)
// is CERTAIN to pass.  The slab distances are formed with the ray's reciprocal direction (<= 1.5 ulp from the reference's
// IEEE quotients); the three conditions tmax > 0, tmax >= tmin, tmin < minT must hold with a 2^-19 relative margin.  Rays
// with a clamped reciprocal (|d| < 1e-30 on some axis), flat boxes and grazing hits fail the certificate and take the
// exact path (box_intersect_exact), so the decision is always the reference's.
// Upper / lower bound of the reference's IEEE quotient given its reciprocal-based approximation v (relative error < 2^-19).
#define DT_CERT_EPS 1.9073486e-6f
__device__ __forceinline__ float dt_cert_up(float v) { return __fadd_rn(v, __fadd_rn(__fmul_rn(fabsf(v), DT_CERT_EPS), 1e-30f)); }
__device__ __forceinline__ float dt_cert_lo(float v) { return __fsub_rn(v, __fadd_rn(__fmul_rn(fabsf(v), DT_CERT_EPS), 1e-30f)); }
// Second-level pass certificate for boxes the one-margin test cannot decide because the binding near and far plane belong to the
// SAME axis -- flat boxes (axis-aligned quads as meshes, leaf boxes of axis-aligned triangles: every wall of a Cornell box) and
// rays that start on a box face.  tmax >= tmin means near_a <= far_b for every pair of axes; for a == b that holds by construction
// (near and far are min / max of the same two quotients, in the reference too), so only the six cross pairs need a margin.
// n* / f* are the per-axis near / far approximations, tmin / tmax their max / min.
__device__ __forceinline__ bool dt_cert_cross_axes(float nx, float fx, float ny, float fy, float nz, float fz, float tmin, float tmax, float minT) {
    return dt_cert_lo(tmax) > 0.0f && dt_cert_up(tmin) < minT &&
           dt_cert_up(nx) < dt_cert_lo(fminf(fy, fz)) && dt_cert_up(ny) < dt_cert_lo(fminf(fx, fz)) && dt_cert_up(nz) < dt_cert_lo(fminf(fx, fy));
}
__device__ __forceinline__ bool dt_leaf_box_certain(const float4 mn, const float4 mx, const DtRayPrep& r, float minT) {
    if (fabsf(r.idx) >= 1e30f || fabsf(r.idy) >= 1e30f || fabsf(r.idz) >= 1e30f) return false;
    const float x1 = __fmul_rn(__fsub_rn(mn.x, r.o.x), r.idx), x2 = __fmul_rn(__fsub_rn(mx.x, r.o.x), r.idx);
    const float y1 = __fmul_rn(__fsub_rn(mn.y, r.o.y), r.idy), y2 = __fmul_rn(__fsub_rn(mx.y, r.o.y), r.idy);
    const float z1 = __fmul_rn(__fsub_rn(mn.z, r.o.z), r.idz), z2 = __fmul_rn(__fsub_rn(mx.z, r.o.z), r.idz);
    const float nx = fminf(x1, x2), fx = fmaxf(x1, x2), ny = fminf(y1, y2), fy = fmaxf(y1, y2), nz = fminf(z1, z2), fz = fmaxf(z1, z2);
    const float tmin = fmaxf(fmaxf(nx, ny), nz);
    const float tmax = fminf(fminf(fx, fy), fz);
    const float m = __fmul_rn(fmaxf(fabsf(tmin), fabsf(tmax)), DT_CERT_EPS);
    if (!(m < 1e30f)) return false;
    if (tmax > 1e-30f && tmax > m && __fsub_rn(tmax, tmin) > __fadd_rn(m, m) && __fadd_rn(tmin, m) < minT) return true;
    return dt_cert_cross_axes(nx, fx, ny, fy, nz, fz, tmin, tmax, minT);
}

// Three-way certificate for the per-shape slab tests (Mesh::bbox, InstancedMesh::bbox): +1 the reference's test certainly
// passes, -1 it certainly fails, 0 too close to call (the caller then runs box_intersect_exact).  Same error model as above.
__device__ __forceinline__ int dt_box_certificate(const float* mn, const float* mx, v3 o, const DtRayPrep& r, float minT) {
    if (fabsf(r.idx) >= 1e30f || fabsf(r.idy) >= 1e30f || fabsf(r.idz) >= 1e30f) return 0;
    const float x1 = __fmul_rn(__fsub_rn(mn[0], o.x), r.idx), x2 = __fmul_rn(__fsub_rn(mx[0], o.x), r.idx);
    const float y1 = __fmul_rn(__fsub_rn(mn[1], o.y), r.idy), y2 = __fmul_rn(__fsub_rn(mx[1], o.y), r.idy);
    const float z1 = __fmul_rn(__fsub_rn(mn[2], o.z), r.idz), z2 = __fmul_rn(__fsub_rn(mx[2], o.z), r.idz);
    const float nx = fminf(x1, x2), fx = fmaxf(x1, x2), ny = fminf(y1, y2), fy = fmaxf(y1, y2), nz = fminf(z1, z2), fz = fmaxf(z1, z2);
    const float tmin = fmaxf(fmaxf(nx, ny), nz);
    const float tmax = fminf(fminf(fx, fy), fz);
    const float m = __fadd_rn(__fmul_rn(fmaxf(fabsf(tmin), fabsf(tmax)), DT_CERT_EPS), 1e-30f);
    if (!(m < 1e30f)) return 0;                                                        // NaN / overflow: undecided
    if (tmax > m && __fsub_rn(tmax, tmin) > __fadd_rn(m, m) && __fadd_rn(tmin, m) < minT) return 1;
    if (tmax < -m || __fsub_rn(tmin, tmax) > __fadd_rn(m, m) || __fsub_rn(tmin, m) > minT) return -1;
    return dt_cert_cross_axes(nx, fx, ny, fy, nz, fz, tmin, tmax, minT) ? 1 : 0;       // flat boxes, rays starting on a face
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
//
// north_star names a "shared-memory short stack": -DDT_SMEM_STACK=n keeps the n entries nearest the stack BOTTOM of every lane
// in shared memory (s[entry][lane]: conflict-free, 8 B x 128 threads x n per block) and only deeper entries in the local
// array.  Measured A/B (profiles/r2_ab_smem_stack.log) decides the default.
#ifndef DT_SMEM_STACK
#define DT_SMEM_STACK 0
#endif
struct DtStack {
    uint2* loc;                 // per-thread local array
#if DT_SMEM_STACK
    uint2* sh;                  // &smem[0][threadIdx.x]; entry k at sh[k * 128]
#endif
};
#if DT_SMEM_STACK
#define DT_DECLARE_STACK(name) uint2 name##_loc[DT_STACK_SIZE - DT_SMEM_STACK]; __shared__ uint2 name##_sh[DT_SMEM_STACK][128]; \
    DtStack name; name.loc = name##_loc; name.sh = &name##_sh[0][threadIdx.x & 127]
#else
#define DT_DECLARE_STACK(name) uint2 name##_loc[DT_STACK_SIZE]; DtStack name; name.loc = name##_loc
#endif
__device__ __forceinline__ void dt_push(const DtStack& s, int& sp, uint2 v) {
    if (sp >= DT_STACK_SIZE) return;
#if DT_SMEM_STACK
    if (sp < DT_SMEM_STACK) s.sh[sp * 128] = v; else s.loc[sp - DT_SMEM_STACK] = v;
#else
    s.loc[sp] = v;
#endif
    sp++;
}
__device__ __forceinline__ uint2 dt_pop(const DtStack& s, int& sp) {
    --sp;
#if DT_SMEM_STACK
    return sp < DT_SMEM_STACK ? s.sh[sp * 128] : s.loc[sp - DT_SMEM_STACK];
#else
    return s.loc[sp];
#endif
}

template <bool ANY>
__device__ __forceinline__ void dt_trav_init(DtTrav& T, const DtSceneDev& S, v3 wo, v3 wd, float mb_time, float tmax_in) {
    T.best.t = ANY ? tmax_in : CUDART_INF_F;
    T.best.shape = -1; T.best.face = -1; T.best.beta = 0.f; T.best.gamma = 0.f;
    dt_prep(T.r, wo, wd);
    DT_STAT(6);
    T.blas_sp = 0; T.cur_shape = -1; T.sp = 0;
    // root as the single "child" of a virtual group; a scene of a few shapes skips the TLAS node test and starts with the
    // shape list (a primitive group), which is what the reference's linear scan does (raytracer.cpp:625-643)
    T.ng = S.tlas_direct > 0 ? make_uint2(0u, ((1u << S.tlas_direct) - 1u) * 0x101u) : make_uint2(0u, 0x80000000u);       // hit slots | leaf mask << 8
    T.tg = make_uint2(0u, 0u);
}

// Visit the nearest pending child node of the current group: load 80 B, test 8 boxes, refill ng / tg.
// ORDERED: visit the hit children front to back along the ray's direction octant.  Closest-hit queries always do; occlusion
// queries take plain slot order (any hit will do, and the permutation is not free).  Measured on config 5 (profiles/
// r2_ab_anyhit_ordered_config5.log): front-to-back any-hit is 5 % slower in the wave kernels and changes nothing for the lone
// shadow rays of k_tail.
#ifndef DT_ANYHIT_ORDERED
#define DT_ANYHIT_ORDERED 0       // A/B knob: 1 = front-to-back occlusion queries in the wave kernels too
#endif
template <bool ANY, bool ORDERED = (DT_ANYHIT_ORDERED != 0) || !ANY>
__device__ __forceinline__ void dt_trav_node(DtTrav& T, const DtStack& stack, const DtSceneDev& S) {
    const uint32_t one = S.one_bits;
    DT_STAT(0);
    const uint32_t hits = T.ng.y;
    const uint32_t imask = T.ng.y & 0xFFu;
    const int child_bit = 31 - __clz(hits);
    T.ng.y &= ~(1u << child_bit);
    if (T.ng.y > 0x00FFFFFFu) dt_push(stack, T.sp, T.ng);
    // occlusion queries visit the children in plain slot order (any hit will do): no octant permutation
    const uint32_t slot = !ORDERED ? (uint32_t)(child_bit - 24) : ((uint32_t)(child_bit - 24) ^ (T.r.oct_inv4 & 0xFFu));
    const uint32_t rel = __popc(imask & ~(0xFFFFFFFFu << slot));
    const uint32_t ni = T.ng.x + rel;
    const uint4* np = (T.cur_shape >= 0 ? S.blas_nodes : S.tlas_nodes) + (size_t)ni * 5;
    const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
    const uint32_t h = dt_node_hits(n0, n2, n3, n4, T.r, T.best.t, one);
    const uint32_t im = n0.w >> 24, lm = n1.z & 0xFFu;
    const uint32_t hi = h & im, hl = h & lm;
    // node group: inner hits in visiting priority at bits 24..31 + the node's inner mask (a group without hits reads as empty);
    // primitive group: hit leaf slots at bits 0..7 + the node's leaf mask at bits 8..15 (slot -> primitive index by popcount)
    T.ng.x = n1.x;
    T.ng.y = hi ? (((!ORDERED ? hi : dt_perm8(hi, T.r.oct_inv4 & 7u)) << 24) | im) : 0u;
    T.tg.x = n1.y;
    T.tg.y = hl ? (hl | (lm << 8)) : 0u;
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

// Scan-order exactness when the visiting order differs from the reference's.  The reference scans (shape, face) in ascending
// order and tests every box with the minT it holds at that moment.  If a candidate c that PRECEDES the current best b in scan
// order arrives later with t_c a few ulps ABOVE t_b (two surfaces meeting at an edge or corner), the reference would have held
// minT = t_c when it reached b, and b's own box tests (shape box, mesh box, BVH2 leaf box: `tmin < minT`) may have rejected b
// -- flat boxes put tmin within an ulp of t_b.  This re-runs b's box tests exactly with minT = t_c.  Rare path (~1e-6 of rays).
// Out of line (it runs for ~1e-6 of the rays) and handed the four pointers it needs BY VALUE: a `const DtSceneDev&` would force
// a 264-byte local-memory copy of the kernel parameter struct in every thread of the traversal kernels (STACK 688 -> 424).
__device__ __noinline__ bool dt_best_survives_impl(const DtShapeDev* shapes, const DtMeshDev* meshes, const uint32_t* face_prim, const float4* leaf_boxes,
                                                   float best_t, int best_shape, int best_face, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, float minT) {
    struct { const DtShapeDev* shapes; const DtMeshDev* meshes; const uint32_t* face_prim; const float4* leaf_boxes; } S = {shapes, meshes, face_prim, leaf_boxes};
    struct { float t; int shape, face; } b = {best_t, best_shape, best_face};
    const DtShapeDev* sh = S.shapes + b.shape;
    if (sh->kind == DT_SHAPE_SPHERE) return true;                          // Sphere::Intersect has no box test
    const float4 o4 = *ray_o, d4 = *ray_d;
    const v3 wo = V(o4.x, o4.y, o4.z), wd = V(d4.x, d4.y, d4.z);
    if (sh->kind == DT_SHAPE_INSTANCE) {
        v3 so = wo;
        if (sh->has_motion_blur) so = vadd(so, vscale(F3(sh->motion_blur), o4.w));
        if (!box_intersect_exact(sh->bbox_min, sh->bbox_max, so, wd, minT)) return false;
    }
    v3 lo, ld;
    dt_to_local(sh, wo, wd, ray_o, lo, ld);
    const DtMeshDev* m = S.meshes + sh->mesh;
    if (!box_intersect_exact(m->bbox_min, m->bbox_max, lo, ld, minT)) return false;
    const uint32_t prim = __ldg(S.face_prim + m->face_base + (uint32_t)b.face);
    const float4 lb0 = __ldg(S.leaf_boxes + (size_t)prim * 2), lb1 = __ldg(S.leaf_boxes + (size_t)prim * 2 + 1);
    const float lmn[3] = {lb0.x, lb0.y, lb0.z}, lmx[3] = {lb1.x, lb1.y, lb1.z};
    return box_intersect_exact(lmn, lmx, lo, ld, minT);
}
__device__ __forceinline__ bool dt_best_survives(const DtSceneDev& S, const DtHit& b, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, float minT) {
    return dt_best_survives_impl(S.shapes, S.meshes, S.face_prim, S.leaf_boxes, b.t, b.shape, b.face, ray_o, ray_d, minT);
}
// c = (shape, face) precedes the best hit in the reference's scan order and lies within a few ulps behind it
__device__ __forceinline__ bool dt_close_behind(float t, int shape, int face, const DtHit& best) {
    return best.shape >= 0 && t > best.t && t <= __fmul_rn(best.t, 1.000001f) && (shape < best.shape || (shape == best.shape && face < best.face));
}

// One primitive of the current primitive group.  Returns true when an ANY query is decided (occluded).
// LIMITED: only shapes with an index below shape_limit exist (the reference's scan up to, not including, that shape: dt_stale_normal).
template <bool ANY, bool LIMITED = false>
__device__ __forceinline__ bool dt_trav_prim(DtTrav& T, const DtStack& stack, const DtSceneDev& S, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, bool& entered_blas,
                                             int shape_limit = 0x7FFFFFFF) {
    DtHit& best = T.best;
    const int bit = __ffs(T.tg.y & 0xFFu) - 1;                                   // next hit leaf slot of the group
    const uint32_t prim = T.tg.x + (uint32_t)__popc((T.tg.y >> 8) & ((1u << bit) - 1u));
    T.tg.y &= ~(1u << bit);
    if ((T.tg.y & 0xFFu) == 0u) T.tg.y = 0u;
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
            } else if (!ANY && t > 0.0f && dt_close_behind(t, T.cur_shape, face, best)) {
                // scanned BEFORE the best hit by the reference: valid on its own (minT = infinity)?  then the best hit must survive minT = t
                const float4 lb0 = __ldg(S.leaf_boxes + (size_t)prim * 2), lb1 = __ldg(S.leaf_boxes + (size_t)prim * 2 + 1);
                const float lmn[3] = {lb0.x, lb0.y, lb0.z}, lmx[3] = {lb1.x, lb1.y, lb1.z};
                if (box_intersect_exact(lmn, lmx, T.r.o, T.r.d, CUDART_INF_F) && !dt_best_survives(S, best, ray_o, ray_d, t)) {
                    best.t = t; best.beta = beta; best.gamma = gamma; best.shape = T.cur_shape; best.face = face;
                }
            }
        }
        return false;
    }
    DT_STAT(2);
    const int si = __ldg(S.tlas_prims + prim);
    if (LIMITED && si >= shape_limit) return false;
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
            } else if (t > 0.0f && dt_close_behind(t, si, -1, best) && !dt_best_survives(S, best, ray_o, ray_d, t)) {
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
    if (T.ng.y > 0x00FFFFFFu) dt_push(stack, T.sp, T.ng);
    if (T.tg.y != 0u) dt_push(stack, T.sp, T.tg);
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
template <bool ANY, bool WW, bool ORDERED = (DT_ANYHIT_ORDERED != 0) || !ANY, bool LIMITED = false>
__device__ __forceinline__ bool dt_trav_step(DtTrav& T, const DtStack& stack, const DtSceneDev& S, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d,
                                             int shape_limit = 0x7FFFFFFF) {
    DT_STAT(5);
    if (WW) {
        while (T.ng.y > 0x00FFFFFFu && T.tg.y == 0u) dt_trav_node<ANY, ORDERED>(T, stack, S);
        if (T.tg.y == 0u && T.ng.y != 0u && T.ng.y <= 0x00FFFFFFu) { T.tg = T.ng; T.ng = make_uint2(0u, 0u); }
    } else {
        if (T.ng.y > 0x00FFFFFFu) dt_trav_node<ANY, ORDERED>(T, stack, S);
        else { T.tg = T.ng; T.ng = make_uint2(0u, 0u); }
    }
    while (T.tg.y != 0u) {
        bool entered = false;
        if (dt_trav_prim<ANY, LIMITED>(T, stack, S, ray_o, ray_d, entered, shape_limit)) return true;
        if (entered) break;
    }
    if (T.ng.y <= 0x00FFFFFFu && T.tg.y == 0u) {
        if (T.cur_shape >= 0 && T.sp == T.blas_sp) {
            T.cur_shape = -1;
            const float4 wo = *ray_o, wd = *ray_d;
            dt_prep(T.r, V(wo.x, wo.y, wo.z), V(wd.x, wd.y, wd.z));
        }
        if (T.sp == 0) { if (ANY) T.best.shape = -1; return true; }
        T.ng = dt_pop(stack, T.sp);
    }
    return false;
}

// The same step in two halves for the batched-triangle variant of k_traverse_dyn (-DDT_TRI_BATCH=n, A/B): lanes whose node visit
// produced primitives PARK (dt_trav_step_nodes leaves them alone) until at least n lanes of the warp hold primitives or no lane can
// advance without testing one; dt_trav_step_prims then drains the parked groups with that many lanes live instead of the 2-4 the
// per-lane order yields.  Each lane still performs exactly the node visits and primitive tests of dt_trav_step, in the same order.
template <bool ANY>
__device__ __forceinline__ bool dt_trav_step_tail(DtTrav& T, const DtStack& stack, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d) {
    if (T.ng.y <= 0x00FFFFFFu && T.tg.y == 0u) {
        if (T.cur_shape >= 0 && T.sp == T.blas_sp) {
            T.cur_shape = -1;
            const float4 wo = *ray_o, wd = *ray_d;
            dt_prep(T.r, V(wo.x, wo.y, wo.z), V(wd.x, wd.y, wd.z));
        }
        if (T.sp == 0) { if (ANY) T.best.shape = -1; return true; }
        T.ng = dt_pop(stack, T.sp);
    }
    return false;
}
template <bool ANY, bool ORDERED = (DT_ANYHIT_ORDERED != 0) || !ANY>
__device__ __forceinline__ bool dt_trav_step_nodes(DtTrav& T, const DtStack& stack, const DtSceneDev& S, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d) {
    DT_STAT(5);
    if (T.ng.y > 0x00FFFFFFu) dt_trav_node<ANY, ORDERED>(T, stack, S);
    else { T.tg = T.ng; T.ng = make_uint2(0u, 0u); }
    return dt_trav_step_tail<ANY>(T, stack, ray_o, ray_d);
}
template <bool ANY>
__device__ __forceinline__ bool dt_trav_step_prims(DtTrav& T, const DtStack& stack, const DtSceneDev& S, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d) {
    while (T.tg.y != 0u) {
        bool entered = false;
        if (dt_trav_prim<ANY>(T, stack, S, ray_o, ray_d, entered)) return true;
        if (entered) break;
    }
    return dt_trav_step_tail<ANY>(T, stack, ray_o, ray_d);
}

// ANY = true: occlusion query (CastShadowRay): finishes as soon as any hit with 0 < t < tmax_in exists,
// skipping Emissive mesh shapes; best.shape >= 0 marks "occluded".
template <bool ANY>
__device__ __forceinline__ void dt_trace(const DtSceneDev& S, v3 wo, v3 wd, float mb_time, float tmax_in, DtHit& best) {
    DtTrav T;
    DT_DECLARE_STACK(stack);
    dt_trav_init<ANY>(T, S, wo, wd, mb_time, tmax_in);
    const float4 ro = make_float4(wo.x, wo.y, wo.z, mb_time), rd = make_float4(wd.x, wd.y, wd.z, tmax_in);
    while (!dt_trav_step<ANY, true>(T, stack, S, &ro, &rd)) {}
    best = T.best;
}
// Closest hit among the shapes [0, shape_limit) only.
__device__ __forceinline__ void dt_trace_prefix(const DtSceneDev& S, v3 wo, v3 wd, float mb_time, int shape_limit, DtHit& best) {
    DtTrav T;
    DT_DECLARE_STACK(stack);
    dt_trav_init<false>(T, S, wo, wd, mb_time, CUDART_INF_F);
    const float4 ro = make_float4(wo.x, wo.y, wo.z, mb_time), rd = make_float4(wd.x, wd.y, wd.z, CUDART_INF_F);
    while (!dt_trav_step<false, true, true, true>(T, stack, S, &ro, &rd, shape_limit)) {}
    best = T.best;
}
