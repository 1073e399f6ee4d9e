// Device restatement of the reference's per-hit surface evaluation and shading terms.
// Each function cites the reference code it follows; arithmetic types (float vs double, narrowing points)
// are kept exactly because they decide LDR rounding.
#pragma once
#include "dt_device.h"
#include "dt_math.cuh"

// ---------------------------------------------------------------- images / textures
// LDRImage::GetSample (LDRImage.h:16-26) / HDRImage::GetSample (HDRImage.h:24-33).  The reference indexes
// without bounds checks; out-of-range reads are clamped to the last texel here.
__device__ __forceinline__ v3 image_sample(const DtSceneDev& S, const DtImageDev& im, int i, int j) {
    const int ch = im.is_hdr ? 3 : im.channels;
    uint32_t idx = (uint32_t)(ch * (i + j * im.width));
    long long k = (long long)idx;
    const long long total = (long long)im.count;
    if (k > total - 3) k = total - 3;
    if (k < 0) k = 0;
    if (im.is_hdr) { const float* p = S.image_f32 + im.offset + k; return V(p[0], p[1], p[2]); }
    const uint8_t* p = S.image_u8 + im.offset + k;
    return V((float)p[0], (float)p[1], (float)p[2]);
}
// std::max(lower, std::min(n, upper)) with the std:: NaN behaviour (a NaN input yields `lower`): imageTexture.h:107-109, tonemapper.h:121-124
__device__ __forceinline__ float clipf(float n, float lo, float hi) { float m = (hi < n) ? hi : n; return (lo < m) ? m : lo; }

// ImageTexture::GetRGBSample / interpolateBilinear (imageTexture.h:60-73, 111-133); PerlinTexture returns a
// constant (perlinTexture.h:41-50).
__device__ inline v3 tex_rgb_sample(const DtSceneDev& S, const dt_texture& t, float u, float v) {
    if (t.kind == DT_TEX_PERLIN) return V(180.f, 30.f, 180.f);
    const DtImageDev im = S.images[t.image];
    if (t.interpolation == DT_INTERP_NEAREST) {
        int i = (int)(u * im.width);
        int j = (int)(v * im.height);
        if (i > im.width - 1) i = im.width - 1;
        if (j > im.height - 1) j = im.height - 1;
        return image_sample(S, im, i, j);
    }
    float i = clipf(u * im.width, 0.0f, (float)(im.width - 1));
    float j = clipf(v * im.height, 0.0f, (float)(im.height - 1));
    float p = floorf(i), q = floorf(j);
    float dx = i - p, dy = j - q;
    float w1 = (1 - dx) * (1 - dy), w2 = dx * (1 - dy), w3 = (1 - dx) * dy, w4 = dx * dy;
    v3 c = vadd(vadd(vadd(vscale(image_sample(S, im, (int)p, (int)q), w1), vscale(image_sample(S, im, (int)(p + 1), (int)q), w2)),
                     vscale(image_sample(S, im, (int)p, (int)(q + 1)), w3)), vscale(image_sample(S, im, (int)(p + 1), (int)(q + 1)), w4));
    return c;
}
__device__ __forceinline__ v3 tex_direct_sample(const DtSceneDev& S, const dt_texture& t, int i, int j) {
    if (t.kind == DT_TEX_PERLIN) return V(180.f, 30.f, 180.f);
    return image_sample(S, S.images[t.image], i, j);
}

// perlinTexture.cpp:5-38 (Ken Perlin's reference permutation; the reference stores it twice, indices < 512)
__device__ const uint8_t DT_PERM[256] = {151,160,137,91,90,15,131,13,201,95,96,53,194,233,7,225,140,36,103,30,69,142,8,99,37,240,21,10,23,
    190,6,148,247,120,234,75,0,26,197,62,94,252,219,203,117,35,11,32,57,177,33,88,237,149,56,87,174,20,125,136,171,168,68,175,
    74,165,71,134,139,48,27,166,77,146,158,231,83,111,229,122,60,211,133,230,220,105,92,41,55,46,245,40,244,102,143,54,65,25,
    63,161,1,216,80,73,209,76,132,187,208,89,18,169,200,196,135,130,116,188,159,86,164,100,109,198,173,186,3,64,52,217,226,250,
    124,123,5,202,38,147,118,126,255,82,85,212,207,206,59,227,47,16,58,17,182,189,28,42,223,183,170,213,119,248,152,2,44,154,
    163,70,221,153,101,155,167,43,172,9,129,22,39,253,19,98,108,110,79,113,224,232,178,185,112,104,218,246,97,228,251,34,242,
    193,238,210,144,12,191,179,162,241,81,51,145,235,249,14,239,107,49,192,214,31,181,199,106,157,184,84,204,176,115,121,50,45,
    127,4,150,254,138,236,205,93,222,114,67,29,24,72,243,141,128,195,78,66,215,61,156,180};
__device__ const float DT_GRAD[12][3] = {{1,1,0},{-1,1,0},{1,-1,0},{-1,-1,0},{1,0,1},{-1,0,1},{1,0,-1},{-1,0,-1},{0,1,1},{0,-1,1},{0,1,-1},{0,-1,-1}};
__device__ __forceinline__ int dt_p(int i) { return DT_PERM[i & 255]; }
__device__ __forceinline__ float dt_pdot(int g, float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(DT_GRAD[g][0], x), __fmul_rn(DT_GRAD[g][1], y)), __fmul_rn(DT_GRAD[g][2], z));
}
__device__ __forceinline__ double dt_pf(float x) {               // perlinTexture.h:153-160
    x = fabsf(x);
    if (x > 1) return 0;
    float xSqr = x * x;
    float xCube = xSqr * x;
    return (-6 * xCube * xSqr) + 15 * xCube * x - 10 * xCube + 1;
}
// PerlinTexture::GetSampleFromWorldPos (perlinTexture.h:57-123)
__device__ inline float perlin_sample(const dt_texture& t, float x, float y, float z) {
    x *= t.noise_scale; y *= t.noise_scale; z *= t.noise_scale;
    int X = (int)floorf(x), Y = (int)floorf(y), Z = (int)floorf(z);
    float dx = x - X, dy = y - Y, dz = z - Z;
    X &= 255; Y &= 255; Z &= 255;
    int ind0 = dt_p(X + dt_p(Y + dt_p(Z))) % 12;
    int ind1 = dt_p(X + dt_p(Y + dt_p(Z + 1))) % 12;
    int ind2 = dt_p(X + dt_p(Y + 1 + dt_p(Z))) % 12;
    int ind3 = dt_p(X + dt_p(Y + 1 + dt_p(Z + 1))) % 12;
    int ind4 = dt_p(X + 1 + dt_p(Y + dt_p(Z))) % 12;
    int ind5 = dt_p(X + 1 + dt_p(Y + dt_p(Z + 1))) % 12;
    int ind6 = dt_p(X + 1 + dt_p(Y + 1 + dt_p(Z))) % 12;
    int ind7 = dt_p(X + 1 + dt_p(Y + 1 + dt_p(Z + 1))) % 12;
    double c0 = dt_pdot(ind0, dx, dy, dz), c1 = dt_pdot(ind4, dx - 1, dy, dz), c2 = dt_pdot(ind2, dx, dy - 1, dz), c3 = dt_pdot(ind6, dx - 1, dy - 1, dz);
    double c4 = dt_pdot(ind1, dx, dy, dz - 1), c5 = dt_pdot(ind5, dx - 1, dy, dz - 1), c6 = dt_pdot(ind3, dx, dy - 1, dz - 1), c7 = dt_pdot(ind7, dx - 1, dy - 1, dz - 1);
    double fdx = dt_pf(dx), fdy = dt_pf(dy), fdz = dt_pf(dz), fdx1 = dt_pf(dx - 1), fdy1 = dt_pf(dy - 1), fdz1 = dt_pf(dz - 1);
    double w0 = fdx * fdy * fdz, w1 = fdx1 * fdy * fdz, w2 = fdx * fdy1 * fdz, w3 = fdx1 * fdy1 * fdz;
    double w4 = fdx * fdy * fdz1, w5 = fdx1 * fdy * fdz1, w6 = fdx * fdy1 * fdz1, w7 = fdx1 * fdy1 * fdz1;
    double total = w0 * c0 + w1 * c1 + w2 * c2 + w3 * c3 + w4 * c4 + w5 * c5 + w6 * c6 + w7 * c7;
    if (t.noise_conversion == DT_NOISE_LINEAR) return (float)((total + 1) / 2.0f);
    return (float)fabs(total);
}
__device__ __forceinline__ float tex_world_sample(const dt_texture& t, float x, float y, float z) {
    return t.kind == DT_TEX_PERLIN ? perlin_sample(t, x, y, z) : 0.0f;          // texture.h:47-49
}
__device__ __forceinline__ float tex_width(const DtSceneDev& S, const dt_texture& t) { return t.kind == DT_TEX_PERLIN ? CUDART_INF_F : (float)S.images[t.image].width; }
__device__ __forceinline__ float tex_height(const DtSceneDev& S, const dt_texture& t) { return t.kind == DT_TEX_PERLIN ? CUDART_INF_F : (float)S.images[t.image].height; }

// ---------------------------------------------------------------- final-hit surface evaluation
__device__ __forceinline__ float floor_tiled(float x) {          // Mesh::GetFloorForTiledUV (mesh.cpp:382-389)
    if (x > 1.0001f) {
        x = x - floorf(x);
        if (x < 0.0001) x = 1.0f;
    }
    return x;
}
// Mesh::GetTangentAndBitangentForTriangle (mesh.cpp:390-422)
__device__ inline void tangent_bitangent(v3 vert0, v3 vert1, v3 vert2, float2 uv0, float2 uv1, float2 uv2, v3& tan, v3& bitan) {
    v3 e1 = vunit(vsub(vert1, vert0));
    v3 e2 = vunit(vsub(vert2, vert1));
    float v0u = floor_tiled(uv0.x), v0v = floor_tiled(uv0.y);
    float v1u = floor_tiled(uv1.x), v1v = floor_tiled(uv1.y);
    float v2u = floor_tiled(uv2.x), v2v = floor_tiled(uv2.y);
    float u1 = v1u - v0u, v1 = v1v - v0v, u2 = v2u - v1u, v2 = v2v - v1v;
    float det = 1.0f / (u1 * v2 - v1 * u2);
    tan.x = det * (v2 * e1.x - v1 * e2.x);
    tan.y = det * (v2 * e1.y - v1 * e2.y);
    tan.z = det * (v2 * e1.z - v1 * e2.z);
    bitan.x = -det * u2 * e1.x + det * u1 * e2.x;
    bitan.y = -det * u2 * e1.y + det * u1 * e2.y;
    bitan.z = -det * u2 * e1.z + det * u1 * e2.z;
    tan = vunit(tan);
    bitan = vunit(bitan);
}
// GetTransformedNormal (helperMath.cpp:86-109): double 3x3 * 3x1, accumulated from 0.0f
__device__ __forceinline__ v3 transformed_normal(v3 tan, v3 bitan, v3 normal, v3 s) {
    double x = 0.0, y = 0.0, z = 0.0;
    x += (double)tan.x * (double)s.x; x += (double)bitan.x * (double)s.y; x += (double)normal.x * (double)s.z;
    y += (double)tan.y * (double)s.x; y += (double)bitan.y * (double)s.y; y += (double)normal.y * (double)s.z;
    z += (double)tan.z * (double)s.x; z += (double)bitan.z * (double)s.y; z += (double)normal.z * (double)s.z;
    return vunit(V((float)x, (float)y, (float)z));
}

struct DtSurface { v3 normal; float u, v; };

// Everything Mesh::IntersectFace does after accepting a hit (mesh.cpp:237-372) plus the caller's part
// (Mesh::Intersect mesh.cpp:176-180 / InstancedMesh::Intersect instancedMesh.cpp:52-58), evaluated once for
// the winning hit.  (lo, ld) is the ray in the shape's local space.
__device__ inline void mesh_surface(const DtSceneDev& S, const DtShapeDev& sh, int face_index, float t, float beta, float gama,
                                    v3 lo, v3 ld, bool smooth, DtSurface& out) {
    const DtShapeDev& ow = S.shapes[sh.owner];
    const DtMeshDev& m = S.meshes[sh.mesh];
    const DtFaceDev fc = S.faces[m.face_base + face_index];
    v3 N = V(fc.nx, fc.ny, fc.nz);
    if (smooth && m.normal_base >= 0) {           // DT_FLAG_SMOOTH_SHADING: barycentric interpolation of the vertex normals
        const float* np = S.vnormals + (size_t)m.normal_base * 3;
        const v3 n0 = F3(np + (size_t)fc.v0 * 3), n1 = F3(np + (size_t)fc.v1 * 3), n2 = F3(np + (size_t)fc.v2 * 3);
        const float w0 = 1.0f - beta - gama;
        const v3 sn = V(n0.x * w0 + n1.x * beta + n2.x * gama, n0.y * w0 + n1.y * beta + n2.y * gama, n0.z * w0 + n1.z * beta + n2.z * gama);
        const float l = vlen(sn);
        if (l > 0.0f) N = vdiv(sn, l);
    }
    v3 normal = N;
    out.u = 0.f; out.v = 0.f;
    if (m.n_uvs > 0) {
        const float* vp = S.verts + (size_t)m.vert_base * 3;
        const float2* up = (const float2*)(S.uvs + (size_t)m.uv_base * 2);
        const int toff = m.texture_offset - m.vertex_offset;
        float2 uv0 = up[fc.v0 + toff], uv1 = up[fc.v1 + toff], uv2 = up[fc.v2 + toff];
        float u = uv0.x + beta * (uv1.x - uv0.x) + gama * (uv2.x - uv0.x);
        float v = uv0.y + beta * (uv1.y - uv0.y) + gama * (uv2.y - uv0.y);
        u = floor_tiled(u); v = floor_tiled(v);
        out.u = u; out.v = v;
        if (ow.tex_normal >= 0) {
            v3 v0 = F3(vp + (size_t)fc.v0 * 3), v1 = F3(vp + (size_t)fc.v1 * 3), v2 = F3(vp + (size_t)fc.v2 * 3);
            v3 s = tex_rgb_sample(S, S.textures[ow.tex_normal], u, v);
            s = vsub(vdiv(s, 127.5f), V(1.f, 1.f, 1.f));
            s = vunit(s);
            v3 tan, bitan;
            tangent_bitangent(v0, v1, v2, uv0, uv1, uv2, tan, bitan);
            normal = transformed_normal(tan, bitan, N, s);
            normal = vunit(apply_transform(ow.invT, normal, 0.0f));
        } else if (ow.tex_bump >= 0) {
            const dt_texture bm = S.textures[ow.tex_bump];
            v3 v0 = F3(vp + (size_t)fc.v0 * 3), v1 = F3(vp + (size_t)fc.v1 * 3), v2 = F3(vp + (size_t)fc.v2 * 3);
            v3 tan, bitan;
            tangent_bitangent(v0, v1, v2, uv0, uv1, uv2, tan, bitan);
            if (bm.kind == DT_TEX_PERLIN) {
                v3 g;
                float eps = 0.001;
                v3 p = vadd(lo, vscale(ld, t));                   // local-space hit point (mesh.cpp:241)
                float bf = bm.sample_multiplier;
                float hxyz = tex_world_sample(bm, p.x, p.y, p.z) * bf;
                g.x = (tex_world_sample(bm, p.x + eps, p.y, p.z) * bf - hxyz) / eps;
                g.y = (tex_world_sample(bm, p.x, p.y + eps, p.z) * bf - hxyz) / eps;
                g.z = (tex_world_sample(bm, p.x, p.y, p.z + eps) * bf - hxyz) / eps;
                v3 gpar = vscale(N, vdot(g, N));
                v3 sg = vsub(g, gpar);
                normal = vunit(vsub(N, sg));
                normal = vunit(apply_transform(ow.invT, normal, 0.0f));
            } else {
                float width = tex_width(S, bm), height = tex_height(S, bm);
                int i = (int)(u * (width - 1));
                int j = (int)(v * (height - 1));
                int nextI = i + 1, nextJ = j + 1;
                if (i == width - 1) nextI = i;
                if (j == height - 1) nextJ = j;
                v3 c0 = tex_direct_sample(S, bm, i, j), c1 = tex_direct_sample(S, bm, nextI, j), c2 = tex_direct_sample(S, bm, i, nextJ);
                float h_uv = (c0.x + c0.y + c0.z) / 3.0f;
                float hDeltaU = (c1.x + c1.y + c1.z) / 3.0f;
                float hDeltaV = (c2.x + c2.y + c2.z) / 3.0f;
                float bumpFactor = bm.sample_multiplier;
                v3 q_u = vadd(tan, vscale(N, ((hDeltaU - h_uv) * bumpFactor)));
                v3 q_v = vadd(bitan, vscale(N, ((hDeltaV - h_uv) * bumpFactor)));
                v3 nn = vcross(q_v, q_u);
                normal = vunit(nn);
                if (nn.x * N.x <= 0 && nn.y * N.y <= 0 && nn.z * N.z <= 0) normal = vscale(normal, -1.f);
                else if (fabsf(nn.y - N.y) > 0.9f || fabsf(nn.x - N.x) > 0.9f || fabsf(nn.z - N.z) > 0.9f) normal = vscale(normal, -1.f);
                normal = vunit(apply_transform(ow.invT, normal, 0.0f));
            }
        }
    } else {
        normal = vunit(apply_transform(ow.invT, normal, 0.0f));
    }
    // Mesh::Intersect :179 (same matrix again for a Mesh) / InstancedMesh::Intersect :57 (the instance's matrix)
    out.normal = vunit(apply_transform(sh.invT, normal, 0.0f));
}

// Sphere::Intersect after the hit is accepted (sphere.cpp:72-178).  stale_normal: what hitInfo.normal held
// before (the reference's normal-map branch leaves it untouched, sphere.cpp:95-115).
__device__ inline void sphere_surface(const DtSceneDev& S, const DtShapeDev& sh, float t, v3 lo, v3 ld, v3 stale_normal, DtSurface& out) {
    v3 center = F3(sh.center);
    float radius = sh.radius;
    v3 localhit = vadd(lo, vscale(ld, t));
    v3 p = vsub(localhit, center);
    float phi = atan2f(p.z, p.x);
    float theta = acosf(p.y / radius);
    float u = (float)((-phi + DT_PI) / (2.0f * DT_PI));
    float v = (float)(theta / DT_PI);
    out.u = u; out.v = v;
    v3 normal = stale_normal;
    if (sh.tex_normal >= 0) {
        // only reads the texture in the reference; the normal is left as it was
    } else if (sh.tex_bump >= 0) {
        const dt_texture bm = S.textures[sh.tex_bump];
        v3 tan, bitan;                                            // sphere.cpp:181-193
        tan.x = (float)(2 * DT_PI * p.z); tan.y = 0; tan.z = (float)(-2 * DT_PI * p.x);
        bitan.x = (float)(DT_PI * p.y * cosf(phi)); bitan.y = (float)(-radius * DT_PI * sinf(theta)); bitan.z = (float)(DT_PI * p.y * sinf(phi));
        tan = vunit(tan); bitan = vunit(bitan);
        v3 N = vunit(vcross(bitan, tan));
        if (bm.kind == DT_TEX_PERLIN) {
            v3 g;
            float eps = 0.001;
            float hxyz = tex_world_sample(bm, p.x, p.y, p.z);
            g.x = (tex_world_sample(bm, p.x + eps, p.y, p.z) - hxyz) / eps;
            g.y = (tex_world_sample(bm, p.x, p.y + eps, p.z) - hxyz) / eps;
            g.z = (tex_world_sample(bm, p.x, p.y, p.z + eps) - hxyz) / eps;
            v3 gpar = vscale(N, vdot(g, N));
            v3 sg = vsub(g, gpar);
            normal = vunit(vsub(N, sg));
        } else {
            float width = tex_width(S, bm), height = tex_height(S, bm);
            int i = (int)(u * width);
            int j = (int)(v * height);
            float normalizer = bm.normalizer, bumpFactor = bm.sample_multiplier;
            v3 a = vdiv(tex_direct_sample(S, bm, i + 1, j), normalizer);
            v3 b = vdiv(tex_direct_sample(S, bm, i, j), normalizer);
            v3 c = vdiv(tex_direct_sample(S, bm, i, j + 1), normalizer);
            float h1 = (a.x + a.y + a.z) * bumpFactor;
            float h_uv = (b.x + b.y + b.z) * bumpFactor;
            float h2 = (c.x + c.y + c.z) * bumpFactor;
            v3 q_u = vadd(tan, vscale(N, (h1 - h_uv)));
            v3 q_v = vadd(bitan, vscale(N, (h2 - h_uv)));
            normal = vunit(vcross(q_v, q_u));
        }
    } else {
        normal = vunit(vsub(localhit, center));
    }
    out.normal = vunit(apply_transform(sh.invT, normal, 0.0f));
}

// ---------------------------------------------------------------- BRDFs (brdf*.cpp)
// The reference evaluates its lobes through degrees: cos(rad(deg(acosf(c)))) in double around a FLOAT acos (std::acos(float),
// helperMath.cpp:156) and pow(., exponent).  Two algebraic shortcuts stay within that float acos' own rounding (<= 1e-7 absolute
// in the cosine; the LDR tolerance is 4e-3, and the LDR images are byte-identical on every scene rendered both ways) and remove
// the bulk of the shading kernel's double-precision transcendental work on path-traced frames (three BRDF evaluations per hit):
//  * cos_deg(angle_between_unit(a, b)) is the clamped dot product itself (the degree <-> radian factors are inverses);
//  * pow(x, e) with an integer exponent (every shipped scene; <Exponent> is parsed as a number) is repeated squaring in double.
// The `theta_i >= 90` cut-off is taken on the cosine: (float)(acos(c) * 180/pi) >= 90.0f  <=>  c <= 6.657903e-08f (computed
// by bisection over all floats).  The two "Original" variants still divide by the cosine of the float-rounded angle: kept as is.
#define DT_COS_90_CUTOFF 6.657903e-08f
__device__ __forceinline__ double dt_cos_between(v3 a, v3 b) { return (double)fminf(1.0f, fmaxf(-1.0f, vdot(a, b))); }
__device__ __forceinline__ double dt_pow_lobe(double x, double e) {
    if (e >= 0.0 && e <= 65536.0 && e == floor(e)) {
        unsigned n = (unsigned)e;
        double r = 1.0, p = x;
        while (n) { if (n & 1u) r *= p; p *= p; n >>= 1; }
        return r;
    }
    return pow(x, e);
}

__device__ inline v3 brdf_apply(const DtSceneDev& S, const dt_material& mat, v3 kd, v3 ks, v3 w_i, v3 w_o, v3 n) {
    const dt_brdf b = S.brdfs[mat.brdf];
    const float exponent = b.exponent;
    if (b.kind == DT_BRDF_MODIFIED_PHONG || b.kind == DT_BRDF_MODIFIED_BLINN_PHONG || b.kind == DT_BRDF_TORRANCE_SPARROW) {
        if (fminf(1.0f, fmaxf(-1.0f, vdot(w_i, n))) <= DT_COS_90_CUTOFF) return V(0, 0, 0);          // angleTheta_i >= 90.0f
    }
    switch (b.kind) {
        case DT_BRDF_MODIFIED_PHONG: {                                         // brdfModifiedPhong.cpp:14-33
            v3 pr = vunit(vsub(vscale(vscale(n, 2.0f), vdot(n, w_i)), w_i));
            const double cosTerm = dt_pow_lobe(dt_cos_between(pr, w_o), (double)exponent);
            if (b.flag) {
                v3 kdTerm = vscale(kd, (float)(1.0f / DT_PI));
                double cons = (exponent + 2) / (2 * DT_PI);
                return vadd(kdTerm, vscale(ks, (float)(cons * cosTerm)));
            }
            return vadd(kd, vscale(ks, (float)cosTerm));
        }
        case DT_BRDF_MODIFIED_BLINN_PHONG: {                                   // brdfModifiedBlinnPhong.cpp:11-29
            v3 s = vadd(w_i, w_o);
            v3 half = vdiv(s, vlen(s));
            const double cosTerm = dt_pow_lobe(dt_cos_between(half, n), (double)exponent);
            if (b.flag) {
                v3 kdTerm = vscale(kd, (float)(1.0f / DT_PI));
                double cons = (exponent + 8) / (8 * DT_PI);
                return vadd(kdTerm, vscale(ks, (float)(cons * cosTerm)));
            }
            return vadd(kd, vscale(ks, (float)cosTerm));
        }
        case DT_BRDF_TORRANCE_SPARROW: {                                       // brdfTorranceSparrow.cpp:15-59
            v3 s = vadd(w_i, w_o);
            v3 half = vdiv(s, vlen(s));
            double e = exponent;
            double d = (e + 2) * dt_pow_lobe((double)vdot(half, n), e) / (2 * DT_PI);
            double ri = mat.refractive_index;
            double r0 = ((ri - 1) * (ri - 1)) / ((ri + 1) * (ri + 1));                                // pow(x, 2.0) is x*x, correctly rounded
            const double om = 1.0 - (double)vdot(half, w_o), om2 = om * om;
            double f = r0 + (1.0 - r0) * (om2 * om2 * om);                                            // pow(., 5.0)
            double ndoth = vdot(n, half), ndotwo = vdot(n, w_o), ndotwi = vdot(n, w_i), wodoth = vdot(w_o, half);
            double g = fmin(1.0, fmin(2.0f * ndoth * ndotwo / wodoth, 2.0 * ndoth * ndotwi / wodoth));
            double kdCoeff = (1.0f / DT_PI);
            if (b.flag) kdCoeff *= (1 - f);
            v3 kdTerm = vscale(kd, (float)kdCoeff);
            double costheta = vdot(n, w_i);
            double cosphi = vdot(n, w_o);
            v3 ksTerm = vscale(ks, (float)((d * f * g) / (4 * costheta * cosphi)));
            return vadd(kdTerm, ksTerm);
        }
        default: break;
    }
    const float angleTheta_i = (float)angle_between_unit(w_i, n);
    switch (b.kind) {
        case DT_BRDF_PHONG: {                                                  // brdfPhong.cpp:11-20
            if (angleTheta_i >= 90.0f || angleTheta_i < 0) return V(0, 0, 0);
            v3 pr = vunit(vsub(vscale(vscale(n, 2.0f), vdot(n, w_i)), w_i));
            double angleR = angle_between_unit(pr, w_o);
            return vadd(kd, vscale(ks, (float)(pow(cos_deg(angleR), (double)exponent) / cos_deg((double)angleTheta_i))));
        }
        case DT_BRDF_BLINN_PHONG: {                                            // brdfBlinnPhong.cpp:11-20
            if (angleTheta_i >= 90.0f) return V(0, 0, 0);
            v3 s = vadd(w_i, w_o);
            v3 half = vdiv(s, vlen(s));
            double a = angle_between_unit(half, n);
            return vadd(kd, vscale(ks, (float)(pow(cos_deg(a), (double)exponent) / cos_deg((double)angleTheta_i))));
        }
        default: return V(0, 0, 0);
    }
}

// Raytracer::Get{Diffuse,Specular}ReflectanceCoeff (raytracer.cpp:478-539).  Both read shape->diffuseTex.
__device__ inline v3 reflectance_coeff(const DtSceneDev& S, const DtShapeDev& sh, const dt_material& mat, v3 hitPoint, float u, float v, bool specular) {
    v3 reflectance = specular ? F3(mat.specular) : F3(mat.diffuse);
    const bool has = specular ? (sh.tex_specular >= 0) : (sh.tex_diffuse >= 0);
    if (!has || sh.tex_diffuse < 0) return reflectance;
    const dt_texture t = S.textures[sh.tex_diffuse];
    v3 tk;
    if (t.kind == DT_TEX_PERLIN) {
        float s = tex_world_sample(t, hitPoint.x, hitPoint.y, hitPoint.z);
        tk = V(s, s, s);
    } else {
        tk = vdiv(tex_rgb_sample(S, t, u, v), 255.0f);
    }
    if (t.decal_mode == DT_DECAL_BLEND_KD) reflectance = vdiv(vadd(tk, F3(mat.diffuse)), 2.0f);
    else reflectance = tk;
    return reflectance;
}

// Raytracer::Shade with the incoming radiance factored out (raytracer.cpp:192-206, 540-554):
//   BRDF material:  res * Li * cos_i          -> returns res (cos applied by caller as in the reference order)
//   otherwise:      kd*Li*cos + ks*Li*pow(..)
// Li is passed in so the float evaluation order is the reference's.
// kd / ks depend on the hit only (material + diffuse texture): evaluated once per hit, not once per light.
__device__ __forceinline__ v3 shade_term(const dt_material& mat, const DtSceneDev& S, v3 kd, v3 ks, v3 normal, v3 w_i, v3 w_o, v3 Li, v3* brdf_res) {
    if (mat.brdf >= 0) {
        float costheta_i = fmaxf(0.0f, vdot(w_i, normal));
        v3 res = brdf_apply(S, mat, kd, ks, w_i, w_o, normal);
        if (brdf_res) *brdf_res = res;
        return vscale(vmul(res, Li), costheta_i);
    }
    float costheta = fmaxf(0.0f, vdot(w_i, normal));
    v3 diffuse = vscale(vmul(kd, Li), costheta);
    v3 s = vadd(w_i, w_o);
    v3 half = vdiv(s, vlen(s));
    float cosAlpha = fmaxf(0.0f, vdot(normal, half));
    v3 spec = vscale(vmul(ks, Li), powf(cosAlpha, mat.phong_exponent));
    return vadd(diffuse, spec);
}

// SphericalEnvironmentLight::GetSample (sphericalEnvironmentLight.h:22-35)
__device__ inline v3 env_sample(const DtSceneDev& S, int li, v3 dir) {
    const DtImageDev im = S.images[S.env_lights[li].image];
    float u = (float)((1 + (atan2f(dir.x, -dir.z) / DT_PI)) / 2.0f);
    float v = (float)(acosf(dir.y) / DT_PI);
    int i = (int)(im.width * u);
    int j = (int)(im.height * v);
    v3 s = image_sample(S, im, i, j);
    return vscale(vscale(s, 2.f), DT_PI_F);
}

// SpotLight::GetIrradiance (spotLight.h:33-57)
__device__ inline v3 spot_irradiance(const dt_spot_light& l, v3 point) {
    v3 pos = F3(l.pos);
    float dist = vlen(vsub(point, pos));
    v3 toPoint = vdiv(vsub(point, pos), dist);
    double alpha = angle_between_unit(F3(l.dir), toPoint);
    if (alpha <= 0 || alpha > (l.coverage_angle / 2.0f)) return V(0, 0, 0);
    float distSqr = dist * dist;
    v3 irr = vdiv(F3(l.intensity), distSqr);
    if (alpha > (l.falloff_angle / 2.0f)) {
        double cosAlpha = cos(alpha * (DT_PI / 180.0f));
        double s = pow((cosAlpha - l.cos_half_coverage) / (l.cos_half_falloff - l.cos_half_coverage), 4.0);
        irr = vscale(irr, (float)s);
    }
    return irr;
}

// Raytracer::Reflect (raytracer.cpp:424-440)
__device__ inline v3 reflect_dir(DtRng& rng, v3 normal, v3 w_o, float roughness) {
    v3 r = vunit(vsub(vscale(vscale(normal, 2.0f), vdot(normal, w_o)), w_o));
    if (roughness > 0.001) {
        v3 u, v;
        orthonormal_basis(r, u, v);
        float psi1 = rng01(rng) - 0.5f;
        float psi2 = rng01(rng) - 0.5f;
        return vunit(vadd(r, vscale(vadd(vscale(u, psi1), vscale(v, psi2)), roughness)));
    }
    return r;
}
