/*
 * dt_oracle.c — CPU restatement of the reference's render hot path.   *** TEST INFRASTRUCTURE ONLY ***
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product path (advanced-cpu-raytracing_b200/) never does and has no CPU fallback.
 *
 * It restates, function by function, the algorithm of dorukb/Advanced-CPU-Raytracing on the flat
 * dt_scene_desc of include/dorktracer.h (the same description the CUDA path consumes), following the
 * reference's recursion and float/double evaluation order.  Each function cites the reference file:line it
 * follows.  Compile with -ffp-contract=off (the reference build contains no FMA).
 *
 * Pinning: tests/test_cpu_oracle_host.py checks this file against the reference's own golden
 * PNGs (archive/hw1_outputs, six "pins" scenes, committed as fixtures under tests/golden/) and against outputs of the
 * compiled reference oracle/_ref/raytracer_probe (hit ids, radiance bits, LDR bytes, ray counts; fixtures + live in the
 * build container); tests/test_cpu_monte_carlo_pin.py does the same for the Monte-Carlo path (reference-RNG mode below),
 * tests/test_cpu_smooth_shading.py for DT_FLAG_SMOOTH_SHADING (off: the reference's image; on: the course's goldens).
 *
 * Random numbers: the reference draws from std::mt19937s that are default-seeded or seeded from an UNSEEDED rand()
 * (raytracer.cpp:10,14, main.cpp:49, areaLight.h:28, meshLight.h:17, sphericalEnvironmentLight.h:19), so with ONE render
 * thread it is deterministic.  Two modes:
 *   dto_render                every pixel owns a SplitMix64 stream keyed by (seed, pixel): thread-count independent, used for
 *                             statistical comparisons with the GPU path (whose RNG is counter-based too);
 *   dto_render_reference_rng  replays the reference's generators -- glibc rand() from its default seed, mt19937, libstdc++'s
 *                             generate_canonical / uniform_real_distribution / uniform_int_distribution -- in the reference's
 *                             call order on one thread, so that a Monte-Carlo render is BIT-IDENTICAL (radiance bits and ray
 *                             counts) to `DT_THREADS=1 oracle/_ref/raytracer_probe` (tests/test_cpu_oracle_host.py).
 */
#include "dorktracer.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define RAD2DEG (180.0f / M_PI)
#define DEG2RAD (M_PI / 180.0f)

typedef struct { float x, y, z; } v3;

/* ---- helperMath.cpp:4-162 ---- */
static inline v3 V(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vdiv(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline v3 vneg(v3 a) { return V(a.x * -1.0f, a.y * -1.0f, a.z * -1.0f); }
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 vcross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float vlen(v3 a) { return sqrtf((a.x * a.x) + (a.y * a.y) + (a.z * a.z)); }
static inline v3 vunit(v3 a) { float l = vlen(a); return V(a.x / l, a.y / l, a.z / l); }
static inline v3 F3(const float* p) { return V(p[0], p[1], p[2]); }

static float determinant(float m[3][3]) {                      /* helperMath.cpp:132-138 */
    float firstTerm = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]);
    float secondTerm = m[1][0] * (m[0][2] * m[2][1] - m[0][1] * m[2][2]);
    float thirdTerm = m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]);
    return firstTerm + secondTerm + thirdTerm;
}
static void orthonormal_basis(v3 r, v3* u, v3* v) {            /* helperMath.cpp:59-85 */
    float ax = fabsf(r.x), ay = fabsf(r.y), az = fabsf(r.z);
    v3 rp = r;
    if (ax < ay) { if (ax < az) rp.x = 1.0f; else rp.z = 1.0f; }
    else { if (ay < az) rp.y = 1.0f; else rp.z = 1.0f; }
    *u = vunit(vcross(rp, r));
    *v = vunit(vcross(r, *u));
}
static double angle_between_unit(v3 a, v3 b) {                 /* helperMath.cpp:154-157 */
    float d = vdot(a, b);
    float c = fminf(1.0f, fmaxf(-1.0f, d));
    return acosf(c) * RAD2DEG;             /* std::acos(float) is the FLOAT overload; the product with RAD2DEG (a double) is double */
}
static double cos_deg(double a) { return cos(a * DEG2RAD); }   /* helperMath.cpp:158-161 */

/* matrix.hpp:86-121: double 4x4 (row-major) applied to (v, w), result rounded to float */
static v3 apply_transform(const double* t, v3 v, float w) {
    v3 r;
    r.x = (float)(t[0] * v.x + t[1] * v.y + t[2] * v.z + t[3] * w);
    r.y = (float)(t[4] * v.x + t[5] * v.y + t[6] * v.z + t[7] * w);
    r.z = (float)(t[8] * v.x + t[9] * v.y + t[10] * v.z + t[11] * w);
    return r;
}
static v3 transformed_normal(v3 tan, v3 bitan, v3 normal, v3 s) {   /* helperMath.cpp:86-109 (double 3x3 * 3x1) */
    double x = 0.0f, y = 0.0f, z = 0.0f;
    x += (double)tan.x * (double)s.x; x += (double)bitan.x * (double)s.y; x += (double)normal.x * (double)s.z;
    y += (double)tan.y * (double)s.x; y += (double)bitan.y * (double)s.y; y += (double)normal.y * (double)s.z;
    z += (double)tan.z * (double)s.x; z += (double)bitan.z * (double)s.y; z += (double)normal.z * (double)s.z;
    return vunit(V((float)x, (float)y, (float)z));
}

/* ---- ray.hpp:10-32 ---- */
typedef struct {
    int hasHit;
    int matId;
    float minT;
    v3 normal, hitPoint;
    float u, v;
    int shape;       /* index into desc->shapes (Shape* hitShape) */
    int face;        /* probe only: canonical face index of the accepted face */
} HitInfo;
typedef struct {
    v3 origin, dir;
    HitInfo hit;
    float n_medium;  /* refractiveIndexOfCurrentMedium */
    float mbTime;
    v3 throughput;
} Ray;

/* ---- the reference's generators (reference-RNG mode) ----
 * std::mt19937 = mersenne_twister_engine<uint_fast32_t, 32, 624, 397, 31, 0x9908b0df, 11, 0xffffffff, 7, 0x9d2c5680, 15,
 * 0xefc60000, 18, 1812433253>, default seed 5489. */
typedef struct { uint32_t mt[624]; int idx; } Mt19937;
static void mt_seed(Mt19937* m, uint32_t seed) {
    m->mt[0] = seed;
    for (int i = 1; i < 624; i++) m->mt[i] = 1812433253u * (m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) + (uint32_t)i;
    m->idx = 624;
}
static uint32_t mt_next(Mt19937* m) {
    if (m->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (m->mt[i] & 0x80000000u) | (m->mt[(i + 1) % 624] & 0x7FFFFFFFu);
            m->mt[i] = m->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->idx = 0;
    }
    uint32_t y = m->mt[m->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
/* libstdc++ std::generate_canonical<double, 53>(mt19937&) (bits/random.tcc): two 32-bit draws, low word first */
static double mt_canonical(Mt19937* m) {
    double sum = (double)mt_next(m);
    sum += (double)mt_next(m) * 4294967296.0;
    double ret = sum / 18446744073709551616.0;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret;
}
/* libstdc++ std::uniform_int_distribution<int>(0, n - 1)(mt19937&) (bits/uniform_int_dist.h, GCC >= 11): the generator's
 * range is exactly 32 bits, so Lemire's nearly-divisionless method with a 64-bit product (_S_nd<uint64_t>) */
static uint32_t mt_uniform_int(Mt19937* m, uint32_t n) {
    uint64_t product = (uint64_t)mt_next(m) * (uint64_t)n;
    uint32_t low = (uint32_t)product;
    if (low < n) {
        uint32_t threshold = (0u - n) % n;
        while (low < threshold) { product = (uint64_t)mt_next(m) * (uint64_t)n; low = (uint32_t)product; }
    }
    return (uint32_t)(product >> 32);
}
/* glibc rand() = random() TYPE_3 (additive feedback, x^31 + x^3 + 1) from the default seed 1: the reference never calls srand */
typedef struct { uint32_t r[34 + 310 + 64]; int next; } GlibcRand;
static void glibc_rand_init(GlibcRand* g, uint32_t seed) {
    int32_t* r = (int32_t*)g->r;
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) { r[i] = (int32_t)((16807LL * r[i - 1]) % 2147483647); if (r[i] < 0) r[i] += 2147483647; }
    for (int i = 31; i < 34; i++) g->r[i] = g->r[i - 31];
    for (int i = 34; i < 34 + 310 + 64; i++) g->r[i] = g->r[i - 31] + g->r[i - 3];
    g->next = 344;
}
static uint32_t glibc_rand(GlibcRand* g) { return g->r[g->next++] >> 1; }      /* 64 values are plenty (one per generator) */

#define DTO_MAX_LIGHT_STREAMS 16
typedef struct {
    Mt19937 main_gen;                          /* main.cpp:49    std::mt19937 randomGen(rand()) of the (single) render thread        */
    Mt19937 rand_gen;                          /* raytracer.cpp:10  Raytracer::randGen = mt19937(rand())                             */
    Mt19937 dof_gen;                           /* raytracer.cpp:14  dofLensSampleGenerator = mt19937(rand())                         */
    Mt19937 area[DTO_MAX_LIGHT_STREAMS];       /* areaLight.h:28    sampleGen = mt19937()                                            */
    Mt19937 mesh[DTO_MAX_LIGHT_STREAMS];       /* meshLight.h:17    sampleGen = mt19937()                                            */
    Mt19937 env[DTO_MAX_LIGHT_STREAMS];        /* sphericalEnvironmentLight.h:19  gen = mt19937(rand()), seeded while parsing        */
} RefRng;
enum { RS_MAIN, RS_RAND, RS_DOF, RS_AREA, RS_MESH, RS_ENV };

typedef struct {
    const dt_scene_desc* sc;
    const dt_camera_desc* cam;
    uint64_t rng;
    uint64_t n_closest, n_shadow;
    RefRng* ref;     /* non-NULL: reference-RNG mode */
} Ctx;

/* SplitMix64 -> double in [0,1) (uniform_real_distribution<double> over mt19937 has 53 random bits too) */
static inline double rnd01(Ctx* c) {
    uint64_t z = (c->rng += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
/* One canonical double in [0,1) from the generator the reference uses at this call site (reference-RNG mode) or from the
 * pixel's SplitMix64 stream.  std::uniform_real_distribution<double>(a, b)(g) = generate_canonical(g) * (b - a) + a. */
static inline double rnd_s(Ctx* c, int stream, int idx) {
    if (!c->ref) return rnd01(c);
    RefRng* r = c->ref;
    if (idx >= DTO_MAX_LIGHT_STREAMS) idx = DTO_MAX_LIGHT_STREAMS - 1;
    switch (stream) {
        case RS_MAIN: return mt_canonical(&r->main_gen);
        case RS_RAND: return mt_canonical(&r->rand_gen);
        case RS_DOF: return mt_canonical(&r->dof_gen);
        case RS_AREA: return mt_canonical(&r->area[idx]);
        case RS_MESH: return mt_canonical(&r->mesh[idx]);
        default: return mt_canonical(&r->env[idx]);
    }
}
static inline float rnd_normalized(Ctx* c) { return (float)(0.0f + (1.0f - 0.0f) * rnd_s(c, RS_RAND, 0)); }   /* raytracer.cpp:24-28 */

/* ---- shape.hpp:78-100 ---- */
static int box_intersect(const float* mn, const float* mx, const Ray* ray) {
    float tx1 = (mn[0] - ray->origin.x) / ray->dir.x;
    float tx2 = (mx[0] - ray->origin.x) / ray->dir.x;
    float tmin = tx1, tmax = tx2;
    if (tx1 > tx2) { tmin = tx2; tmax = tx1; }
    float ty1 = (mn[1] - ray->origin.y) / ray->dir.y;
    float ty2 = (mx[1] - ray->origin.y) / ray->dir.y;
    tmin = fmaxf(tmin, fminf(ty1, ty2));
    tmax = fminf(tmax, fmaxf(ty1, ty2));
    float tz1 = (mn[2] - ray->origin.z) / ray->dir.z;
    float tz2 = (mx[2] - ray->origin.z) / ray->dir.z;
    tmin = fmaxf(tmin, fminf(tz1, tz2));
    tmax = fminf(tmax, fmaxf(tz1, tz2));
    return tmax > 0 && tmax >= tmin && tmin < ray->hit.minT;
}

/* ---- images / textures ---- */
static v3 image_sample(const dt_image* im, int i, int j) {      /* LDRImage.h:16-26, HDRImage.h:24-33 */
    int64_t total = (int64_t)im->width * im->height * (im->is_hdr ? 3 : im->channels);
    uint32_t idx = (uint32_t)((im->is_hdr ? 3 : im->channels) * (i + j * im->width));
    int64_t k = idx;
    if (k > total - 3) k = total - 3;     /* the reference reads out of bounds here (UB); we clamp */
    if (k < 0) k = 0;
    if (im->is_hdr) { const float* p = (const float*)im->data; return V(p[k], p[k + 1], p[k + 2]); }
    const uint8_t* p = (const uint8_t*)im->data;
    return V((float)p[k], (float)p[k + 1], (float)p[k + 2]);
}
/* std::max(lower, std::min(n, upper)) with the std:: NaN behaviour: min(a,b) = (b<a)?b:a, max(a,b) = (a<b)?b:a, so a NaN
   input yields `lower` (imageTexture.h:107-109, tonemapper.h:121-124) */
static float clipf(float n, float lo, float hi) { float m = (hi < n) ? hi : n; return (lo < m) ? m : lo; }

static v3 tex_rgb_sample(const dt_scene_desc* sc, const dt_texture* t, float u, float v) {
    if (t->kind == DT_TEX_PERLIN) return V(180, 30, 180);        /* perlinTexture.h:46-50 */
    const dt_image* im = &sc->images[t->image];
    if (t->interpolation == DT_INTERP_NEAREST) {                 /* imageTexture.h:60-73 */
        int i = (int)(u * im->width);
        int j = (int)(v * im->height);
        if (i > im->width - 1) i = im->width - 1;
        if (j > im->height - 1) j = im->height - 1;
        return image_sample(im, i, j);
    }
    /* interpolateBilinear, imageTexture.h:111-133 */
    float i = clipf(u * im->width, 0.0f, (float)(im->width - 1));
    float j = clipf(v * im->height, 0.0f, (float)(im->height - 1));
    float p = floorf(i), q = floorf(j);
    float dx = i - p, dy = j - q;
    float w1 = (1 - dx) * (1 - dy), w2 = dx * (1 - dy), w3 = (1 - dx) * dy, w4 = dx * dy;
    v3 c = vadd(vadd(vadd(vscale(image_sample(im, (int)p, (int)q), w1), vscale(image_sample(im, (int)(p + 1), (int)q), w2)),
                     vscale(image_sample(im, (int)p, (int)(q + 1)), w3)), vscale(image_sample(im, (int)(p + 1), (int)(q + 1)), w4));
    return c;
}
static v3 tex_direct_sample(const dt_scene_desc* sc, const dt_texture* t, int i, int j) {
    if (t->kind == DT_TEX_PERLIN) return V(180, 30, 180);
    return image_sample(&sc->images[t->image], i, j);
}
static float tex_width(const dt_scene_desc* sc, const dt_texture* t) { return t->kind == DT_TEX_PERLIN ? INFINITY : (float)sc->images[t->image].width; }
static float tex_height(const dt_scene_desc* sc, const dt_texture* t) { return t->kind == DT_TEX_PERLIN ? INFINITY : (float)sc->images[t->image].height; }

/* perlinTexture.cpp:5-38 (Ken Perlin's reference permutation, doubled) */
static const int PERM[256] = {151,160,137,91,90,15,131,13,201,95,96,53,194,233,7,225,140,36,103,30,69,142,8,99,37,240,21,10,23,
    190,6,148,247,120,234,75,0,26,197,62,94,252,219,203,117,35,11,32,57,177,33,88,237,149,56,87,174,20,125,136,171,168,68,175,
    74,165,71,134,139,48,27,166,77,146,158,231,83,111,229,122,60,211,133,230,220,105,92,41,55,46,245,40,244,102,143,54,65,25,
    63,161,1,216,80,73,209,76,132,187,208,89,18,169,200,196,135,130,116,188,159,86,164,100,109,198,173,186,3,64,52,217,226,250,
    124,123,5,202,38,147,118,126,255,82,85,212,207,206,59,227,47,16,58,17,182,189,28,42,223,183,170,213,119,248,152,2,44,154,
    163,70,221,153,101,155,167,43,172,9,129,22,39,253,19,98,108,110,79,113,224,232,178,185,112,104,218,246,97,228,251,34,242,
    193,238,210,144,12,191,179,162,241,81,51,145,235,249,14,239,107,49,192,214,31,181,199,106,157,184,84,204,176,115,121,50,45,
    127,4,150,254,138,236,205,93,222,114,67,29,24,72,243,141,128,195,78,66,215,61,156,180};
static inline int P(int i) { return PERM[i & 255]; }   /* p[512] is p[256] twice; indices stay < 512 */
static const float GRAD[12][3] = {{1,1,0},{-1,1,0},{1,-1,0},{-1,-1,0},{1,0,1},{-1,0,1},{1,0,-1},{-1,0,-1},{0,1,1},{0,-1,1},{0,1,-1},{0,-1,-1}};
static inline float pdot(int g, float x, float y, float z) { return GRAD[g][0] * x + GRAD[g][1] * y + GRAD[g][2] * z; }
static double pf(float x) {                                      /* perlinTexture.h:153-160 */
    x = fabsf(x);
    if (x > 1) return 0;
    float xSqr = x * x;
    float xCube = xSqr * x;
    return (-6 * xCube * xSqr) + 15 * xCube * x - 10 * xCube + 1;
}
static float perlin_sample(const dt_texture* t, float x, float y, float z) {   /* perlinTexture.h:57-123 */
    x *= t->noise_scale; y *= t->noise_scale; z *= t->noise_scale;
    int X = (int)floorf(x), Y = (int)floorf(y), Z = (int)floorf(z);
    float dx = x - X, dy = y - Y, dz = z - Z;
    X &= 255; Y &= 255; Z &= 255;
    int ind0 = P(X + P(Y + P(Z))) % 12;
    int ind1 = P(X + P(Y + P(Z + 1))) % 12;
    int ind2 = P(X + P(Y + 1 + P(Z))) % 12;
    int ind3 = P(X + P(Y + 1 + P(Z + 1))) % 12;
    int ind4 = P(X + 1 + P(Y + P(Z))) % 12;
    int ind5 = P(X + 1 + P(Y + P(Z + 1))) % 12;
    int ind6 = P(X + 1 + P(Y + 1 + P(Z))) % 12;
    int ind7 = P(X + 1 + P(Y + 1 + P(Z + 1))) % 12;
    double c0 = pdot(ind0, dx, dy, dz), c1 = pdot(ind4, dx - 1, dy, dz), c2 = pdot(ind2, dx, dy - 1, dz), c3 = pdot(ind6, dx - 1, dy - 1, dz);
    double c4 = pdot(ind1, dx, dy, dz - 1), c5 = pdot(ind5, dx - 1, dy, dz - 1), c6 = pdot(ind3, dx, dy - 1, dz - 1), c7 = pdot(ind7, dx - 1, dy - 1, dz - 1);
    double fdx = pf(dx), fdy = pf(dy), fdz = pf(dz), fdx1 = pf(dx - 1), fdy1 = pf(dy - 1), fdz1 = pf(dz - 1);
    double w0 = fdx * fdy * fdz, w1 = fdx1 * fdy * fdz, w2 = fdx * fdy1 * fdz, w3 = fdx1 * fdy1 * fdz;
    double w4 = fdx * fdy * fdz1, w5 = fdx1 * fdy * fdz1, w6 = fdx * fdy1 * fdz1, w7 = fdx1 * fdy1 * fdz1;
    double total = w0 * c0 + w1 * c1 + w2 * c2 + w3 * c3 + w4 * c4 + w5 * c5 + w6 * c6 + w7 * c7;
    if (t->noise_conversion == DT_NOISE_LINEAR) return (float)((total + 1) / 2.0f);
    return (float)fabs(total);
}
static float tex_world_sample(const dt_texture* t, float x, float y, float z) {
    return t->kind == DT_TEX_PERLIN ? perlin_sample(t, x, y, z) : 0.0f;         /* texture.h:47-49 */
}

/* ---- mesh.cpp:382-422 ---- */
static float floor_tiled(float x) {
    if (x > 1.0001f) {
        x = x - floorf(x);
        if (x < 0.0001) x = 1.0f;
    }
    return x;
}
static void tangent_bitangent(v3 vert0, v3 vert1, v3 vert2, const float* uv0, const float* uv1, const float* uv2, v3* tan, v3* bitan) {
    v3 e1 = vunit(vsub(vert1, vert0));
    v3 e2 = vunit(vsub(vert2, vert1));
    float v0u = floor_tiled(uv0[0]), v0v = floor_tiled(uv0[1]);
    float v1u = floor_tiled(uv1[0]), v1v = floor_tiled(uv1[1]);
    float v2u = floor_tiled(uv2[0]), v2v = floor_tiled(uv2[1]);
    float u1 = v1u - v0u, v1 = v1v - v0v, u2 = v2u - v1u, v2 = v2v - v1v;
    float det = 1.0f / (u1 * v2 - v1 * u2);
    tan->x = det * (v2 * e1.x - v1 * e2.x);
    tan->y = det * (v2 * e1.y - v1 * e2.y);
    tan->z = det * (v2 * e1.z - v1 * e2.z);
    bitan->x = -det * u2 * e1.x + det * u1 * e2.x;
    bitan->y = -det * u2 * e1.y + det * u1 * e2.y;
    bitan->z = -det * u2 * e1.z + det * u1 * e2.z;
    *tan = vunit(*tan);
    *bitan = vunit(*bitan);
}
static inline float greyscale3(v3 c) { return (c.x + c.y + c.z) / 3.0f; }     /* mesh.cpp:195-197 */

static inline v3 mesh_vertex(const dt_mesh* m, int id) { const float* p = &m->vertices[(size_t)(id - 1 + m->vertex_offset) * 3]; return V(p[0], p[1], p[2]); }
static inline const float* mesh_uv(const dt_mesh* m, int id) { return &m->uvs[(size_t)(id - 1 + m->texture_offset) * 2]; }

static int g_dto_smooth = 0;                               /* dto_set_render_flags: DT_FLAG_SMOOTH_SHADING */
static int g_dto_origin_leak = 0;                          /* dto_set_render_flags: DTO_FLAG_ORIGIN_LEAK (oracle only, see instance_intersect) */

/* ---- Mesh::IntersectFace, mesh.cpp:201-372.  `owner` = index of the Mesh shape that owns the geometry ---- */
static int intersect_face(const dt_scene_desc* sc, Ray* ray, int owner, uint32_t faceIdx) {
    const dt_shape* sh = &sc->shapes[owner];
    const dt_mesh* m = &sc->meshes[sh->mesh];
    const dt_face* face = &m->faces[faceIdx];
    v3 v0 = mesh_vertex(m, face->v0_id), v1 = mesh_vertex(m, face->v1_id), v2 = mesh_vertex(m, face->v2_id);
    v3 o = ray->origin, d = ray->dir;

    float matrixA[3][3] = {{v0.x - v1.x, v0.x - v2.x, d.x}, {v0.y - v1.y, v0.y - v2.y, d.y}, {v0.z - v1.z, v0.z - v2.z, d.z}};
    float detA = determinant(matrixA);
    if (detA == 0) return 0;
    float matrixBeta[3][3] = {{v0.x - o.x, v0.x - v2.x, d.x}, {v0.y - o.y, v0.y - v2.y, d.y}, {v0.z - o.z, v0.z - v2.z, d.z}};
    float beta = determinant(matrixBeta) / detA;
    if (beta < 0) return 0;
    float matrixGama[3][3] = {{v0.x - v1.x, v0.x - o.x, d.x}, {v0.y - v1.y, v0.y - o.y, d.y}, {v0.z - v1.z, v0.z - o.z, d.z}};
    float gama = determinant(matrixGama) / detA;
    if (gama < 0 || gama + beta > 1) return 0;
    float matrixT[3][3] = {{v0.x - v1.x, v0.x - v2.x, v0.x - o.x}, {v0.y - v1.y, v0.y - v2.y, v0.y - o.y}, {v0.z - v1.z, v0.z - v2.z, v0.z - o.z}};
    float t = determinant(matrixT) / detA;
    if (!(t > 0.0f && t < ray->hit.minT)) return 0;

    ray->hit.minT = t;
    ray->hit.hasHit = 1;
    v3 N = F3(face->n);
    if (g_dto_smooth && m->vertex_normals) {               /* DT_FLAG_SMOOTH_SHADING (not in the reference): interpolated vertex normals */
        const float* n0 = &m->vertex_normals[(size_t)(face->v0_id - 1 + m->vertex_offset) * 3];
        const float* n1 = &m->vertex_normals[(size_t)(face->v1_id - 1 + m->vertex_offset) * 3];
        const float* n2 = &m->vertex_normals[(size_t)(face->v2_id - 1 + m->vertex_offset) * 3];
        const float w0 = 1.0f - beta - gama;
        v3 sn = V(n0[0] * w0 + n1[0] * beta + n2[0] * gama, n0[1] * w0 + n1[1] * beta + n2[1] * gama, n0[2] * w0 + n1[2] * beta + n2[2] * gama);
        const float l = vlen(sn);
        if (l > 0.0f) N = vdiv(sn, l);
    }
    ray->hit.normal = N;
    ray->hit.hitPoint = vadd(ray->origin, vscale(ray->dir, ray->hit.minT));
    const double* invT = sh->inverse_transpose_transform;
    if (m->n_uvs > 0) {
        const float* uv0 = mesh_uv(m, face->v0_id); const float* uv1 = mesh_uv(m, face->v1_id); const float* uv2 = mesh_uv(m, face->v2_id);
        float u = uv0[0] + beta * (uv1[0] - uv0[0]) + gama * (uv2[0] - uv0[0]);
        float v = uv0[1] + beta * (uv1[1] - uv0[1]) + gama * (uv2[1] - uv0[1]);
        u = floor_tiled(u); v = floor_tiled(v);
        ray->hit.u = u; ray->hit.v = v;
        if (sh->tex_normal >= 0) {
            const dt_texture* nm = &sc->textures[sh->tex_normal];
            v3 s = tex_rgb_sample(sc, nm, u, v);
            s = vsub(vdiv(s, 127.5f), V(1, 1, 1));
            s = vunit(s);
            v3 tan, bitan;
            tangent_bitangent(v0, v1, v2, uv0, uv1, uv2, &tan, &bitan);
            ray->hit.normal = transformed_normal(tan, bitan, N, s);
            ray->hit.normal = vunit(apply_transform(invT, ray->hit.normal, 0.0f));
        } else if (sh->tex_bump >= 0) {
            const dt_texture* bm = &sc->textures[sh->tex_bump];
            v3 tan, bitan;
            tangent_bitangent(v0, v1, v2, uv0, uv1, uv2, &tan, &bitan);
            if (bm->kind == DT_TEX_PERLIN) {
                v3 g;
                float eps = 0.001;
                v3 p = ray->hit.hitPoint;
                float bf = bm->sample_multiplier;
                float hxyz = tex_world_sample(bm, p.x, p.y, p.z) * bf;
                g.x = (tex_world_sample(bm, p.x + eps, p.y, p.z) * bf - hxyz) / eps;
                g.y = (tex_world_sample(bm, p.x, p.y + eps, p.z) * bf - hxyz) / eps;
                g.z = (tex_world_sample(bm, p.x, p.y, p.z + eps) * bf - hxyz) / eps;
                v3 gpar = vscale(N, vdot(g, N));
                v3 sg = vsub(g, gpar);
                ray->hit.normal = vunit(vsub(N, sg));
                ray->hit.normal = vunit(apply_transform(invT, ray->hit.normal, 0.0f));
            } else {
                float width = tex_width(sc, bm), height = tex_height(sc, bm);
                int i = (int)(u * (width - 1));
                int j = (int)(v * (height - 1));
                int nextI = i + 1, nextJ = j + 1;
                if (i == width - 1) nextI = i;
                if (j == height - 1) nextJ = j;
                float h_uv = greyscale3(tex_direct_sample(sc, bm, i, j));
                float hDeltaU = greyscale3(tex_direct_sample(sc, bm, nextI, j));
                float hDeltaV = greyscale3(tex_direct_sample(sc, bm, i, nextJ));
                float bumpFactor = bm->sample_multiplier;
                v3 q_u = vadd(tan, vscale(N, ((hDeltaU - h_uv) * bumpFactor)));
                v3 q_v = vadd(bitan, vscale(N, ((hDeltaV - h_uv) * bumpFactor)));
                v3 nn = vcross(q_v, q_u);
                ray->hit.normal = vunit(nn);
                if (nn.x * N.x <= 0 && nn.y * N.y <= 0 && nn.z * N.z <= 0) ray->hit.normal = vscale(ray->hit.normal, -1);
                else if (fabsf(nn.y - N.y) > 0.9f || fabsf(nn.x - N.x) > 0.9f || fabsf(nn.z - N.z) > 0.9f) ray->hit.normal = vscale(ray->hit.normal, -1);
                ray->hit.normal = vunit(apply_transform(invT, ray->hit.normal, 0.0f));
            }
        }
    } else {
        ray->hit.normal = vunit(apply_transform(invT, ray->hit.normal, 0.0f));
    }
    ray->hit.matId = sh->material;
    ray->hit.shape = owner;
    ray->hit.face = (int)faceIdx;
    return 1;
}

/* ---- BVH::IntersectBVH, bvh.cpp:5-31 (iterative left-then-right DFS == the reference's recursion order) ---- */
static int intersect_bvh(const dt_scene_desc* sc, Ray* ray, int owner) {
    const dt_mesh* m = &sc->meshes[sc->shapes[owner].mesh];
    int stack[256];
    int sp = 0, hasHit = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const dt_bvh2_node* nd = &m->bvh[stack[--sp]];
        if (!box_intersect(nd->bmin, nd->bmax, ray)) continue;
        if (nd->left < 0 && nd->right < 0 && nd->face_count > 0) {
            for (uint32_t i = nd->first_face; i < nd->first_face + nd->face_count; i++)
                if (intersect_face(sc, ray, owner, i)) hasHit = 1;
        } else if (nd->left >= 0 && nd->right >= 0) {
            if (sp + 2 > 256) continue;   /* deeper than any tree the midpoint build produces */
            stack[sp++] = nd->right;
            stack[sp++] = nd->left;
        }
    }
    return hasHit;
}

/* ---- Mesh::Intersect, mesh.cpp:158-188 ---- */
static int mesh_intersect(const dt_scene_desc* sc, Ray* ray, int si) {
    const dt_shape* sh = &sc->shapes[si];
    const dt_mesh* m = &sc->meshes[sh->mesh];
    v3 oc = ray->origin, dc = ray->dir;
    ray->origin = apply_transform(sh->inverse_transform, ray->origin, 1.0f);
    ray->dir = apply_transform(sh->inverse_transform, ray->dir, 0.0f);
    if (sh->has_motion_blur) ray->origin = vadd(ray->origin, vscale(F3(sh->motion_blur), ray->mbTime));
    if (box_intersect(m->bbox_min, m->bbox_max, ray)) {
        int hasHit = intersect_bvh(sc, ray, si);
        ray->origin = oc; ray->dir = dc;
        if (hasHit) {
            ray->hit.hitPoint = vadd(ray->origin, vscale(ray->dir, ray->hit.minT));
            ray->hit.normal = vunit(apply_transform(sh->inverse_transpose_transform, ray->hit.normal, 0.0f));
        }
        return hasHit;
    }
    ray->origin = oc; ray->dir = dc;
    return 0;
}

/* ---- InstancedMesh::Intersect, instancedMesh.cpp:16-66 ---- */
static int instance_intersect(const dt_scene_desc* sc, Ray* ray, int si) {
    const dt_shape* sh = &sc->shapes[si];
    int hasHit = 0;
    v3 oc = ray->origin, dc = ray->dir;
    if (sh->has_motion_blur) ray->origin = vadd(ray->origin, vscale(F3(sh->motion_blur), ray->mbTime));
    if (box_intersect(sh->bbox_min, sh->bbox_max, ray)) {
        ray->origin = oc;
        ray->origin = apply_transform(sh->inverse_transform, ray->origin, 1.0f);
        ray->dir = apply_transform(sh->inverse_transform, ray->dir, 0.0f);
        if (sh->has_motion_blur) ray->origin = vadd(ray->origin, vscale(F3(sh->motion_blur), ray->mbTime));
        hasHit = intersect_bvh(sc, ray, sh->base_shape);
        if (hasHit) {
            ray->hit.hitPoint = vadd(oc, vscale(dc, ray->hit.minT));
            ray->hit.matId = sh->material;
            ray->hit.shape = si;
            ray->hit.normal = vunit(apply_transform(sh->inverse_transpose_transform, ray->hit.normal, 0.0f));
        }
        ray->origin = oc; ray->dir = dc;
    } else {
        /* the reference leaves the motion-blur shifted origin in place here (instancedMesh.cpp:22-29,62-65 restore only inside the
           if): a box miss of a moving instance leaks the shift into every later shape test of the scan and into the shading of the
           ray.  The product path does not reproduce that; the oracle does on request (DTO_FLAG_ORIGIN_LEAK), which is how the tests
           show that the leak is the ONLY difference. */
        if (!g_dto_origin_leak) ray->origin = oc;
    }
    return hasHit;
}

/* ---- Sphere::Intersect, sphere.cpp:13-193 ---- */
static inline float greyscale_sum(v3 c) { return c.x + c.y + c.z; }           /* sphere.cpp:9-11 */
static int sphere_intersect(const dt_scene_desc* sc, Ray* r, int si) {
    const dt_shape* sh = &sc->shapes[si];
    v3 center = F3(sh->center);
    float radius = sh->radius;
    v3 oc0 = r->origin, dc0 = r->dir;
    r->origin = apply_transform(sh->inverse_transform, r->origin, 1.0f);
    r->dir = apply_transform(sh->inverse_transform, r->dir, 0.0f);
    if (sh->has_motion_blur) r->origin = vadd(r->origin, vscale(F3(sh->motion_blur), r->mbTime));
    v3 oc = vsub(r->origin, center);
    float t;
    float c = vdot(oc, oc) - (radius * radius);
    float b = 2 * vdot(r->dir, oc);
    float a = vdot(r->dir, r->dir);
    float delta = b * b - (4 * a * c);
    if (delta < 0.0f) { r->origin = oc0; r->dir = dc0; return 0; }
    delta = sqrtf(delta);
    a = (float)(2.0 * a);
    float t1 = (-b + delta) / a;
    float t2 = (-b - delta) / a;
    t = t1 < t2 ? t1 : t2;
    if (t1 < t2) { if (t1 > 0.0f) t = t1; else t = t2; }
    else if (t2 < t1) { if (t2 > 0.0f) t = t2; else t = t1; }
    v3 localhit = vadd(r->origin, vscale(r->dir, t));
    r->origin = oc0; r->dir = dc0;
    if (!(t < r->hit.minT && t > 0.0f)) return 0;
    r->hit.minT = t;
    r->hit.matId = sh->material;
    r->hit.shape = si;
    r->hit.face = -1;
    r->hit.hasHit = 1;
    r->hit.hitPoint = vadd(r->origin, vscale(r->dir, t));
    v3 p = vsub(localhit, center);
    float phi = atan2f(p.z, p.x);
    float theta = acosf(p.y / radius);
    float u = (float)((-phi + M_PI) / (2.0f * M_PI));
    float v = (float)(theta / M_PI);
    r->hit.u = u; r->hit.v = v;
    if (sh->tex_normal >= 0) {
        /* sphere.cpp:95-115: the normal-map branch only reads the texture; the normal keeps its previous value */
    } else if (sh->tex_bump >= 0) {
        const dt_texture* bm = &sc->textures[sh->tex_bump];
        v3 tan, bitan;                                                          /* sphere.cpp:181-193 */
        tan.x = (float)(2 * M_PI * p.z); tan.y = 0; tan.z = (float)(-2 * M_PI * p.x);
        bitan.x = (float)(M_PI * p.y * cosf(phi)); bitan.y = (float)(-radius * M_PI * sinf(theta)); bitan.z = (float)(M_PI * p.y * sinf(phi));
        tan = vunit(tan); bitan = vunit(bitan);
        v3 N = vunit(vcross(bitan, tan));
        if (bm->kind == DT_TEX_PERLIN) {
            v3 g;
            float eps = 0.001;
            float hxyz = tex_world_sample(bm, p.x, p.y, p.z);
            g.x = (tex_world_sample(bm, p.x + eps, p.y, p.z) - hxyz) / eps;
            g.y = (tex_world_sample(bm, p.x, p.y + eps, p.z) - hxyz) / eps;
            g.z = (tex_world_sample(bm, p.x, p.y, p.z + eps) - hxyz) / eps;
            v3 gpar = vscale(N, vdot(g, N));
            v3 sg = vsub(g, gpar);
            r->hit.normal = vunit(vsub(N, sg));
        } else {
            float width = tex_width(sc, bm), height = tex_height(sc, bm);
            int i = (int)(u * width);
            int j = (int)(v * height);
            float normalizer = bm->normalizer, bumpFactor = bm->sample_multiplier;
            float h1 = greyscale_sum(vdiv(tex_direct_sample(sc, bm, i + 1, j), normalizer)) * bumpFactor;
            float h_uv = greyscale_sum(vdiv(tex_direct_sample(sc, bm, i, j), normalizer)) * bumpFactor;
            float h2 = greyscale_sum(vdiv(tex_direct_sample(sc, bm, i, j + 1), normalizer)) * bumpFactor;
            v3 q_u = vadd(tan, vscale(N, (h1 - h_uv)));
            v3 q_v = vadd(bitan, vscale(N, (h2 - h_uv)));
            r->hit.normal = vunit(vcross(q_v, q_u));
        }
    } else {
        r->hit.normal = vunit(vsub(localhit, center));
    }
    r->hit.normal = vunit(apply_transform(sh->inverse_transpose_transform, r->hit.normal, 0.0f));
    return 1;
}

static int shape_intersect(const dt_scene_desc* sc, Ray* ray, int si) {
    switch (sc->shapes[si].kind) {
        case DT_SHAPE_MESH: return mesh_intersect(sc, ray, si);
        case DT_SHAPE_INSTANCE: return instance_intersect(sc, ray, si);
        default: return sphere_intersect(sc, ray, si);
    }
}

/* ---- Raytracer::IntersectObjects, raytracer.cpp:625-643 ---- */
static void intersect_objects(Ctx* c, Ray* ray) {
    c->n_closest++;
    for (int i = 0; i < c->sc->n_shapes; i++) shape_intersect(c->sc, ray, i);
}

/* ---- Raytracer::CastShadowRay / IsInShadow / IsInShadowDirectional, raytracer.cpp:555-623 ---- */
static int cast_shadow_ray(Ctx* c, Ray* sr, float lightT) {
    const dt_scene_desc* sc = c->sc;
    c->n_shadow++;
    for (int i = 0; i < sc->n_mesh_shapes; i++) {
        if (sc->materials[sc->shapes[i].material - 1].type == DT_MAT_EMISSIVE) continue;
        shape_intersect(sc, sr, i);
        if (sr->hit.hasHit && sr->hit.minT < lightT) return 1;
    }
    for (int i = sc->n_mesh_shapes; i < sc->n_shapes; i++) {
        shape_intersect(sc, sr, i);
        if (sr->hit.hasHit && sr->hit.minT < lightT) return 1;
    }
    return 0;
}
static int is_in_shadow(Ctx* c, const Ray* orig, v3 lightPos) {
    Ray sr; memset(&sr, 0, sizeof sr);
    sr.dir = vsub(lightPos, orig->hit.hitPoint);
    float lightT = vlen(sr.dir);
    sr.dir = vdiv(sr.dir, lightT);
    sr.origin = vadd(orig->hit.hitPoint, vscale(orig->hit.normal, c->sc->shadow_ray_epsilon));
    sr.hit.hasHit = 0;
    sr.hit.minT = lightT + 0.01f;
    sr.mbTime = orig->mbTime;
    return cast_shadow_ray(c, &sr, lightT);
}
static int is_in_shadow_directional(Ctx* c, const Ray* orig, v3 lightDir) {
    Ray sr; memset(&sr, 0, sizeof sr);
    sr.dir = vneg(lightDir);
    sr.origin = vadd(orig->hit.hitPoint, vscale(orig->hit.normal, c->sc->shadow_ray_epsilon));
    sr.hit.hasHit = 0;
    sr.hit.minT = INFINITY;
    sr.mbTime = orig->mbTime;
    return cast_shadow_ray(c, &sr, INFINITY);
}

/* ---- BRDFs ---- */
static v3 brdf_apply(const dt_scene_desc* sc, const dt_material* mat, v3 kd, v3 ks, v3 w_i, v3 w_o, v3 n) {
    const dt_brdf* b = &sc->brdfs[mat->brdf];
    float exponent = b->exponent;
    float angleTheta_i = (float)angle_between_unit(w_i, n);
    switch (b->kind) {
        case DT_BRDF_PHONG: {                                                  /* brdfPhong.cpp:11-20 */
            if (angleTheta_i >= 90.0f || angleTheta_i < 0) return V(0, 0, 0);
            v3 pr = vunit(vsub(vscale(vscale(n, 2.0f), vdot(n, w_i)), w_i));
            double angleR = angle_between_unit(pr, w_o);
            return vadd(kd, vscale(ks, (float)(pow(cos_deg(angleR), exponent) / cos_deg(angleTheta_i))));
        }
        case DT_BRDF_BLINN_PHONG: {                                            /* brdfBlinnPhong.cpp:11-20 */
            if (angleTheta_i >= 90.0f) return V(0, 0, 0);
            v3 s = vadd(w_i, w_o);
            v3 half = vdiv(s, vlen(s));
            double a = angle_between_unit(half, n);
            return vadd(kd, vscale(ks, (float)(pow(cos_deg(a), exponent) / cos_deg(angleTheta_i))));
        }
        case DT_BRDF_MODIFIED_PHONG: {                                         /* brdfModifiedPhong.cpp:14-33 */
            if (angleTheta_i >= 90.0f || angleTheta_i < 0) return V(0, 0, 0);
            v3 pr = vunit(vsub(vscale(vscale(n, 2.0f), vdot(n, w_i)), w_i));
            double angleR = angle_between_unit(pr, w_o);
            if (b->flag) {
                v3 kdTerm = vscale(kd, (float)(1.0f / M_PI));
                double cons = (exponent + 2) / (2 * M_PI);
                double cosTerm = pow(cos_deg(angleR), exponent);
                v3 ksTerm = vscale(ks, (float)(cons * cosTerm));
                return vadd(kdTerm, ksTerm);
            }
            return vadd(kd, vscale(ks, (float)pow(cos_deg(angleR), exponent)));
        }
        case DT_BRDF_MODIFIED_BLINN_PHONG: {                                   /* brdfModifiedBlinnPhong.cpp:11-29 */
            if (angleTheta_i >= 90.0f) return V(0, 0, 0);
            v3 s = vadd(w_i, w_o);
            v3 half = vdiv(s, vlen(s));
            double a = angle_between_unit(half, n);
            if (b->flag) {
                v3 kdTerm = vscale(kd, (float)(1.0f / M_PI));
                double cons = (exponent + 8) / (8 * M_PI);
                double cosTerm = pow(cos_deg(a), exponent);
                v3 ksTerm = vscale(ks, (float)(cons * cosTerm));
                return vadd(kdTerm, ksTerm);
            }
            return vadd(kd, vscale(ks, (float)pow(cos_deg(a), exponent)));
        }
        default: {                                                             /* brdfTorranceSparrow.cpp:15-59 */
            if (angleTheta_i >= 90.0f) return V(0, 0, 0);
            v3 s = vadd(w_i, w_o);
            v3 half = vdiv(s, vlen(s));
            double e = exponent;
            double d = (e + 2) * pow((double)vdot(half, n), e) / (2 * M_PI);
            double ri = mat->refractive_index;
            double r0 = pow(ri - 1, 2) / pow(ri + 1, 2);
            double f = r0 + (1.0 - r0) * pow((1.0 - (double)vdot(half, w_o)), 5.0);
            double ndoth = vdot(n, half), ndotwo = vdot(n, w_o), ndotwi = vdot(n, w_i), wodoth = vdot(w_o, half);
            double g = fmin(1.0, fmin(2.0f * ndoth * ndotwo / wodoth, 2.0 * ndoth * ndotwi / wodoth));
            double kdCoeff = (1.0f / M_PI);
            if (b->flag) kdCoeff *= (1 - f);
            v3 kdTerm = vscale(kd, (float)kdCoeff);
            double costheta = vdot(n, w_i);
            double cosphi = vdot(n, w_o);
            v3 ksTerm = vscale(ks, (float)((d * f * g) / (4 * costheta * cosphi)));
            return vadd(kdTerm, ksTerm);
        }
    }
}

/* ---- Raytracer::Get{Diffuse,Specular}ReflectanceCoeff, raytracer.cpp:478-539 ---- */
static v3 reflectance_coeff(const dt_scene_desc* sc, const Ray* ray, const dt_material* mat, int specular) {
    const dt_shape* sh = &sc->shapes[ray->hit.shape];
    v3 reflectance = specular ? F3(mat->specular) : F3(mat->diffuse);
    int has = specular ? (sh->tex_specular >= 0) : (sh->tex_diffuse >= 0);
    if (!has) return reflectance;
    /* both paths read shape->diffuseTex (the specular path too, raytracer.cpp:516-531).  A shape with a specular
       but no diffuse texture dereferences nullptr in the reference; we leave the coefficient untouched then. */
    if (sh->tex_diffuse < 0) return reflectance;
    const dt_texture* t = &sc->textures[sh->tex_diffuse];
    v3 tk;
    if (t->kind == DT_TEX_PERLIN) {
        v3 hp = ray->hit.hitPoint;
        float s = tex_world_sample(t, hp.x, hp.y, hp.z);
        tk = V(s, s, s);
    } else {
        tk = vdiv(tex_rgb_sample(sc, t, ray->hit.u, ray->hit.v), 255.0f);
    }
    if (t->decal_mode == DT_DECAL_BLEND_KD) reflectance = vdiv(vadd(tk, F3(mat->diffuse)), 2.0f);
    else reflectance = tk;
    return reflectance;
}

/* ---- Raytracer::Shade / GetDiffuse / GetSpecular, raytracer.cpp:192-206, 540-554 ---- */
static v3 shade(const dt_scene_desc* sc, Ray* ray, const dt_material* mat, v3 w_i, v3 w_o, v3 Li) {
    if (mat->brdf >= 0) {
        float costheta_i = fmaxf(0.0f, vdot(w_i, ray->hit.normal));
        v3 kd = reflectance_coeff(sc, ray, mat, 0);
        v3 ks = reflectance_coeff(sc, ray, mat, 1);
        v3 res = brdf_apply(sc, mat, kd, ks, w_i, w_o, ray->hit.normal);
        ray->throughput = vmul(ray->throughput, res);
        return vscale(vmul(res, Li), costheta_i);
    }
    v3 kd = reflectance_coeff(sc, ray, mat, 0);
    float costheta = fmaxf(0.0f, vdot(w_i, ray->hit.normal));
    v3 diffuse = vscale(vmul(kd, Li), costheta);
    v3 ks = reflectance_coeff(sc, ray, mat, 1);
    v3 s = vadd(w_i, w_o);
    v3 half = vdiv(s, vlen(s));
    float cosAlpha = fmaxf(0.0f, vdot(ray->hit.normal, half));
    v3 spec = vscale(vmul(ks, Li), powf(cosAlpha, mat->phong_exponent));
    return vadd(diffuse, spec);
}

/* ---- environment light, sphericalEnvironmentLight.h:22-65 ---- */
static v3 env_sample(const dt_scene_desc* sc, v3 dir) {
    const dt_image* im = &sc->images[sc->env_lights[0].image];
    float u = (float)((1 + (atan2f(dir.x, -dir.z) / M_PI)) / 2.0f);
    float v = (float)(acosf(dir.y) / M_PI);
    int i = (int)(im->width * u);
    int j = (int)(im->height * v);
    v3 s = image_sample(im, i, j);
    return vscale(vscale(s, 2), (float)M_PI);   /* Vec3f * 2 * M_PI: both factors narrow to float (helperMath.cpp:22) */
}
static v3 env_sample_idx(const dt_scene_desc* sc, int li, v3 dir) {
    const dt_image* im = &sc->images[sc->env_lights[li].image];
    float u = (float)((1 + (atan2f(dir.x, -dir.z) / M_PI)) / 2.0f);
    float v = (float)(acosf(dir.y) / M_PI);
    int i = (int)(im->width * u);
    int j = (int)(im->height * v);
    v3 s = image_sample(im, i, j);
    return vscale(vscale(s, 2), (float)M_PI);   /* Vec3f * 2 * M_PI: both factors narrow to float (helperMath.cpp:22) */
}
static v3 env_direction(Ctx* c, int li, v3 normal) {
    v3 n = vunit(normal);
    v3 cand;
    for (int guard = 0;; guard++) {
        cand.x = (float)(-1.0f + 2.0 * rnd_s(c, RS_ENV, li));
        cand.y = (float)(-1.0f + 2.0 * rnd_s(c, RS_ENV, li));
        cand.z = (float)(-1.0f + 2.0 * rnd_s(c, RS_ENV, li));
        float length = vlen(cand);
        if (length <= 1.0f && vdot(n, cand) > 0.0f) break;    /* `candidate / length;` is a no-op in the reference */
        if (guard > 100000) break;
    }
    return cand;
}

/* ---- spotLight.h:33-57 ---- */
static v3 spot_irradiance(const dt_spot_light* l, v3 point) {
    v3 pos = F3(l->pos);
    float dist = vlen(vsub(point, pos));
    v3 toPoint = vdiv(vsub(point, pos), dist);
    double alpha = angle_between_unit(F3(l->dir), toPoint);
    if (alpha <= 0 || alpha > (l->coverage_angle / 2.0f)) return V(0, 0, 0);
    float distSqr = dist * dist;
    v3 irr = vdiv(F3(l->intensity), distSqr);
    if (alpha > (l->falloff_angle / 2.0f)) {
        double cosAlpha = cos(alpha * DEG2RAD);
        double s = pow((cosAlpha - l->cos_half_coverage) / (l->cos_half_falloff - l->cos_half_coverage), 4.0f);
        irr = vscale(irr, (float)s);
    }
    return irr;
}

/* ---- Raytracer::SampleDirectLighting, raytracer.cpp:701-806 ---- */
static v3 sample_direct_lighting(Ctx* c, Ray* ray, const dt_material* mat, v3 w_o, int lightIdToSkip) {
    const dt_scene_desc* sc = c->sc;
    v3 color = V(0, 0, 0);
    for (int i = 0; i < sc->n_point_lights; i++) {
        const dt_point_light* l = &sc->point_lights[i];
        v3 lp = F3(l->position);
        if (is_in_shadow(c, ray, lp)) continue;
        v3 w_i = vunit(vsub(lp, ray->hit.hitPoint));
        float dist = vlen(vsub(lp, ray->hit.hitPoint));
        v3 E = vdiv(F3(l->intensity), (dist * dist));
        color = vadd(color, shade(sc, ray, mat, w_i, w_o, E));
    }
    for (int i = 0; i < sc->n_area_lights; i++) {
        const dt_area_light* l = &sc->area_lights[i];
        float offU = (float)(-0.5f + rnd_s(c, RS_AREA, i));     /* areaLight.h:34-40 */
        float offV = (float)(-0.5f + rnd_s(c, RS_AREA, i));
        v3 sp = vadd(vadd(F3(l->position), vscale(F3(l->u), (l->extent * offU))), vscale(F3(l->v), (l->extent * offV)));
        if (is_in_shadow(c, ray, sp)) continue;
        v3 w_i = vsub(sp, ray->hit.hitPoint);
        float dist = vlen(w_i);
        float dSqr = dist * dist;
        w_i = vdiv(w_i, dist);
        float lCos = vdot(F3(l->normal), vneg(w_i));
        if (lCos < 0) lCos = vdot(F3(l->normal), w_i);
        float area = l->extent * l->extent;
        v3 E = vscale(F3(l->radiance), (area * lCos / dSqr));
        color = vadd(color, shade(sc, ray, mat, w_i, w_o, E));
    }
    for (int i = 0; i < sc->n_env_lights; i++) {
        v3 dir = env_direction(c, i, ray->hit.normal);
        v3 E = env_sample_idx(sc, i, dir);
        v3 w_i = ray->hit.normal;
        color = vadd(color, shade(sc, ray, mat, w_i, w_o, E));
    }
    for (int i = 0; i < sc->n_directional_lights; i++) {
        const dt_directional_light* l = &sc->directional_lights[i];
        if (is_in_shadow_directional(c, ray, F3(l->dir))) continue;
        v3 w_i = vneg(F3(l->dir));
        color = vadd(color, shade(sc, ray, mat, w_i, w_o, F3(l->radiance)));
    }
    for (int i = 0; i < sc->n_spot_lights; i++) {
        const dt_spot_light* l = &sc->spot_lights[i];
        if (is_in_shadow(c, ray, F3(l->pos))) continue;
        v3 w_i = vunit(vsub(F3(l->pos), ray->hit.hitPoint));
        v3 E = spot_irradiance(l, ray->hit.hitPoint);
        color = vadd(color, shade(sc, ray, mat, w_i, w_o, E));
    }
    for (int i = 0; i < sc->n_mesh_lights; i++) {
        const dt_mesh_light* l = &sc->mesh_lights[i];
        if (l->id == lightIdToSkip) continue;
        const dt_shape* lsh = &sc->shapes[l->shape];
        const dt_mesh* lm = &sc->meshes[lsh->mesh];
        /* meshLight.h:27-47 with patch P2 (uniform face pick over [0, faceCount-1]) */
        int fi;
        if (c->ref) fi = (int)mt_uniform_int(&c->ref->mesh[i < DTO_MAX_LIGHT_STREAMS ? i : DTO_MAX_LIGHT_STREAMS - 1], (uint32_t)lm->n_faces);
        else fi = (int)(rnd01(c) * lm->n_faces);
        if (fi >= lm->n_faces) fi = lm->n_faces - 1;
        const dt_face* face = &lm->faces[fi];
        double weight = face->area / lm->surface_area;
        double rand1 = rnd_s(c, RS_MESH, i), rand2 = rnd_s(c, RS_MESH, i);
        v3 a = mesh_vertex(lm, face->v0_id), b = mesh_vertex(lm, face->v1_id), cc = mesh_vertex(lm, face->v2_id);
        v3 q = vadd(vscale(b, (float)(1 - rand2)), vscale(cc, (float)rand2));
        v3 pos = vadd(vscale(a, (float)(1 - sqrt(rand1))), vscale(q, (float)sqrt(rand1)));
        pos = apply_transform(lsh->transform, pos, 1.0f);
        v3 lightNormal = F3(face->n);
        if (is_in_shadow(c, ray, pos)) continue;
        v3 w_i = vsub(pos, ray->hit.hitPoint);
        float dist = vlen(w_i);
        w_i = vdiv(w_i, dist);
        (void)lightNormal;
        v3 rad = vscale(vscale(vscale(F3(l->radiance), (float)weight), 2), (float)M_PI);
        color = vadd(color, shade(sc, ray, mat, w_i, w_o, rad));
    }
    return color;
}

static v3 perform_shading(Ctx* c, Ray* ray, v3 eyePos, int recDepth);

static Ray secondary_ray(const Ray* orig, v3 dir, v3 origin) {                 /* raytracer.cpp:645-660 */
    Ray r; memset(&r, 0, sizeof r);
    r.dir = dir; r.origin = origin;
    r.hit.hasHit = 0; r.hit.minT = INFINITY; r.hit.shape = -1; r.hit.face = -1;
    r.n_medium = orig->n_medium;
    r.mbTime = orig->mbTime;
    r.throughput = orig->throughput;
    return r;
}
static v3 reflect_dir(Ctx* c, v3 normal, v3 w_o, float roughness) {            /* raytracer.cpp:424-440 */
    v3 r = vunit(vsub(vscale(vscale(normal, 2.0f), vdot(normal, w_o)), w_o));
    if (roughness > 0.001) {
        v3 u, v;
        orthonormal_basis(r, &u, &v);
        float psi1 = rnd_normalized(c) - 0.5f;
        float psi2 = rnd_normalized(c) - 0.5f;
        return vunit(vadd(r, vscale(vadd(vscale(u, psi1), vscale(v, psi2)), roughness)));
    }
    return r;
}
static v3 beers_law(float x, v3 cf, v3 L0) {                                   /* raytracer.cpp:416-423 */
    return V(L0.x * expf(-cf.x * x), L0.y * expf(-cf.y * x), L0.z * expf(-cf.z * x));
}

/* ---- ComputeMirrorReflection, raytracer.cpp:442-472 ---- */
static v3 mirror_reflection(Ctx* c, Ray* orig, const dt_material* mat, v3 w_o, int recDepth) {
    if (recDepth <= 0) return V(0, 0, 0);
    v3 w_r = reflect_dir(c, orig->hit.normal, w_o, mat->roughness);
    v3 origin = vadd(orig->hit.hitPoint, vscale(orig->hit.normal, c->sc->shadow_ray_epsilon));
    Ray rr = secondary_ray(orig, w_r, origin);
    rr.n_medium = 1.0f;
    intersect_objects(c, &rr);
    if (rr.hit.hasHit) return vmul(F3(mat->mirror), perform_shading(c, &rr, rr.origin, recDepth - 1));
    if (c->sc->n_env_lights > 0) return vmul(F3(mat->mirror), env_sample(c->sc, rr.dir));
    return V(0, 0, 0);
}

/* ---- ComputeConductorFresnelReflection, raytracer.cpp:208-254 ---- */
static v3 conductor_reflection(Ctx* c, Ray* orig, const dt_material* mat, v3 w_o, int recDepth) {
    if (recDepth <= 0) return V(0, 0, 0);
    v3 d = vneg(w_o);
    float cosTheta = -vdot(d, orig->hit.normal);
    float n2 = mat->refractive_index, k2 = mat->conductor_absorption_index;
    float n2k2 = n2 * n2 + k2 * k2;
    float n2cosTheta2 = 2 * n2 * cosTheta;
    float cosThetaSqr = cosTheta * cosTheta;
    float rs = (n2k2 - n2cosTheta2 + cosThetaSqr) / (n2k2 + n2cosTheta2 + cosThetaSqr);
    float rp = (n2k2 * cosThetaSqr - n2cosTheta2 + 1) / (n2k2 * cosThetaSqr + n2cosTheta2 + 1);
    float reflectRatio = (float)(0.5 * (rs + rp));
    if (reflectRatio > 0.0001) {
        v3 col;
        v3 w_r = reflect_dir(c, orig->hit.normal, w_o, mat->roughness);
        v3 origin = vadd(orig->hit.hitPoint, vscale(orig->hit.normal, c->sc->shadow_ray_epsilon));
        Ray rr = secondary_ray(orig, w_r, origin);
        rr.n_medium = 1.0f;
        intersect_objects(c, &rr);
        if (rr.hit.hasHit) col = vmul(F3(mat->mirror), perform_shading(c, &rr, rr.origin, recDepth - 1));
        else col = V(0, 0, 0);
        return vscale(col, reflectRatio);
    }
    return V(0, 0, 0);
}

/* ---- ComputeDielectricFresnelReflectionAndRefraction, raytracer.cpp:261-415 ---- */
static v3 dielectric(Ctx* c, Ray* orig, const dt_material* mat, v3 w_o, float n1, float n2, int recDepth) {
    const dt_scene_desc* sc = c->sc;
    if (recDepth <= 0) return V(0, 0, 0);
    v3 d = vneg(w_o);
    v3 mn = orig->hit.normal;
    float cosTheta = -vdot(d, mn);
    int isEntering = cosTheta > 0.f;
    float objN = n2;
    if (!isEntering) {
        n1 = n2;
        n2 = 1.0f;
        objN = 1.0f;
        cosTheta = fabsf(cosTheta);
        mn = vneg(mn);
    }
    float r = n1 / n2;
    float sinThetaSqr = 1 - (cosTheta * cosTheta);
    float criticalTerm = r * r * sinThetaSqr;
    v3 absorb = F3(mat->absorption_coefficient);
    if (criticalTerm > 1) {
        v3 w_r = reflect_dir(c, mn, w_o, mat->roughness);
        v3 no = vadd(orig->hit.hitPoint, vscale(mn, sc->shadow_ray_epsilon));
        Ray rr = secondary_ray(orig, w_r, no);
        intersect_objects(c, &rr);
        v3 col = V(0, 0, 0);
        if (rr.hit.hasHit) {
            col = perform_shading(c, &rr, rr.origin, recDepth - 1);
            if (rr.n_medium > 1.0001) col = beers_law(rr.hit.minT, absorb, col);
        }
        return col;
    }
    float cosPhi = sqrtf(1 - criticalTerm);
    float n2cosTheta = n2 * cosTheta;
    float n1cosPhi = n1 * cosPhi;
    float rparallel = (n2cosTheta - n1cosPhi) / (n2cosTheta + n1cosPhi);
    float rperp = (n1 * cosTheta - n2 * cosPhi) / (n1 * cosTheta + n2 * cosPhi);
    float rReflect = (rparallel * rparallel + rperp * rperp) / 2;
    float rRefract = 1 - rReflect;

    v3 w_reflected = reflect_dir(c, mn, w_o, mat->roughness);
    v3 no = vadd(orig->hit.hitPoint, vscale(mn, sc->shadow_ray_epsilon));
    Ray rr = secondary_ray(orig, w_reflected, no);
    intersect_objects(c, &rr);
    rr.n_medium = isEntering ? objN : 1.0f;
    v3 reflCol = V(0, 0, 0);
    if (rr.hit.hasHit) {
        reflCol = perform_shading(c, &rr, rr.origin, recDepth - 1);
        if (rr.n_medium > 1.00001f) reflCol = beers_law(rr.hit.minT, absorb, reflCol);
    } else if (sc->n_env_lights > 0) {
        reflCol = env_sample(sc, rr.dir);
    }
    v3 refrCol;
    {
        v3 w_t = vsub(vscale(vadd(d, vscale(mn, cosTheta)), r), vscale(mn, cosPhi));
        if (mat->roughness > 0.001) {
            v3 u, v;
            orthonormal_basis(w_t, &u, &v);
            float psi1 = rnd_normalized(c) - 0.5f;
            float psi2 = rnd_normalized(c) - 0.5f;
            w_t = vunit(vadd(w_t, vscale(vadd(vscale(u, psi1), vscale(v, psi2)), mat->roughness)));
        } else w_t = vunit(w_t);
        v3 no2 = vadd(orig->hit.hitPoint, vscale(vneg(mn), sc->shadow_ray_epsilon));
        Ray tr = secondary_ray(orig, w_t, no2);
        tr.n_medium = isEntering ? objN : 1.0f;
        intersect_objects(c, &tr);
        refrCol = V(0, 0, 0);
        if (tr.hit.hasHit) {
            refrCol = perform_shading(c, &tr, tr.origin, recDepth - 1);
            if (tr.n_medium > 1.001f) refrCol = beers_law(tr.hit.minT, absorb, refrCol);
        } else if (sc->n_env_lights > 0) {
            refrCol = env_sample(sc, rr.dir);          /* uses reflectedRay.dir (raytracer.cpp:408) */
        }
    }
    return vadd(vscale(reflCol, rReflect), vscale(refrCol, rRefract));
}

/* ---- ComputeGlobalIllumination, raytracer.cpp:135-191 ---- */
static v3 global_illumination(Ctx* c, Ray* ray, const dt_material* orgMat, v3 w_o, int recDepth, int* hitMeshLightId) {
    const dt_camera_desc* cam = c->cam;
    if (cam->russian_roulette) {
        float probTest = rnd_normalized(c);
        float maxT = fmaxf(ray->throughput.x, fmaxf(ray->throughput.x, ray->throughput.z));
        if (probTest > maxT && recDepth <= 0) return V(0, 0, 0);
        ray->throughput = vdiv(ray->throughput, maxT);
    } else if (recDepth <= 0) return V(0, 0, 0);
    float rand1 = rnd_normalized(c);
    float rand2 = rnd_normalized(c);
    float phi = (float)(2 * M_PI * rand1);
    float theta;
    if (cam->importance_sampling) theta = asinf(sqrtf(rand2));
    else theta = acosf(rand2);
    v3 u, v;
    orthonormal_basis(ray->hit.normal, &u, &v);
    v3 nd = vadd(vadd(vscale(vscale(u, sinf(theta)), cosf(phi)), vscale(ray->hit.normal, cosf(theta))), vscale(vscale(v, sinf(theta)), sinf(phi)));
    nd = vunit(nd);
    v3 no = vadd(ray->hit.hitPoint, vscale(ray->hit.normal, (float)0.0001));
    Ray gr = secondary_ray(ray, nd, no);
    intersect_objects(c, &gr);
    v3 col = V(0, 0, 0);
    if (gr.hit.hasHit) {
        const dt_material* m = &c->sc->materials[gr.hit.matId - 1];
        if (m->type == DT_MAT_EMISSIVE) *hitMeshLightId = c->sc->shapes[gr.hit.shape].id;
        v3 Li = perform_shading(c, &gr, gr.origin, recDepth - 1);
        v3 s = shade(c->sc, ray, orgMat, gr.dir, w_o, Li);
        col = vscale(vscale(s, 2.0f), (float)M_PI);
    }
    return col;
}

/* ---- PerformShading, raytracer.cpp:65-134 ---- */
static v3 perform_shading(Ctx* c, Ray* ray, v3 eyePos, int recDepth) {
    const dt_scene_desc* sc = c->sc;
    const dt_camera_desc* cam = c->cam;
    ray->hit.hitPoint = vadd(ray->origin, vscale(ray->dir, ray->hit.minT));
    v3 color = V(0, 0, 0);
    const dt_material* mat = &sc->materials[ray->hit.matId - 1];
    const dt_shape* shape = &sc->shapes[ray->hit.shape];
    v3 w_o = vunit(vsub(eyePos, ray->hit.hitPoint));
    float vac = 1.00001;
    int inside = ray->n_medium > vac;
    if (mat->type == DT_MAT_EMISSIVE) return vscale(vscale(F3(mat->radiance), 2.0f), (float)M_PI);
    if (shape->tex_replace_all >= 0) return tex_rgb_sample(sc, &sc->textures[shape->tex_replace_all], ray->hit.u, ray->hit.v);
    int hitLightMeshId = -1;
    if (cam->path_tracing) color = vadd(color, global_illumination(c, ray, mat, w_o, recDepth, &hitLightMeshId));
    int sampleDirect = !cam->path_tracing || (cam->path_tracing && cam->next_event_estimation);
    if (!inside && sampleDirect) {
        color = vadd(color, vmul(F3(sc->ambient_light), F3(mat->ambient)));
        color = vadd(color, sample_direct_lighting(c, ray, mat, w_o, hitLightMeshId));
    }
    if (mat->type == DT_MAT_MIRROR) color = vadd(color, mirror_reflection(c, ray, mat, w_o, recDepth));
    else if (mat->type == DT_MAT_DIELECTRIC) color = vadd(color, dielectric(c, ray, mat, w_o, ray->n_medium, mat->refractive_index, recDepth));
    else if (mat->type == DT_MAT_CONDUCTOR) color = vadd(color, conductor_reflection(c, ray, mat, w_o, recDepth));
    return color;
}

/* ---- Camera::GetImagePlanePosition (camera.cpp:74-80) + Raytracer::GenerateRay (raytracer.cpp:661-699) ---- */
static Ray generate_ray(Ctx* c, int i, int j) {
    const dt_camera_desc* cam = c->cam;
    Ray ray; memset(&ray, 0, sizeof ray);
    float su = (float)((i + 0.5) * (cam->right_ - cam->left) / cam->width);
    float sv = (float)((j + 0.5) * (cam->top - cam->bottom) / cam->height);
    v3 ipp = vadd(vadd(F3(cam->q), vscale(F3(cam->right), su)), vscale(F3(cam->up), -sv));
    ray.origin = F3(cam->position);
    ray.throughput = V(1.0f, 1.0f, 1.0f);
    if (cam->aperture_size > 0.0001) {
        v3 ap = ray.origin;
        float first01 = (float)(-1.0f + 2.0 * rnd_s(c, RS_DOF, 0));
        ap = vadd(ap, vscale(F3(cam->up), (first01 * cam->aperture_size * 0.5f)));
        float second01 = (float)(-1.0f + 2.0 * rnd_s(c, RS_DOF, 0));
        ap = vadd(ap, vscale(F3(cam->right), (second01 * cam->aperture_size * 0.5f)));
        v3 dir = vunit(vsub(ray.origin, ipp));
        float tFd = cam->focus_distance / vdot(dir, F3(cam->gaze));
        v3 bent = vadd(ray.origin, vscale(dir, tFd));
        ray.dir = vunit(vsub(bent, ap));
        ray.origin = ap;
    } else {
        ray.dir = vunit(vsub(ipp, ray.origin));
    }
    ray.hit.hasHit = 0; ray.hit.minT = INFINITY; ray.hit.shape = -1; ray.hit.face = -1;
    ray.n_medium = 1.0f;
    ray.mbTime = rnd_normalized(c);
    return ray;
}

/* ---- Raytracer::PerPixel, raytracer.cpp:38-63 ---- */
static v3 per_pixel(Ctx* c, int x, int y) {
    const dt_scene_desc* sc = c->sc;
    Ray ray = generate_ray(c, x, y);
    intersect_objects(c, &ray);
    if (ray.hit.hasHit) return perform_shading(c, &ray, F3(c->cam->position), sc->max_recursion_depth);
    if (sc->bg_texture >= 0) {
        float u = x / (float)c->cam->width;
        float v = y / (float)c->cam->height;
        return tex_rgb_sample(sc, &sc->textures[sc->bg_texture], u, v);
    }
    if (sc->n_env_lights > 0) return env_sample(sc, ray.dir);
    return V((float)sc->background_color[0], (float)sc->background_color[1], (float)sc->background_color[2]);
}

/* helperMath.cpp:140-152: (int) is cvttss2si: NaN / out-of-range -> INT_MIN -> clamps to 0 */
static int clamp_channel(float f) {
    int v;
    if (!(f > -2147483904.0f && f < 2147483648.0f)) v = INT32_MIN; else v = (int)f;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

/* ---- renderThreadMain, main.cpp:26-130 ---- */
typedef struct {
    const dt_scene_desc* sc; const dt_camera_desc* cam; uint64_t seed;
    int y0, y1; uint8_t* ldr; float* hdr; uint64_t n_closest, n_shadow;
    RefRng* ref;
} Job;

static uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static void* render_rows(void* arg) {
    Job* job = (Job*)arg;
    const dt_camera_desc* cam = job->cam;
    Ctx c; c.sc = job->sc; c.cam = cam; c.n_closest = c.n_shadow = 0; c.rng = 0; c.ref = job->ref;
    int width = cam->width;
    int spp = cam->samples_per_pixel;
    int nRows = (int)sqrt((double)spp), nCols = nRows;
    float sigma = 1.0f / 6.0f;                                  /* gaussian.h:3-21 */
    float sigmaSqr = sigma * sigma;
    float c1 = (float)(1.0f / (2.0f * M_PI * sigmaSqr));
    float* sx = (float*)calloc((size_t)(spp > 0 ? spp : 1), sizeof(float));
    float* sy = (float*)calloc((size_t)(spp > 0 ? spp : 1), sizeof(float));
    for (int y = job->y0; y < job->y1; y++) {
        for (int x = 0; x < width; x++) {
            if (!c.ref) c.rng = mix64(job->seed ^ mix64((uint64_t)(x + (uint64_t)y * (uint64_t)width) + 0x51ED270B7F4A7C15ull));
            v3 color = V(0, 0, 0);
            if (spp > 1) {
                int i = 0;
                for (int row = 0; row < nRows; row++) for (int col = 0; col < nCols; col++) {
                    float psi1 = (float)rnd_s(&c, RS_MAIN, 0);
                    float psi2 = (float)rnd_s(&c, RS_MAIN, 0);
                    sx[i] = (col + psi1) / nCols;
                    sy[i] = (row + psi2) / nRows;
                    i++;
                }
                /* the reference iterates samplesPerPixel entries (main.cpp:81); with a non-square count the entries past nRows^2 are
                   never written: they keep the zeros `std::vector<Vec2f> samples(samplesPerPixel)` was created with (main.cpp:47), i.e.
                   extra rays through the pixel at sample position (0,0), Gaussian weight of the pixel corner */
                int ns = spp;
                float sumW = 0.0f;
                for (i = 0; i < ns; i++) {
                    /* RenderPixel(int,int): the float sample position is truncated (main.cpp:83, raytracer.hpp:19) */
                    v3 col = per_pixel(&c, (int)(sx[i] + x), (int)(sy[i] + y));
                    float xd = sx[i] - 0.5f, yd = sy[i] - 0.5f;
                    float exponent = (float)(-0.5 * ((xd * xd + yd * yd) / sigmaSqr));
                    float w = c1 * expf(exponent);
                    color.x += col.x * w; color.y += col.y * w; color.z += col.z * w;
                    sumW += w;
                }
                color.x = color.x / sumW; color.y = color.y / sumW; color.z = color.z / sumW;
            } else {
                color = per_pixel(&c, x, y);
            }
            size_t idx = 3 * ((size_t)x + (size_t)y * width);
            if (job->hdr) { job->hdr[idx] = color.x; job->hdr[idx + 1] = color.y; job->hdr[idx + 2] = color.z; }
            if (job->ldr && !cam->has_tonemapper) {
                job->ldr[idx] = (uint8_t)clamp_channel(color.x);
                job->ldr[idx + 1] = (uint8_t)clamp_channel(color.y);
                job->ldr[idx + 2] = (uint8_t)clamp_channel(color.z);
            }
        }
    }
    free(sx); free(sy);
    job->n_closest = c.n_closest; job->n_shadow = c.n_shadow;
    return NULL;
}

/* ---- Tonemapper::Tonemap, tonemapper.h:28-119 ---- */
static int cmp_float(const void* a, const void* b) { float x = *(const float*)a, y = *(const float*)b; return (x > y) - (x < y); }

int dto_tonemap(const float* hdr, int32_t width, int32_t height, float key, float burn, float saturation, float gamma, uint8_t* ldr) {
    size_t n = (size_t)width * height;
    float* sorted = (float*)malloc(sizeof(float) * n * 3);
    if (!sorted) return DT_ERR_INVALID;
    double logSum = 0.0f;
    float delta = 0.01f;
    for (size_t i = 0; i < n; i++) {
        double r = sorted[3 * i] = hdr[3 * i], g = sorted[3 * i + 1] = hdr[3 * i + 1], b = sorted[3 * i + 2] = hdr[3 * i + 2];
        double lum = 0.2126 * r + 0.7152 * g + 0.0722 * b;
        logSum += log(delta + lum);
    }
    long pixelCount = (long)width * height;
    double avgLum = exp(logSum / (double)pixelCount);
    qsort(sorted, n * 3, sizeof(float), cmp_float);
    for (size_t i = 0; i < n; i++) {
        double R = hdr[3 * i], G = hdr[3 * i + 1], B = hdr[3 * i + 2];
        double y_i = 0.2126 * R + 0.7152 * G + 0.0722 * B;
        /* Reinhard(): returns float */
        float y_of;
        {
            double Lxy = (key * y_i) / avgLum;
            if (burn > 0.01) {
                float thresholdPerct = (100.0f - burn) / 100;
                int lastIdx = (int)(n * 3) - 1;
                int bi = (int)(thresholdPerct * lastIdx);
                if (bi > lastIdx) bi = lastIdx;
                double thr = sorted[bi];
                thr = thr * key / avgLum;
                double LwhiteSqr = thr * thr;
                double res = (Lxy * (1 + (Lxy / LwhiteSqr))) / (1.0f + Lxy);
                y_of = (float)res;
            } else y_of = (float)(Lxy / (1 + Lxy));
        }
        double y_o = y_of;
        double r_o = clipf((float)(y_o * pow((R / y_i), saturation)), 0.0f, 1.0f);
        double g_o = clipf((float)(y_o * pow((G / y_i), saturation)), 0.0f, 1.0f);
        double b_o = clipf((float)(y_o * pow((B / y_i), saturation)), 0.0f, 1.0f);
        double gammaInv = 1.0f / gamma;
        int cr = (int)floor(fmin(255.0, 255 * pow(r_o, gammaInv)));
        int cg = (int)floor(fmin(255.0, 255 * pow(g_o, gammaInv)));
        int cb = (int)floor(fmin(255.0, 255 * pow(b_o, gammaInv)));
        ldr[3 * i] = (uint8_t)cr; ldr[3 * i + 1] = (uint8_t)cg; ldr[3 * i + 2] = (uint8_t)cb;
    }
    free(sorted);
    return DT_OK;
}

/* Render one camera like main.cpp:142-196.  n_threads row bands (the reference hard-codes 8; rows H mod
 * n_threads at the bottom are rendered here too).  hdr may be NULL unless the camera has a tonemapper. */
/* Render flags of include/dorktracer.h that change the image and are not part of the reference (process-wide; tests only). */
#define DTO_FLAG_ORIGIN_LEAK (1 << 30)
void dto_set_render_flags(int flags) { g_dto_smooth = (flags & DT_FLAG_SMOOTH_SHADING) ? 1 : 0; g_dto_origin_leak = (flags & DTO_FLAG_ORIGIN_LEAK) ? 1 : 0; }

int dto_render(const dt_scene_desc* sc, const dt_camera_desc* cam, uint64_t seed, int n_threads,
               uint8_t* ldr, float* hdr, dt_stats* stats) {
    if (!sc || !cam || !ldr) return DT_ERR_INVALID;
    int H = cam->height, W = cam->width;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > H) n_threads = H;
    float* own_hdr = NULL;
    if (!hdr && cam->has_tonemapper) { own_hdr = (float*)malloc(sizeof(float) * (size_t)W * H * 3); hdr = own_hdr; }
    Job* jobs = (Job*)calloc((size_t)n_threads, sizeof(Job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    /* interleaved thin bands keep the threads balanced; the per-pixel RNG keying makes the result band-independent */
    int band = (H + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; t++) {
        jobs[t].sc = sc; jobs[t].cam = cam; jobs[t].seed = seed; jobs[t].ldr = ldr; jobs[t].hdr = hdr;
        jobs[t].y0 = t * band; jobs[t].y1 = (t + 1) * band > H ? H : (t + 1) * band;
        if (jobs[t].y0 > H) jobs[t].y0 = H;
        pthread_create(&th[t], NULL, render_rows, &jobs[t]);
    }
    uint64_t nc = 0, ns = 0;
    for (int t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); nc += jobs[t].n_closest; ns += jobs[t].n_shadow; }
    if (cam->has_tonemapper) dto_tonemap(hdr, W, H, cam->tm_key, cam->tm_burn, cam->tm_saturation, cam->tm_gamma, ldr);
    if (stats) { memset(stats, 0, sizeof *stats); stats->rays_closest = nc; stats->rays_shadow = ns; }
    free(jobs); free(th); free(own_hdr);
    return DT_OK;
}

/* Reference-RNG mode: the FIRST camera of a scene rendered by the reference with one render thread (DT_THREADS=1; patch P3 of
 * oracle/build_ref.py).  Order of the reference's rand() calls from program start: one per SphericalDirectionalLight while
 * parsing (sphericalEnvironmentLight.h:19), Raytracer::randGen, Raytracer::dofLensSampleGenerator (raytracer.cpp:10,14), then
 * the render thread's sample generator (main.cpp:49).  Pixels in row-major order, all generators shared by all pixels. */
int dto_render_reference_rng(const dt_scene_desc* sc, const dt_camera_desc* cam, uint8_t* ldr, float* hdr, dt_stats* stats) {
    if (!sc || !cam || !ldr) return DT_ERR_INVALID;
    if (sc->n_area_lights > DTO_MAX_LIGHT_STREAMS || sc->n_mesh_lights > DTO_MAX_LIGHT_STREAMS || sc->n_env_lights > DTO_MAX_LIGHT_STREAMS) return DT_ERR_UNSUPPORTED;
    RefRng* ref = (RefRng*)malloc(sizeof(RefRng));
    GlibcRand g;
    glibc_rand_init(&g, 1u);
    for (int i = 0; i < DTO_MAX_LIGHT_STREAMS; i++) { mt_seed(&ref->area[i], 5489u); mt_seed(&ref->mesh[i], 5489u); mt_seed(&ref->env[i], 5489u); }
    for (int i = 0; i < sc->n_env_lights; i++) mt_seed(&ref->env[i], glibc_rand(&g));
    mt_seed(&ref->rand_gen, glibc_rand(&g));
    mt_seed(&ref->dof_gen, glibc_rand(&g));
    mt_seed(&ref->main_gen, glibc_rand(&g));
    int H = cam->height, W = cam->width;
    float* own_hdr = NULL;
    if (!hdr && cam->has_tonemapper) { own_hdr = (float*)malloc(sizeof(float) * (size_t)W * H * 3); hdr = own_hdr; }
    Job job; memset(&job, 0, sizeof job);
    job.sc = sc; job.cam = cam; job.seed = 0; job.ldr = ldr; job.hdr = hdr; job.y0 = 0; job.y1 = H; job.ref = ref;
    render_rows(&job);
    if (cam->has_tonemapper) dto_tonemap(hdr, W, H, cam->tm_key, cam->tm_burn, cam->tm_saturation, cam->tm_gamma, ldr);
    if (stats) { memset(stats, 0, sizeof *stats); stats->rays_closest = job.n_closest; stats->rays_shadow = job.n_shadow; }
    free(own_hdr); free(ref);
    return DT_OK;
}

/* Test hook: the raw streams behind the reference-RNG mode, compared in tests/test_cpu_oracle_host.py with what g++'s own
 * libstdc++ / glibc produce.  what = 0: rand() after srand(seed); 1: mt19937(seed)(); 2: uniform_real_distribution<>(0,1)
 * (= generate_canonical) over mt19937(seed); 3: uniform_int_distribution<>(0, param - 1) over mt19937(seed). */
int dto_debug_reference_rng(int what, uint32_t seed, uint32_t param, int n, double* out) {
    if (!out || n < 0) return DT_ERR_INVALID;
    if (what == 0) {
        if (n > 64) return DT_ERR_INVALID;
        GlibcRand g; glibc_rand_init(&g, seed);
        for (int i = 0; i < n; i++) out[i] = (double)glibc_rand(&g);
        return DT_OK;
    }
    Mt19937 m; mt_seed(&m, seed);
    for (int i = 0; i < n; i++) {
        if (what == 1) out[i] = (double)mt_next(&m);
        else if (what == 2) out[i] = mt_canonical(&m);
        else if (what == 3 && param > 0) out[i] = (double)mt_uniform_int(&m, param);
        else return DT_ERR_INVALID;
    }
    return DT_OK;
}

int dto_primary_hits(const dt_scene_desc* sc, const dt_camera_desc* cam, int32_t* shape, int32_t* face, float* t) {
    if (!sc || !cam) return DT_ERR_INVALID;
    Ctx c; c.sc = sc; c.cam = cam; c.rng = 1; c.n_closest = c.n_shadow = 0; c.ref = NULL;
    for (int y = 0; y < cam->height; y++) for (int x = 0; x < cam->width; x++) {
        Ray ray = generate_ray(&c, x, y);
        intersect_objects(&c, &ray);
        size_t i = (size_t)x + (size_t)y * cam->width;
        shape[i] = ray.hit.hasHit ? ray.hit.shape : -1;
        face[i] = ray.hit.hasHit ? ray.hit.face : -1;
        t[i] = ray.hit.hasHit ? ray.hit.minT : INFINITY;
    }
    return DT_OK;
}

int dto_trace_closest(const dt_scene_desc* sc, const float* origins, const float* dirs, int64_t n, int32_t* shape, int32_t* face, float* t) {
    if (!sc) return DT_ERR_INVALID;
    Ctx c; c.sc = sc; c.cam = NULL; c.rng = 1; c.n_closest = c.n_shadow = 0; c.ref = NULL;
    for (int64_t i = 0; i < n; i++) {
        Ray ray; memset(&ray, 0, sizeof ray);
        ray.origin = F3(&origins[3 * i]); ray.dir = F3(&dirs[3 * i]);
        ray.hit.minT = INFINITY; ray.hit.shape = -1; ray.hit.face = -1;
        intersect_objects(&c, &ray);
        shape[i] = ray.hit.hasHit ? ray.hit.shape : -1;
        face[i] = ray.hit.hasHit ? ray.hit.face : -1;
        t[i] = ray.hit.hasHit ? ray.hit.minT : INFINITY;
    }
    return DT_OK;
}

int dto_trace_occluded(const dt_scene_desc* sc, const float* origins, const float* dirs, const float* tmax, int64_t n, uint8_t* occluded) {
    if (!sc) return DT_ERR_INVALID;
    Ctx c; c.sc = sc; c.cam = NULL; c.rng = 1; c.n_closest = c.n_shadow = 0; c.ref = NULL;
    for (int64_t i = 0; i < n; i++) {
        Ray sr; memset(&sr, 0, sizeof sr);
        sr.origin = F3(&origins[3 * i]); sr.dir = F3(&dirs[3 * i]);
        sr.hit.hasHit = 0;
        sr.hit.minT = isinf(tmax[i]) ? INFINITY : tmax[i] + 0.01f;
        occluded[i] = (uint8_t)cast_shadow_ray(&c, &sr, tmax[i]);
    }
    return DT_OK;
}
