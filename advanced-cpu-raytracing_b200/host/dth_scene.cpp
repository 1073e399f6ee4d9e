// Host mirror of the reference's scene layer -> flat dt_scene_desc.  See include/dorktracer_host.h.
// Every routine cites the reference code whose behaviour (including float/double evaluation order) it
// restates.  Must be compiled WITHOUT FMA contraction (-ffp-contract=off, no -march=native) so that face
// normals, centroids, boxes and the BVH face permutation are bit-identical to the reference build's
// (SURVEY.md 8a "Numerics contract").
#include "../../include/dorktracer_host.h"
#include "dth_xml.h"
#include "dth_io.h"

#include <array>
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <limits>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

// ---- helperMath.h / helperMath.cpp:4-138 (float, unfused) ----
struct V3 { float x = 0, y = 0, z = 0; };
inline V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
inline V3 neg(V3 a) { return v3(a.x * -1.0f, a.y * -1.0f, a.z * -1.0f); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float len(V3 a) { return sqrtf((a.x * a.x) + (a.y * a.y) + (a.z * a.z)); }
inline V3 make_unit(V3 a) { float l = len(a); return v3(a.x / l, a.y / l, a.z / l); }
inline float comp(V3 a, int k) { return k == 0 ? a.x : k == 1 ? a.y : a.z; }

// helperMath.cpp:59-85
void orthonormal_basis(V3 r, V3& u, V3& v) {
    float ax = fabsf(r.x), ay = fabsf(r.y), az = fabsf(r.z);
    V3 rp = r;
    if (ax < ay) { if (ax < az) rp.x = 1.0f; else rp.z = 1.0f; }
    else { if (ay < az) rp.y = 1.0f; else rp.z = 1.0f; }
    u = make_unit(cross(rp, r));
    v = make_unit(cross(r, u));
}

// ---- matrix.hpp (double 4x4) ----
struct M4 { double m[4][4]; };
M4 m4_zero() { M4 r; memset(&r, 0, sizeof r); return r; }
M4 m4_identity() { M4 r = m4_zero(); for (int i = 0; i < 4; i++) r.m[i][i] = 1.0f; return r; }
M4 m4_mul(const M4& a, const M4& b) {           // matrix.hpp:123-141
    M4 r;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
        double s = 0.0f;
        for (int k = 0; k < 4; k++) s += a.m[i][k] * b.m[k][j];
        r.m[i][j] = s;
    }
    return r;
}
M4 m4_transpose(const M4& a) { M4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[j][i] = a.m[i][j]; return r; }
M4 m4_translation(double x, double y, double z) { M4 t = m4_identity(); t.m[0][3] = x; t.m[1][3] = y; t.m[2][3] = z; return t; }
M4 m4_scale(double x, double y, double z) { M4 s = m4_zero(); s.m[0][0] = x; s.m[1][1] = y; s.m[2][2] = z; s.m[3][3] = 1.0f; return s; }
M4 m4_rot_x(double a) { M4 r = m4_zero(); r.m[0][0] = r.m[3][3] = 1.0f; r.m[1][1] = r.m[2][2] = std::cos(a); r.m[1][2] = -std::sin(a); r.m[2][1] = std::sin(a); return r; }
M4 m4_rot_y(double a) { M4 r = m4_zero(); r.m[0][0] = r.m[2][2] = std::cos(a); r.m[1][1] = r.m[3][3] = 1.0f; r.m[0][2] = std::sin(a); r.m[2][0] = -std::sin(a); return r; }
M4 m4_rot_z(double a) { M4 r = m4_zero(); r.m[0][0] = r.m[1][1] = std::cos(a); r.m[2][2] = r.m[3][3] = 1.0f; r.m[1][0] = std::sin(a); r.m[0][1] = -std::sin(a); return r; }
// matrix.hpp:86-111
V3 m4_apply(const M4& t, V3 v, float w) {
    V3 r;
    r.x = (float)(t.m[0][0] * v.x + t.m[0][1] * v.y + t.m[0][2] * v.z + t.m[0][3] * w);
    r.y = (float)(t.m[1][0] * v.x + t.m[1][1] * v.y + t.m[1][2] * v.z + t.m[1][3] * w);
    r.z = (float)(t.m[2][0] * v.x + t.m[2][1] * v.y + t.m[2][2] * v.z + t.m[2][3] * w);
    return r;
}

struct Box { V3 mn, mx; };

// ---- owned storage ----
// Per-face arrays of a 10 M-triangle mesh are ~1 GB: resize() must not zero them on one thread before the parallel workers fill
// them, so these vectors default-initialise their PODs.
template <class T> struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { typedef NoInitAlloc<U> other; };
    NoInitAlloc() = default;
    template <class U> NoInitAlloc(const NoInitAlloc<U>&) {}
    template <class U> void construct(U* p) { ::new ((void*)p) U; }
    template <class U, class A0, class... A> void construct(U* p, A0&& a0, A&&... a) { ::new ((void*)p) U(std::forward<A0>(a0), std::forward<A>(a)...); }
};
template <class T> using BigVec = std::vector<T, NoInitAlloc<T>>;

struct MeshStore {
    std::shared_ptr<std::vector<float>> vertices;   // shared for inline meshes (the reference copies vertex_data per mesh)
    std::shared_ptr<std::vector<float>> uvs;
    int vertex_offset = 0, texture_offset = 0;
    BigVec<dt_face> faces;
    BigVec<V3> centers;                             // build-time only (Face::center)
    BigVec<Box> fboxes;                             // build-time only (Face::bbox)
    BigVec<dt_bvh2_node> bvh;
    bool smooth = false;                            // <Mesh shadingMode="smooth"> (the reference ignores the attribute)
    std::vector<float> vnormals;                    // smooth meshes: one unit normal per vertex of *vertices (zero where no face of this mesh touches it)
    Box bbox;
    double surface_area = 0.0;                      // the reference leaves Mesh::surfaceArea uninitialised (mesh.hpp:19); we start at 0
    V3 vertex(int id) const { const float* p = &(*vertices)[(size_t)(id - 1 + vertex_offset) * 3]; return v3(p[0], p[1], p[2]); }
};

struct CameraStore { dt_camera_desc desc; std::string image_name; };
struct ImageStore { dth::ImageData data; std::string path; int id = 0; bool loaded = false; };

}  // namespace

struct dth_scene {
    dt_scene_desc desc;
    std::vector<dt_material> materials;
    std::vector<dt_brdf> brdfs; std::vector<int> brdf_ids;
    std::vector<dt_point_light> point_lights;
    std::vector<dt_area_light> area_lights;
    std::vector<dt_directional_light> directional_lights;
    std::vector<dt_spot_light> spot_lights;
    std::vector<dt_env_light> env_lights;
    std::vector<dt_mesh_light> mesh_lights;
    std::vector<dt_image> images; std::deque<ImageStore> image_store;
    std::vector<dt_texture> textures; std::vector<int> texture_ids;
    std::vector<dt_mesh> meshes; std::deque<MeshStore> mesh_store;
    std::vector<dt_shape> mesh_shapes, sphere_shapes, shapes;
    std::vector<CameraStore> cameras;
    std::shared_ptr<std::vector<float>> vertex_data, tex_coords;
    std::vector<V3> translations, scalings; std::vector<std::array<float, 4>> rotations;
    std::string xml_dir;
    void finalize();
};

namespace {

// ---- text -> numbers.  The reference pushes GetText() into a stringstream and extracts with >>
// (parser.cpp:51-52 ...); strtof/strtol give the same correctly rounded values. ----
std::vector<float> parse_floats(const std::string& s) {
    std::vector<float> r;
    const char* p = s.c_str();
    for (;;) {
        while (*p && isspace((unsigned char)*p)) p++;
        if (!*p) break;
        char* e = nullptr;
        float v = strtof(p, &e);
        if (e == p) break;
        r.push_back(v);
        p = e;
    }
    return r;
}
std::vector<int> parse_ints(const std::string& s) {
    std::vector<int> r;
    const char* p = s.c_str();
    for (;;) {
        while (*p && isspace((unsigned char)*p)) p++;
        if (!*p) break;
        char* e = nullptr;
        long v = strtol(p, &e, 10);
        if (e == p) break;
        r.push_back((int)v);
        p = e;
    }
    return r;
}
std::string first_token(const std::string& s) { std::istringstream ss(s); std::string t; ss >> t; return t; }
float f_or(const dth::XmlNode* n, const char* child, float dflt) {
    const dth::XmlNode* c = n ? n->child(child) : nullptr;
    if (!c) return dflt;
    auto v = parse_floats(c->text);
    return v.empty() ? dflt : v[0];
}
bool v3_of(const dth::XmlNode* n, const char* child, V3& out) {
    const dth::XmlNode* c = n ? n->child(child) : nullptr;
    if (!c) return false;
    auto v = parse_floats(c->text);
    if (v.size() < 3) return false;
    out = v3(v[0], v[1], v[2]);
    return true;
}
void put3(float* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
void putm(double* d, const M4& m) { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) d[i * 4 + j] = m.m[i][j]; }

bool file_exists(const std::string& p) { FILE* f = fopen(p.c_str(), "rb"); if (!f) return false; fclose(f); return true; }
std::string resolve_path(const dth_scene& sc, const std::string& rel) {
    if (file_exists(rel)) return rel;                       // cwd-relative, as the reference (parser.cpp:1404)
    std::string a = sc.xml_dir + "/" + rel;
    if (file_exists(a)) return a;
    return rel;
}

// ---- Scene::computeFaceProperties (parser.cpp:579-611, 725-748) ----
void face_properties(const MeshStore& m, dt_face& f, V3& center, Box& fb) {
    V3 a = m.vertex(f.v0_id), b = m.vertex(f.v1_id), c = m.vertex(f.v2_id);
    center = (a + b + c) / 3.0f;                                           // computeFaceCenter
    V3 n = make_unit(cross(b - a, c - a));                                 // computeFaceNormal
    put3(f.n, n);
    fb.mn = v3(std::min(std::min(a.x, b.x), c.x), std::min(std::min(a.y, b.y), c.y), std::min(std::min(a.z, b.z), c.z));
    fb.mx = v3(std::max(std::max(a.x, b.x), c.x), std::max(std::max(a.y, b.y), c.y), std::max(std::max(a.z, b.z), c.z));
    double e1 = len(a - b), e2 = len(a - c), e3 = len(b - c);              // computeFaceArea (Heron)
    double s = (e1 + e2 + e3) / 2.0f;
    double area = std::sqrt(s * (s - e1) * (s - e2) * (s - e3));
    f.area = area;                                                         // the caller adds it to Mesh::surfaceArea, in face order
}
void add_face(MeshStore& m, int v0, int v1, int v2, Box* mesh_box) {
    dt_face f; f.v0_id = v0; f.v1_id = v1; f.v2_id = v2;
    V3 c; Box fb;
    face_properties(m, f, c, fb);
    m.surface_area += f.area;
    if (mesh_box) {                                                         // parser.cpp:1453-1460 / updateBBox :816-826
        mesh_box->mn = v3(std::min(fb.mn.x, mesh_box->mn.x), std::min(fb.mn.y, mesh_box->mn.y), std::min(fb.mn.z, mesh_box->mn.z));
        mesh_box->mx = v3(std::max(fb.mx.x, mesh_box->mx.x), std::max(fb.mx.y, mesh_box->mx.y), std::max(fb.mx.z, mesh_box->mx.z));
    }
    m.faces.push_back(f); m.centers.push_back(c); m.fboxes.push_back(fb);
}

// shadingMode="smooth" (not in the reference): vertex normal = normalised sum of the AREA-WEIGHTED normals of the
// faces of this mesh that share the vertex, accumulated in double in face order BEFORE the BVH build permutes the faces.  This is
// the convention of the course's golden renders of the *_smooth scenes (archive/hw1_outputs/akif_uslu): 43.9 / 47.7 / 54.3 dB on
// low_poly / berserker / tower against 30.4 / 34.2 / 29.2 dB with unweighted sums and 28.1 / 30.1 / 24.9 dB flat
// (tests/test_cpu_smooth_shading.py).
void compute_vertex_normals(MeshStore& m) {
    const size_t nv = m.vertices->size() / 3;
    std::vector<double> acc(nv * 3, 0.0);
    for (const dt_face& f : m.faces) {
        const int ids[3] = {f.v0_id, f.v1_id, f.v2_id};
        for (int k = 0; k < 3; k++) {
            const size_t vi = (size_t)(ids[k] - 1 + m.vertex_offset);
            if (vi >= nv) continue;
            acc[vi * 3] += (double)f.n[0] * f.area; acc[vi * 3 + 1] += (double)f.n[1] * f.area; acc[vi * 3 + 2] += (double)f.n[2] * f.area;
        }
    }
    m.vnormals.assign(nv * 3, 0.0f);
    for (size_t v = 0; v < nv; v++) {
        const double l = std::sqrt(acc[v * 3] * acc[v * 3] + acc[v * 3 + 1] * acc[v * 3 + 1] + acc[v * 3 + 2] * acc[v * 3 + 2]);
        if (l > 0.0) for (int a = 0; a < 3; a++) m.vnormals[v * 3 + a] = (float)(acc[v * 3 + a] / l);
    }
}

// ---- Mesh::ConstructBVH / RecursiveBVHBuild / RecomputeBoundingBox (mesh.cpp:23-156) ----
void bvh_recompute_box(MeshStore& m, dt_bvh2_node& node) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = 0; i < node.face_count; i++) {
        const Box& fb = m.fboxes[node.first_face + i];
        mn[0] = std::min(mn[0], fb.mn.x); mn[1] = std::min(mn[1], fb.mn.y); mn[2] = std::min(mn[2], fb.mn.z);
        mx[0] = std::max(mx[0], fb.mx.x); mx[1] = std::max(mx[1], fb.mx.y); mx[2] = std::max(mx[2], fb.mx.z);
    }
    memcpy(node.bmin, mn, 12); memcpy(node.bmax, mx, 12);
}
dth_bvh_builder g_bvh_builder = nullptr;
int32_t g_bvh_builder_min_faces = 0;
thread_local double g_bvh_seconds = 0.0;

bool bvh_build_host(MeshStore& m);
// Mesh::ConstructBVH through the registered builder (dt_bvh2_build): the builder returns the tree and the face order, the
// host applies the order to the per-face arrays.
bool bvh_build(MeshStore& m) {
    const auto t0 = std::chrono::steady_clock::now();
    bool ok = true;
    const int n = (int)m.faces.size();
    if (!g_bvh_builder || n < g_bvh_builder_min_faces || n < 2) ok = bvh_build_host(m);
    else {
        static_assert(sizeof(V3) == 12 && sizeof(Box) == 24, "V3 / Box must be plain float triples");
        BigVec<uint32_t> order((size_t)n);
        m.bvh.resize((size_t)n * 2 - 1);                                    // filled by the builder (not zeroed first: ~800 MB on config 5)
        uint32_t n_nodes = 0;
        float mn[3], mx[3]; put3(mn, m.bbox.mn); put3(mx, m.bbox.mx);
        float ms_device = 0.f;
        const int rc = g_bvh_builder(n, &m.centers[0].x, &m.fboxes[0].mn.x, mn, mx, order.data(), m.bvh.data(), (uint32_t)m.bvh.size(), &n_nodes, &ms_device);
        const auto t1 = std::chrono::steady_clock::now();
        if (rc != 0) { g_err = "BVH builder failed with status " + std::to_string(rc); ok = false; }
        else {
            m.bvh.resize(n_nodes);
            BigVec<dt_face> f((size_t)n); BigVec<V3> c((size_t)n); BigVec<Box> b((size_t)n);
            dth::parallel_for((size_t)n, 1 << 15, [&](size_t lo, size_t hi) {
                for (size_t i = lo; i < hi; i++) { f[i] = m.faces[order[i]]; c[i] = m.centers[order[i]]; b[i] = m.fboxes[order[i]]; }
            });
            m.faces.swap(f); m.centers.swap(c); m.fboxes.swap(b);
            if (getenv("DTH_DEBUG_TIMING"))
                fprintf(stderr, "[dth] BVH builder: call %.3f s (device work %.3f s, the rest is allocation + copies), face permutation %.3f s\n",
                        std::chrono::duration<double>(t1 - t0).count(), ms_device * 1e-3, std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
        }
    }
    g_bvh_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return ok;
}
bool bvh_build_host(MeshStore& m) {
    int n = (int)m.faces.size();
    if (n <= 0) { m.bvh.clear(); return true; }
    m.bvh.resize((size_t)n * 2 - 1);
    for (auto& nd : m.bvh) { nd.left = nd.right = -1; nd.first_face = nd.face_count = 0; memset(nd.bmin, 0, 12); memset(nd.bmax, 0, 12); }
    dt_bvh2_node& root = m.bvh[0];
    put3(root.bmin, m.bbox.mn); put3(root.bmax, m.bbox.mx);
    root.first_face = 0; root.face_count = (uint32_t)n;
    uint32_t next_free = 1;
    // The reference recurses left-first and allocates both children before descending; an explicit stack
    // with the right child pushed first reproduces the same node numbering.
    std::vector<uint32_t> stack; stack.push_back(0);
    while (!stack.empty()) {
        uint32_t ni = stack.back(); stack.pop_back();
        dt_bvh2_node& node = m.bvh[ni];
        if (node.face_count < 2) continue;
        float lenX = node.bmax[0] - node.bmin[0], lenY = node.bmax[1] - node.bmin[1], lenZ = node.bmax[2] - node.bmin[2];
        float split; int axis;
        if (lenX > lenY) { if (lenX > lenZ) { split = node.bmin[0] + lenX * 0.5f; axis = 0; } else { split = node.bmin[2] + lenZ * 0.5f; axis = 2; } }
        else { if (lenY > lenZ) { split = node.bmin[1] + lenY * 0.5f; axis = 1; } else { split = node.bmin[2] + lenZ * 0.5f; axis = 2; } }
        int i = (int)node.first_face, j = i + (int)node.face_count - 1;
        while (i <= j) {
            if (comp(m.centers[i], axis) < split) i++;
            else { std::swap(m.faces[i], m.faces[j]); std::swap(m.centers[i], m.centers[j]); std::swap(m.fboxes[i], m.fboxes[j]); j--; }
        }
        int left_count = i - (int)node.first_face;
        if (left_count == 0 || left_count == (int)node.face_count) continue;
        uint32_t li = next_free++, ri = next_free++;
        m.bvh[li].first_face = node.first_face; m.bvh[li].face_count = (uint32_t)left_count;
        m.bvh[ri].first_face = (uint32_t)i; m.bvh[ri].face_count = node.face_count - (uint32_t)left_count;
        node.left = (int32_t)li; node.right = (int32_t)ri; node.face_count = 0;
        bvh_recompute_box(m, m.bvh[li]); bvh_recompute_box(m, m.bvh[ri]);
        stack.push_back(ri); stack.push_back(li);
    }
    m.bvh.resize(next_free);   // unused tail of the 2n-1 allocation is never referenced
    return true;
}

// ---- Scene::computeTransform (parser.cpp:651-723): raw-text indexing, single-digit ids ----
struct Xform { M4 transform, inverse, inverse_transpose; };
bool compute_transform(const dth_scene& sc, Xform& x, const std::string& str) {
    std::vector<M4> inv;
    size_t idx = 0;
    while (str.size() >= 1 && idx < str.size() - 1) {
        char c = str[idx];
        int id = (int)(str[idx + 1] - '0');
        if (c == 'r') {
            if (id < 1 || id > (int)sc.rotations.size()) { g_err = "rotation id out of range in '" + str + "'"; return false; }
            auto rot = sc.rotations[id - 1];      // {angle, x, y, z}
            float angle = (float)(rot[0] * (M_PI / 180.0f));
            M4 r = m4_zero(), ir = m4_zero();
            bool ok = false;
            if (rot[1] >= 0.99 && rot[2] <= 0.001 && rot[3] <= 0.0001) { r = m4_rot_x(angle); ir = m4_rot_x(-angle); ok = true; }
            if (rot[2] >= 0.99 && rot[1] <= 0.001 && rot[3] <= 0.0001) { r = m4_rot_y(angle); ir = m4_rot_y(-angle); ok = true; }
            if (rot[3] >= 0.99 && rot[1] <= 0.001 && rot[2] <= 0.0001) { r = m4_rot_z(angle); ir = m4_rot_z(-angle); ok = true; }
            if (!ok) { g_err = "only axis-aligned rotation axes are supported (parser.cpp:672-683)"; return false; }
            inv.push_back(ir); x.transform = m4_mul(r, x.transform);
        } else if (c == 't') {
            if (id < 1 || id > (int)sc.translations.size()) { g_err = "translation id out of range in '" + str + "'"; return false; }
            V3 t = sc.translations[id - 1];
            inv.push_back(m4_translation(-t.x, -t.y, -t.z));
            x.transform = m4_mul(m4_translation(t.x, t.y, t.z), x.transform);
        } else if (c == 's') {
            if (id < 1 || id > (int)sc.scalings.size()) { g_err = "scaling id out of range in '" + str + "'"; return false; }
            V3 s = sc.scalings[id - 1];
            inv.push_back(m4_scale(1.0f / s.x, 1.0f / s.y, 1.0f / s.z));
            x.transform = m4_mul(m4_scale(s.x, s.y, s.z), x.transform);
        }
        idx += 3;
    }
    x.inverse = m4_identity();
    for (auto& t : inv) x.inverse = m4_mul(x.inverse, t);
    x.inverse_transpose = m4_transpose(x.inverse);
    return true;
}

// ---- Scene::transformBoundingBox (parser.cpp:749-805) ----
Box transform_box(const Box& o, const M4& t) {
    Box res; res.mx = v3(-INFINITY, -INFINITY, -INFINITY); res.mn = v3(INFINITY, INFINITY, INFINITY);
    V3 ext = o.mx - o.mn;
    V3 corners[8];
    corners[0] = o.mx; corners[1] = o.mn;
    V3 c = o.mx; c.x -= ext.x; corners[2] = c;
    c = o.mx; c.y -= ext.y; corners[3] = c;
    c = o.mx; c.z -= ext.z; corners[4] = c;
    c = o.mx; c.x -= ext.x; c.y -= ext.y; corners[5] = c;
    c = o.mx; c.x -= ext.x; c.z -= ext.z; corners[6] = c;
    c = o.mx; c.y -= ext.y; c.z -= ext.z; corners[7] = c;
    for (int i = 0; i < 8; i++) {
        V3 p = m4_apply(t, corners[i], 1.0f);
        res.mx = v3(std::max(p.x, res.mx.x), std::max(p.y, res.mx.y), std::max(p.z, res.mx.z));
        res.mn = v3(std::min(p.x, res.mn.x), std::min(p.y, res.mn.y), std::min(p.z, res.mn.z));
    }
    return res;
}

// ---- Scene::SetupTextures (parser.cpp:613-650) ----
void setup_textures(const dth_scene& sc, dt_shape& sh, std::string ids) {
    ids += " ";
    size_t last = 0, next = 0;
    while ((next = ids.find(" ", last)) != std::string::npos) {
        std::string s = ids.substr(last, next - last);
        char* e = nullptr;
        const char* b = s.c_str();
        long id = strtol(b, &e, 10);
        if (e == b) break;                          // std::stoi would throw here; we stop instead
        int ti = -1;
        for (size_t i = 0; i < sc.texture_ids.size(); i++) if (sc.texture_ids[i] == (int)id) { ti = (int)i; break; }
        if (ti < 0) break;
        switch (sc.textures[ti].decal_mode) {       // Texture::SetTextureType (texture.h:59-81)
            case DT_DECAL_REPLACE_KD: case DT_DECAL_BLEND_KD: sh.tex_diffuse = ti; break;
            case DT_DECAL_REPLACE_KS: sh.tex_specular = ti; break;
            case DT_DECAL_REPLACE_ALL: sh.tex_replace_all = ti; break;
            case DT_DECAL_REPLACE_NORMAL: sh.tex_normal = ti; break;
            case DT_DECAL_BUMP_NORMAL: sh.tex_bump = ti; break;
            default: break;                         // replace_background: `type` is left uninitialised in the reference
        }
        last = next + 1;
    }
}

dt_shape blank_shape() {
    dt_shape s; memset(&s, 0, sizeof s);
    s.mesh = -1; s.base_shape = -1;
    s.tex_diffuse = s.tex_specular = s.tex_normal = s.tex_bump = s.tex_replace_all = -1;
    return s;
}
void set_xform(dt_shape& s, const Xform& x) { putm(s.transform, x.transform); putm(s.inverse_transform, x.inverse); putm(s.inverse_transpose_transform, x.inverse_transpose); }
M4 getm(const double* d) { M4 m; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m.m[i][j] = d[i * 4 + j]; return m; }

bool parse_motion_blur(const dth::XmlNode* el, dt_shape& sh) {
    V3 mb;
    if (v3_of(el, "MotionBlur", mb)) { sh.has_motion_blur = 1; put3(sh.motion_blur, mb); return true; }
    sh.has_motion_blur = 0; sh.motion_blur[0] = sh.motion_blur[1] = sh.motion_blur[2] = 0.f;
    return false;
}

// ---- camera.cpp:5-72 ----
void camera_image_plane(dt_camera_desc& c) {            // CalculateImagePlaneParams
    V3 pos = v3(c.position[0], c.position[1], c.position[2]);
    V3 gaze = v3(c.gaze[0], c.gaze[1], c.gaze[2]);
    V3 up = v3(c.up[0], c.up[1], c.up[2]);
    V3 w = neg(gaze);
    V3 right = cross(up, w);
    V3 middle = pos + gaze * c.near_dist;
    V3 q = middle + right * c.left + up * c.top;
    put3(c.right, right); put3(c.q, q);
}
void camera_default(dt_camera_desc& c, V3 pos, V3 gaze_dir, V3 up_dir, const float np[4], float near_dist, int w, int h) {
    put3(c.position, pos); c.near_dist = near_dist; c.width = w; c.height = h;
    c.left = np[0]; c.right_ = np[1]; c.bottom = np[2]; c.top = np[3];
    V3 gaze = make_unit(gaze_dir);
    V3 tmp_up = make_unit(up_dir);
    float dotvu = dot(tmp_up, gaze);                    // GetOrthonormal (camera.cpp:51-59)
    float rel = dot(gaze, gaze);
    V3 proj = gaze * (dotvu / rel);
    V3 up = make_unit(tmp_up - proj);
    put3(c.gaze, gaze); put3(c.up, up);
    camera_image_plane(c);
}
void camera_look_at(dt_camera_desc& c, V3 pos, V3 gaze_point, V3 up_dir, float near_dist, float fov_y, int w, int h) {
    put3(c.position, pos); c.near_dist = near_dist; c.width = w; c.height = h;
    float aspect = (float)w / h;
    c.top = (float)(near_dist * std::tan((fov_y * (M_PI / 180.0f) / 2.0f)));
    c.right_ = c.top * aspect;
    c.bottom = -c.top; c.left = -c.right_;
    V3 gaze = make_unit(gaze_point - pos);
    V3 tmp_up = make_unit(up_dir);
    V3 tmp_right = make_unit(cross(tmp_up, gaze));
    V3 up = make_unit(cross(gaze, tmp_right));
    put3(c.gaze, gaze); put3(c.up, up);
    camera_image_plane(c);
}

int find_image(const dth_scene& sc, int id) {
    for (size_t i = 0; i < sc.image_store.size(); i++) if (sc.image_store[i].id == id) return (int)i;
    return -1;
}

bool parse_scene(dth_scene& sc, const dth::XmlNode* root) {
    dt_scene_desc& d = sc.desc;
    memset(&d, 0, sizeof d);
    d.abi_version = DT_ABI_VERSION;
    d.bg_texture = -1;
    d.shadow_ray_epsilon = 0.001f;                      // scene.cpp:4
    if (auto e = root->child("BackgroundColor")) { auto v = parse_ints(e->text); for (int k = 0; k < 3 && k < (int)v.size(); k++) d.background_color[k] = v[k]; }
    if (auto e = root->child("ShadowRayEpsilon")) { auto v = parse_floats(e->text); if (!v.empty()) d.shadow_ray_epsilon = v[0]; }
    d.max_recursion_depth = 0;
    if (auto e = root->child("MaxRecursionDepth")) { auto v = parse_ints(e->text); if (!v.empty()) d.max_recursion_depth = v[0]; }

    // ---- parseCameras (parser.cpp:1498-1636); `camera` persists across iterations like the reference's ----
    {
        auto cams = root->child("Cameras");
        if (!cams) { g_err = "scene has no <Cameras>"; return false; }
        dt_camera_desc cam; memset(&cam, 0, sizeof cam);
        for (auto el : cams->children_named("Camera")) {
            bool look_at = el->attr_is("type", "lookAt");
            V3 pos, up;
            if (!v3_of(el, "Position", pos) || !v3_of(el, "Up", up)) { g_err = "camera lacks Position/Up"; return false; }
            float near_dist = f_or(el, "NearDistance", 1.f);
            auto res = el->child("ImageResolution") ? parse_floats(el->child("ImageResolution")->text) : std::vector<float>();
            if (res.size() < 2) { g_err = "camera lacks ImageResolution"; return false; }
            int w = (int)res[0], h = (int)res[1];
            CameraStore cs;
            cs.image_name = el->child("ImageName") ? first_token(el->child("ImageName")->text) : "out.png";
            if (look_at) {
                V3 gp;
                if (!v3_of(el, "GazePoint", gp) && !v3_of(el, "Gaze", gp)) { g_err = "lookAt camera lacks GazePoint"; return false; }
                float fov = f_or(el, "FovY", 45.f);
                camera_look_at(cam, pos, gp, up, near_dist, fov, w, h);
            } else {
                V3 gd;
                if (!v3_of(el, "Gaze", gd)) { g_err = "camera lacks Gaze"; return false; }
                auto np = el->child("NearPlane") ? parse_floats(el->child("NearPlane")->text) : std::vector<float>();
                if (np.size() < 4) { g_err = "camera lacks NearPlane"; return false; }
                camera_default(cam, pos, gd, up, np.data(), near_dist, w, h);
            }
            cam.samples_per_pixel = 1;
            if (auto c = el->child("NumSamples")) { auto v = parse_ints(c->text); if (!v.empty()) cam.samples_per_pixel = v[0]; }
            cam.focus_distance = f_or(el, "FocusDistance", 0.f);
            cam.aperture_size = f_or(el, "ApertureSize", 0.f);
            if (auto c = el->child("Renderer")) {
                if (first_token(c->text) == "PathTracing") {
                    bool is = false, rr = false, nee = false;
                    if (auto pz = el->child("RendererParams")) {
                        std::istringstream ss(pz->text); std::string p;
                        while (ss >> p) { if (p == "NextEventEstimation") nee = true; else if (p == "RussianRoulette") rr = true; else if (p == "ImportanceSampling") is = true; }
                    }
                    cam.path_tracing = 1; cam.importance_sampling = is; cam.next_event_estimation = nee; cam.russian_roulette = rr;
                }
            }
            if (auto tm = el->child("Tonemap")) {          // parseTonemapper (parser.cpp:828-869)
                cam.has_tonemapper = 1;
                cam.tm_key = 0.18f; cam.tm_burn = 1.0f; cam.tm_saturation = 1.0f; cam.tm_gamma = 2.2f;
                if (auto o = tm->child("TMOOptions")) { auto v = parse_floats(o->text); if (v.size() > 0) cam.tm_key = v[0]; if (v.size() > 1) cam.tm_burn = v[1]; }
                cam.tm_saturation = f_or(tm, "Saturation", 1.0f);
                cam.tm_gamma = f_or(tm, "Gamma", 2.2f);
            }
            cs.desc = cam;
            sc.cameras.push_back(cs);
        }
    }

    // ---- parseLights (parser.cpp:984-1107) ----
    if (auto lights = root->child("Lights")) {
        V3 amb;
        if (v3_of(lights, "AmbientLight", amb)) put3(d.ambient_light, amb);
        for (auto l : lights->children_named("PointLight")) {
            dt_point_light pl; V3 p, i;
            if (!v3_of(l, "Position", p) || !v3_of(l, "Intensity", i)) { g_err = "PointLight lacks Position/Intensity"; return false; }
            put3(pl.position, p); put3(pl.intensity, i); sc.point_lights.push_back(pl);
        }
        for (auto l : lights->children_named("AreaLight")) {
            dt_area_light al; V3 p, n, r;
            if (!v3_of(l, "Position", p) || !v3_of(l, "Normal", n) || !v3_of(l, "Radiance", r)) { g_err = "AreaLight incomplete"; return false; }
            put3(al.position, p); put3(al.normal, n); put3(al.radiance, r);
            al.extent = f_or(l, "Size", 1.f);
            V3 u, v; orthonormal_basis(n, u, v);           // areaLight.h:31
            put3(al.u, u); put3(al.v, v);
            sc.area_lights.push_back(al);
        }
        for (auto l : lights->children_named("DirectionalLight")) {
            dt_directional_light dl; V3 dir, r;
            if (!v3_of(l, "Direction", dir) || !v3_of(l, "Radiance", r)) { g_err = "DirectionalLight incomplete"; return false; }
            put3(dl.dir, make_unit(dir)); put3(dl.radiance, r); sc.directional_lights.push_back(dl);
        }
        for (auto l : lights->children_named("SpotLight")) {
            dt_spot_light sl; V3 p, dir, i;
            if (!v3_of(l, "Position", p) || !v3_of(l, "Direction", dir) || !v3_of(l, "Intensity", i)) { g_err = "SpotLight incomplete"; return false; }
            put3(sl.pos, p); put3(sl.dir, make_unit(dir)); put3(sl.intensity, i);
            sl.coverage_angle = f_or(l, "CoverageAngle", 0.f); sl.falloff_angle = f_or(l, "FalloffAngle", 0.f);
            sl.cos_half_coverage = std::cos((sl.coverage_angle * (M_PI / 180.0f) / 2.0f));    // spotLight.h:29-30
            sl.cos_half_falloff = std::cos((sl.falloff_angle * (M_PI / 180.0f) / 2.0f));
            sc.spot_lights.push_back(sl);
        }
    }

    // ---- parseBRDFs (parser.cpp:870-982): push order = ModifiedBlinnPhong, OriginalBlinnPhong, OriginalPhong, ModifiedPhong, TorranceSparrow
    if (auto bs = root->child("BRDFs")) {
        auto add = [&](const char* tag, int kind, const char* flag_attr) {
            for (auto b : bs->children_named(tag)) {
                dt_brdf br; br.kind = kind; br.exponent = f_or(b, "Exponent", 0.f);
                br.flag = (flag_attr && b->attr_is(flag_attr, "true")) ? 1 : 0;
                if (kind == DT_BRDF_TORRANCE_SPARROW) { /* isEnergyConserving is always true; flag = kdFresnel */ }
                int id = -1; if (b->attr("id")) id = atoi(b->attr("id"));
                sc.brdfs.push_back(br); sc.brdf_ids.push_back(id);
            }
        };
        add("ModifiedBlinnPhong", DT_BRDF_MODIFIED_BLINN_PHONG, "normalized");
        add("OriginalBlinnPhong", DT_BRDF_BLINN_PHONG, nullptr);
        add("OriginalPhong", DT_BRDF_PHONG, nullptr);
        add("ModifiedPhong", DT_BRDF_MODIFIED_PHONG, "normalized");
        add("TorranceSparrow", DT_BRDF_TORRANCE_SPARROW, "kdfresnel");
    }

    // ---- parseMaterials (parser.cpp:1109-1278): `material` persists across iterations (fields not given keep
    // the previous material's values, exactly like the reference's loop variable) ----
    {
        auto mats = root->child("Materials");
        if (!mats) { g_err = "scene has no <Materials>"; return false; }
        dt_material m; memset(&m, 0, sizeof m); m.brdf = -1; m.type = DT_MAT_DEFAULT;
        for (auto el : mats->children_named("Material")) {
            if (const char* b = el->attr("BRDF")) {
                int id = atoi(b); int bi = -1;
                for (size_t i = 0; i < sc.brdf_ids.size(); i++) if (sc.brdf_ids[i] == id) { bi = (int)i; break; }
                m.brdf = bi;
            }
            if (el->attr_is("type", "mirror")) m.type = DT_MAT_MIRROR;
            else if (el->attr_is("type", "dielectric")) m.type = DT_MAT_DIELECTRIC;
            else if (el->attr_is("type", "conductor")) m.type = DT_MAT_CONDUCTOR;
            else m.type = DT_MAT_DEFAULT;
            bool degamma = el->attr_is("degamma", "true");
            float gamma = 2.2f;
            auto rd = [&](const char* tag, float* dst) -> bool {
                V3 v; if (!v3_of(el, tag, v)) return false;
                if (degamma) { v.x = std::pow(v.x, gamma); v.y = std::pow(v.y, gamma); v.z = std::pow(v.z, gamma); }
                put3(dst, v); return true;
            };
            rd("AmbientReflectance", m.ambient);
            rd("DiffuseReflectance", m.diffuse);
            rd("SpecularReflectance", m.specular);
            if (!rd("MirrorReflectance", m.mirror)) m.mirror[0] = m.mirror[1] = m.mirror[2] = 0.f;
            m.refractive_index = f_or(el, "RefractionIndex", 1.0f);
            V3 ac; if (v3_of(el, "AbsorptionCoefficient", ac)) put3(m.absorption_coefficient, ac); else m.absorption_coefficient[0] = m.absorption_coefficient[1] = m.absorption_coefficient[2] = 0.f;
            m.conductor_absorption_index = f_or(el, "AbsorptionIndex", 0.0f);
            m.phong_exponent = f_or(el, "PhongExponent", 1.0f);
            m.roughness = f_or(el, "Roughness", 0.0f);
            sc.materials.push_back(m);
        }
    }

    // ---- Textures (parser.cpp:85-228) ----
    if (auto tex = root->child("Textures")) {
        if (auto imgs = tex->child("Images")) {
            for (auto im : imgs->children_named("Image")) {
                sc.image_store.emplace_back();
                ImageStore& is = sc.image_store.back();
                is.id = im->attr("id") ? atoi(im->attr("id")) : 0;
                std::string fn = first_token(im->text);
                is.path = fn;
                std::string p1 = resolve_path(sc, "inputs/" + fn);          // parser.cpp:107,110
                if (!file_exists(p1)) p1 = resolve_path(sc, fn);
                std::string err;
                bool hdr = fn.find(".exr") != std::string::npos;
                bool ok = false;
                if (hdr) ok = dth::exr_load(p1, is.data, err);
                else if (fn.size() > 4 && (fn.substr(fn.size() - 4) == ".png" || fn.substr(fn.size() - 4) == ".PNG")) ok = dth::png_load(p1, is.data, err);
                is.loaded = ok;
                if (!ok) {   // placeholder 1x1; the caller may supply pixels with dth_scene_set_image
                    is.data.width = is.data.height = 1; is.data.is_hdr = hdr; is.data.channels = 3;
                    is.data.u8.assign(8, 0); is.data.f32.assign(8, 0.f);
                }
            }
        }
        for (auto tm : tex->children_named("TextureMap")) {
            dt_texture t; memset(&t, 0, sizeof t);
            int id = tm->attr("id") ? atoi(tm->attr("id")) : 0;
            std::string type = tm->attr("type") ? tm->attr("type") : "";
            std::string mode = tm->child("DecalMode") ? first_token(tm->child("DecalMode")->text) : "";
            t.decal_mode = mode == "replace_kd" ? DT_DECAL_REPLACE_KD : mode == "blend_kd" ? DT_DECAL_BLEND_KD :
                           mode == "replace_ks" ? DT_DECAL_REPLACE_KS : mode == "replace_background" ? DT_DECAL_REPLACE_BG :
                           mode == "replace_normal" ? DT_DECAL_REPLACE_NORMAL : mode == "bump_normal" ? DT_DECAL_BUMP_NORMAL :
                           mode == "replace_all" ? DT_DECAL_REPLACE_ALL : DT_DECAL_REPLACE_KD;
            t.image = -1;
            if (type == "image") {
                t.kind = DT_TEX_IMAGE;
                int image_id = 0;
                if (auto c = tm->child("ImageId")) { auto v = parse_ints(c->text); if (!v.empty()) image_id = v[0]; }
                std::string interp = "nearest";
                if (auto c = tm->child("Interpolation")) interp = first_token(c->text);
                t.interpolation = interp == "nearest" ? DT_INTERP_NEAREST : DT_INTERP_BILINEAR;   // imageTexture.h:24-27
                t.normalizer = f_or(tm, "Normalizer", 255.0f);
                t.sample_multiplier = f_or(tm, "BumpFactor", 1.0f);
                t.image = find_image(sc, image_id);
                if (t.image < 0) { g_err = "TextureMap references unknown ImageId"; return false; }
            } else if (type == "perlin") {
                t.kind = DT_TEX_PERLIN;
                std::string conv = "linear";
                if (auto c = tm->child("NoiseConversion")) conv = first_token(c->text);
                t.noise_conversion = conv == "absval" ? DT_NOISE_ABSVAL : DT_NOISE_LINEAR;
                t.noise_scale = f_or(tm, "NoiseScale", 1.0f);
                t.sample_multiplier = f_or(tm, "BumpFactor", 1.0f);
                t.normalizer = 1.0f;
            } else continue;                                                  // "checkerboard": not implemented in the reference either
            sc.textures.push_back(t); sc.texture_ids.push_back(id);
            if (mode == "replace_background") d.bg_texture = (int)sc.textures.size() - 1;
        }
    }
    // SphericalDirectionalLight (parser.cpp:232-261)
    if (auto lights = root->child("Lights")) {
        for (auto l : lights->children_named("SphericalDirectionalLight")) {
            int image_id = 0;
            if (auto c = l->child("ImageId")) { auto v = parse_ints(c->text); if (!v.empty()) image_id = v[0]; }
            dt_env_light e; e.image = find_image(sc, image_id);
            if (e.image < 0) { g_err = "SphericalDirectionalLight references unknown ImageId"; return false; }
            sc.env_lights.push_back(e);
        }
    }

    // ---- VertexData / TexCoordData / Transformations (parser.cpp:264-345) ----
    sc.vertex_data = std::make_shared<std::vector<float>>();
    sc.tex_coords = std::make_shared<std::vector<float>>();
    if (auto e = root->child("VertexData")) { auto v = parse_floats(e->text); v.resize(v.size() / 3 * 3); *sc.vertex_data = v; }
    if (auto e = root->child("TexCoordData")) { auto v = parse_floats(e->text); v.resize(v.size() / 2 * 2); *sc.tex_coords = v; }
    if (auto tr = root->child("Transformations")) {
        for (auto c : tr->children_named("Translation")) { auto v = parse_floats(c->text); v.resize(3); sc.translations.push_back(v3(v[0], v[1], v[2])); }
        for (auto c : tr->children_named("Scaling")) { auto v = parse_floats(c->text); v.resize(3); sc.scalings.push_back(v3(v[0], v[1], v[2])); }
        for (auto c : tr->children_named("Rotation")) { auto v = parse_floats(c->text); v.resize(4); sc.rotations.push_back({v[0], v[1], v[2], v[3]}); }
    }

    auto objects = root->child("Objects");
    if (!objects) { g_err = "scene has no <Objects>"; return false; }

    // ---- parseMeshes("Mesh") then parseMeshes("LightMesh") (parser.cpp:1280-1496) ----
    for (int pass = 0; pass < 2; pass++) {
        const char* tag = pass == 0 ? "Mesh" : "LightMesh";
        for (auto el : objects->children_named(tag)) {
            auto faces_el = el->child("Faces");
            if (!faces_el) { g_err = std::string(tag) + " lacks <Faces>"; return false; }
            const char* ply = faces_el->attr("plyFile");
            sc.mesh_store.emplace_back();
            MeshStore& ms = sc.mesh_store.back();
            if (ply) { ms.vertices = std::make_shared<std::vector<float>>(); ms.uvs = std::make_shared<std::vector<float>>(); }
            else { ms.vertices = sc.vertex_data; ms.uvs = sc.tex_coords; }
            dt_shape sh = blank_shape();
            sh.kind = DT_SHAPE_MESH;
            sh.id = el->attr("id") ? atoi(el->attr("id")) : 0;
            V3 radiance;
            if (pass == 1) v3_of(el, "Radiance", radiance);
            if (auto c = el->child("Textures")) setup_textures(sc, sh, c->text);
            Xform x; x.transform = m4_identity(); x.inverse = m4_identity(); x.inverse_transpose = m4_identity();
            if (auto c = el->child("Transformations")) { if (!compute_transform(sc, x, c->text)) return false; }
            set_xform(sh, x);
            if (auto c = el->child("Material")) { auto v = parse_ints(c->text); sh.material = v.empty() ? 0 : v[0]; }
            else { g_err = std::string(tag) + " lacks <Material>"; return false; }
            parse_motion_blur(el, sh);
            if (const char* a = faces_el->attr("vertexOffset")) ms.vertex_offset = atoi(a);
            if (const char* a = faces_el->attr("textureOffset")) ms.texture_offset = atoi(a);
            Box bbox;
            bbox.mx = v3(FLT_MIN, FLT_MIN, FLT_MIN);                           // parser.cpp:1393 (numeric_limits<float>::min())
            bbox.mn = v3(FLT_MAX, FLT_MAX, FLT_MAX);
            if (ply) {
                dth::PlyMesh pm; std::string err;
                const auto t_ply0 = std::chrono::steady_clock::now();
                if (!dth::ply_load(resolve_path(sc, ply), pm, err)) { g_err = err; return false; }
                const auto t_ply1 = std::chrono::steady_clock::now();
                if (pm.streamed) {
                    // the decoded arrays ARE the mesh arrays (no per-mesh copy of the vertex data, mesh.cpp:7-13), and the
                    // per-face properties (parser.cpp:579-611) are filled in place by parallel workers; only Mesh::surfaceArea, a sum
                    // whose order the light weights depend on, is accumulated sequentially.
                    ms.vertices->swap(pm.positions_f32);
                    const size_t nf = pm.triangles.size() / 3;
                    ms.faces.resize(nf); ms.centers.resize(nf); ms.fboxes.resize(nf);
                    std::mutex box_mu;
                    dth::parallel_for(nf, 1 << 15, [&](size_t b, size_t e) {
                        Box local = bbox;
                        for (size_t i = b; i < e; i++) {
                            dt_face f; f.v0_id = pm.triangles[i * 3] + 1; f.v1_id = pm.triangles[i * 3 + 1] + 1; f.v2_id = pm.triangles[i * 3 + 2] + 1;
                            face_properties(ms, f, ms.centers[i], ms.fboxes[i]);
                            ms.faces[i] = f;
                            const Box& fb = ms.fboxes[i];
                            local.mn = v3(std::min(fb.mn.x, local.mn.x), std::min(fb.mn.y, local.mn.y), std::min(fb.mn.z, local.mn.z));
                            local.mx = v3(std::max(fb.mx.x, local.mx.x), std::max(fb.mx.y, local.mx.y), std::max(fb.mx.z, local.mx.z));
                        }
                        std::lock_guard<std::mutex> lk(box_mu);
                        bbox.mn = v3(std::min(local.mn.x, bbox.mn.x), std::min(local.mn.y, bbox.mn.y), std::min(local.mn.z, bbox.mn.z));
                        bbox.mx = v3(std::max(local.mx.x, bbox.mx.x), std::max(local.mx.y, bbox.mx.y), std::max(local.mx.z, bbox.mx.z));
                    });
                    for (size_t i = 0; i < nf; i++) ms.surface_area += ms.faces[i].area;
                    if (getenv("DTH_DEBUG_TIMING"))
                        fprintf(stderr, "[dth] %s: read + decode %.3f s, face properties %.3f s (%zu faces, streamed)\n", ply, std::chrono::duration<double>(t_ply1 - t_ply0).count(),
                                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_ply1).count(), nf);
                }
                if (!pm.streamed) ms.vertices->resize(pm.positions.size());
                for (size_t i = 0; i < pm.positions.size(); i++) (*ms.vertices)[i] = (float)pm.positions[i];
                size_t nf = pm.face_counts.size();
                ms.faces.reserve(nf); ms.centers.reserve(nf); ms.fboxes.reserve(nf);
                size_t off = 0;
                for (size_t i = 0; i < nf; i++) {
                    int cnt = pm.face_counts[i]; const int* ix = &pm.face_indices[off]; off += (size_t)cnt;
                    if (cnt == 3) add_face(ms, ix[0] + 1, ix[1] + 1, ix[2] + 1, &bbox);
                    else if (cnt == 4) { add_face(ms, ix[0] + 1, ix[1] + 1, ix[2] + 1, &bbox); add_face(ms, ix[2] + 1, ix[3] + 1, ix[0] + 1, &bbox); }
                }
            } else {
                auto ids = parse_ints(faces_el->text);
                int nverts = (int)(ms.vertices->size() / 3);
                for (size_t i = 0; i + 2 < ids.size(); i += 3) {
                    for (int k = 0; k < 3; k++) { int vi = ids[i + k] - 1 + ms.vertex_offset; if (vi < 0 || vi >= nverts) { g_err = "face vertex id out of range"; return false; } }
                    add_face(ms, ids[i], ids[i + 1], ids[i + 2], &bbox);
                }
            }
            if (ms.faces.empty()) { g_err = std::string(tag) + " has no faces"; return false; }
            ms.bbox = bbox;
            ms.smooth = el->attr_is("shadingMode", "smooth");
            if (ms.smooth) compute_vertex_normals(ms);
            if (!bvh_build(ms)) return false;
            sh.mesh = (int)sc.mesh_store.size() - 1;
            sc.mesh_shapes.push_back(sh);
            if (pass == 1) {
                dt_mesh_light ml; ml.shape = (int)sc.mesh_shapes.size() - 1; ml.id = sh.id; put3(ml.radiance, radiance);
                sc.mesh_lights.push_back(ml);
                if (sh.material < 1 || sh.material > (int)sc.materials.size()) { g_err = "LightMesh material out of range"; return false; }
                dt_material& mat = sc.materials[sh.material - 1];
                mat.type = DT_MAT_EMISSIVE; put3(mat.radiance, radiance);
            }
        }
    }

    // ---- MeshInstances (parser.cpp:350-455) ----
    for (auto el : objects->children_named("MeshInstance")) {
        bool reset = el->attr_is("resetTransform", "true");
        int own_id = el->attr("id") ? atoi(el->attr("id")) : 0;
        int base_id = el->attr("baseMeshId") ? atoi(el->attr("baseMeshId")) : 0;
        int parent = -1;
        for (size_t i = 0; i < sc.mesh_shapes.size(); i++) if (sc.mesh_shapes[i].id == base_id) parent = (int)i;   // last match wins
        if (parent < 0) { g_err = "MeshInstance references unknown baseMeshId"; return false; }
        int base = parent;
        while (sc.mesh_shapes[base].kind == DT_SHAPE_INSTANCE) base = sc.mesh_shapes[base].base_shape;
        dt_shape sh = blank_shape();
        sh.kind = DT_SHAPE_INSTANCE; sh.id = own_id; sh.base_shape = base;
        if (auto c = el->child("Textures")) setup_textures(sc, sh, c->text);
        if (auto c = el->child("Material")) { auto v = parse_ints(c->text); sh.material = v.empty() ? 0 : v[0]; }
        else sh.material = sc.mesh_shapes[base].material;
        parse_motion_blur(el, sh);
        Xform x; x.transform = m4_identity(); x.inverse = m4_identity(); x.inverse_transpose = m4_zero();   // Matrix(4,4) is all-zero until set
        if (auto c = el->child("Transformations")) {
            if (!compute_transform(sc, x, c->text)) return false;
            if (!reset) {
                M4 pt = getm(sc.mesh_shapes[parent].transform), pi = getm(sc.mesh_shapes[parent].inverse_transform);
                x.transform = m4_mul(x.transform, pt);
                x.inverse = m4_mul(pi, x.inverse);
                x.inverse_transpose = m4_transpose(x.inverse);
            }
        }
        set_xform(sh, x);
        const MeshStore& bm = sc.mesh_store[sc.mesh_shapes[base].mesh];
        Box wb = transform_box(bm.bbox, x.transform);
        put3(sh.bbox_min, wb.mn); put3(sh.bbox_max, wb.mx);
        sc.mesh_shapes.push_back(sh);
    }

    // ---- Triangles (parser.cpp:458-512): each becomes a one-face Mesh over the shared vertex_data ----
    for (auto el : objects->children_named("Triangle")) {
        sc.mesh_store.emplace_back();
        MeshStore& ms = sc.mesh_store.back();
        ms.vertices = sc.vertex_data; ms.uvs = sc.tex_coords;
        dt_shape sh = blank_shape();
        sh.kind = DT_SHAPE_MESH;
        sh.id = 0;                                                            // Shape::id is never set for triangles
        Xform x; x.transform = m4_identity(); x.inverse = m4_identity(); x.inverse_transpose = m4_identity();
        if (auto c = el->child("Transformations")) { if (!compute_transform(sc, x, c->text)) return false; }
        set_xform(sh, x);
        if (auto c = el->child("Textures")) setup_textures(sc, sh, c->text);
        if (auto c = el->child("Material")) { auto v = parse_ints(c->text); sh.material = v.empty() ? 0 : v[0]; }
        else { g_err = "Triangle lacks <Material>"; return false; }
        auto ids = el->child("Indices") ? parse_ints(el->child("Indices")->text) : std::vector<int>();
        if (ids.size() < 3) { g_err = "Triangle lacks <Indices>"; return false; }
        int nverts = (int)(ms.vertices->size() / 3);
        for (int k = 0; k < 3; k++) if (ids[k] < 1 || ids[k] > nverts) { g_err = "triangle vertex id out of range"; return false; }
        add_face(ms, ids[0], ids[1], ids[2], nullptr);
        ms.bbox = ms.fboxes[0];
        if (!bvh_build(ms)) return false;
        sh.mesh = (int)sc.mesh_store.size() - 1;
        sc.mesh_shapes.push_back(sh);
    }

    // ---- Spheres (parser.cpp:514-574) ----
    for (auto el : objects->children_named("Sphere")) {
        dt_shape sh = blank_shape();
        sh.kind = DT_SHAPE_SPHERE;
        sh.id = 0;
        Xform x; x.transform = m4_identity(); x.inverse = m4_identity(); x.inverse_transpose = m4_identity();
        if (auto c = el->child("Transformations")) { if (!compute_transform(sc, x, c->text)) return false; }
        set_xform(sh, x);
        if (auto c = el->child("Textures")) setup_textures(sc, sh, c->text);
        if (auto c = el->child("Material")) { auto v = parse_ints(c->text); sh.material = v.empty() ? 0 : v[0]; }
        int cid = 0;
        if (auto c = el->child("Center")) { auto v = parse_ints(c->text); if (!v.empty()) cid = v[0]; }
        int nverts = (int)(sc.vertex_data->size() / 3);
        if (cid < 1 || cid > nverts) { g_err = "sphere centre vertex id out of range"; return false; }
        sh.center[0] = (*sc.vertex_data)[(size_t)(cid - 1) * 3]; sh.center[1] = (*sc.vertex_data)[(size_t)(cid - 1) * 3 + 1]; sh.center[2] = (*sc.vertex_data)[(size_t)(cid - 1) * 3 + 2];
        sh.radius = f_or(el, "Radius", 1.f);
        parse_motion_blur(el, sh);
        sc.sphere_shapes.push_back(sh);
    }
    for (auto& s : sc.mesh_shapes) if (s.material < 1 || s.material > (int)sc.materials.size()) { g_err = "shape material id out of range"; return false; }
    for (auto& s : sc.sphere_shapes) if (s.material < 1 || s.material > (int)sc.materials.size()) { g_err = "sphere material id out of range"; return false; }
    return true;
}

}  // namespace

void dth_scene::finalize() {
    meshes.clear();
    for (auto& ms : mesh_store) {
        dt_mesh m; memset(&m, 0, sizeof m);
        m.vertices = ms.vertices->data(); m.n_vertices = (int)(ms.vertices->size() / 3);
        m.uvs = ms.uvs->data(); m.n_uvs = (int)(ms.uvs->size() / 2);
        m.vertex_offset = ms.vertex_offset; m.texture_offset = ms.texture_offset;
        m.faces = ms.faces.data(); m.n_faces = (int)ms.faces.size();
        m.bvh = ms.bvh.data(); m.n_bvh_nodes = (int)ms.bvh.size();
        put3(m.bbox_min, ms.bbox.mn); put3(m.bbox_max, ms.bbox.mx);
        m.surface_area = ms.surface_area;
        m.vertex_normals = ms.smooth && !ms.vnormals.empty() ? ms.vnormals.data() : nullptr;
        meshes.push_back(m);
    }
    images.clear();
    for (auto& is : image_store) {
        dt_image im; im.width = is.data.width; im.height = is.data.height; im.channels = is.data.channels; im.is_hdr = is.data.is_hdr;
        im.data = is.data.is_hdr ? (const void*)is.data.f32.data() : (const void*)is.data.u8.data();
        images.push_back(im);
    }
    shapes = mesh_shapes;
    shapes.insert(shapes.end(), sphere_shapes.begin(), sphere_shapes.end());
    desc.materials = materials.data(); desc.n_materials = (int)materials.size();
    desc.brdfs = brdfs.data(); desc.n_brdfs = (int)brdfs.size();
    desc.point_lights = point_lights.data(); desc.n_point_lights = (int)point_lights.size();
    desc.area_lights = area_lights.data(); desc.n_area_lights = (int)area_lights.size();
    desc.directional_lights = directional_lights.data(); desc.n_directional_lights = (int)directional_lights.size();
    desc.spot_lights = spot_lights.data(); desc.n_spot_lights = (int)spot_lights.size();
    desc.env_lights = env_lights.data(); desc.n_env_lights = (int)env_lights.size();
    desc.mesh_lights = mesh_lights.data(); desc.n_mesh_lights = (int)mesh_lights.size();
    desc.images = images.data(); desc.n_images = (int)images.size();
    desc.textures = textures.data(); desc.n_textures = (int)textures.size();
    desc.meshes = meshes.data(); desc.n_meshes = (int)meshes.size();
    desc.shapes = shapes.data(); desc.n_shapes = (int)shapes.size();
    desc.n_mesh_shapes = (int)mesh_shapes.size();
}

extern "C" {

void dth_set_bvh_builder(dth_bvh_builder builder, int32_t min_faces) { g_bvh_builder = builder; g_bvh_builder_min_faces = min_faces; }
double dth_last_bvh_build_seconds(void) { return g_bvh_seconds; }

int dth_scene_load_xml(const char* xml_path, dth_scene** out) {
    if (!xml_path || !out) { g_err = "null argument"; return DT_ERR_INVALID; }
    *out = nullptr;
    g_bvh_seconds = 0.0;
    std::ifstream f(xml_path, std::ios::binary);
    if (!f) { g_err = std::string("Error: The xml file cannot be loaded: ") + xml_path; return DT_ERR_INVALID; }
    std::stringstream ss; ss << f.rdbuf();
    std::string err;
    auto root = dth::xml_parse(ss.str(), err);
    if (!root) { g_err = "XML parse error: " + err; return DT_ERR_INVALID; }
    std::unique_ptr<dth_scene> sc(new dth_scene());
    std::string p(xml_path);
    size_t sl = p.find_last_of('/');
    sc->xml_dir = sl == std::string::npos ? "." : p.substr(0, sl);
    // free build-time face data after the BVH is built? kept: small relative to faces.
    if (!parse_scene(*sc, root.get())) return DT_ERR_INVALID;
    for (auto& ms : sc->mesh_store) { BigVec<V3>().swap(ms.centers); BigVec<Box>().swap(ms.fboxes); }
    sc->finalize();
    *out = sc.release();
    return DT_OK;
}

void dth_scene_free(dth_scene* s) { delete s; }
const dt_scene_desc* dth_scene_desc(const dth_scene* s) { return s ? &s->desc : nullptr; }
int dth_scene_num_cameras(const dth_scene* s) { return s ? (int)s->cameras.size() : 0; }
const dt_camera_desc* dth_scene_camera(const dth_scene* s, int i) { return (s && i >= 0 && i < (int)s->cameras.size()) ? &s->cameras[i].desc : nullptr; }
const char* dth_scene_camera_image_name(const dth_scene* s, int i) { return (s && i >= 0 && i < (int)s->cameras.size()) ? s->cameras[i].image_name.c_str() : nullptr; }

int dth_scene_set_image(dth_scene* s, int index, int width, int height, int channels, int is_hdr, const void* data) {
    if (!s || index < 0 || index >= (int)s->image_store.size() || !data || width <= 0 || height <= 0) { g_err = "bad image argument"; return DT_ERR_INVALID; }
    ImageStore& is = s->image_store[index];
    is.data.width = width; is.data.height = height; is.data.is_hdr = is_hdr != 0;
    if (is_hdr) { is.data.channels = 3; is.data.f32.assign((const float*)data, (const float*)data + (size_t)width * height * 3); }
    else { is.data.channels = channels; is.data.u8.assign((const uint8_t*)data, (const uint8_t*)data + (size_t)width * height * channels); }
    is.loaded = true;
    s->finalize();
    return DT_OK;
}
const char* dth_scene_image_path(const dth_scene* s, int i) { return (s && i >= 0 && i < (int)s->image_store.size()) ? s->image_store[i].path.c_str() : nullptr; }
int dth_scene_image_loaded(const dth_scene* s, int i) { return (s && i >= 0 && i < (int)s->image_store.size()) ? (s->image_store[i].loaded ? 1 : 0) : 0; }

int dth_camera_look_at(const float pos[3], const float gp[3], const float up[3], float near_dist, float fov_y, int w, int h, dt_camera_desc* out) {
    if (!out) return DT_ERR_INVALID;
    memset(out, 0, sizeof *out);
    camera_look_at(*out, v3(pos[0], pos[1], pos[2]), v3(gp[0], gp[1], gp[2]), v3(up[0], up[1], up[2]), near_dist, fov_y, w, h);
    out->samples_per_pixel = 1;
    return DT_OK;
}
int dth_camera_default(const float pos[3], const float gd[3], const float up[3], const float np[4], float near_dist, int w, int h, dt_camera_desc* out) {
    if (!out) return DT_ERR_INVALID;
    memset(out, 0, sizeof *out);
    camera_default(*out, v3(pos[0], pos[1], pos[2]), v3(gd[0], gd[1], gd[2]), v3(up[0], up[1], up[2]), np, near_dist, w, h);
    out->samples_per_pixel = 1;
    return DT_OK;
}

int dth_write_png(const char* path, int w, int h, const uint8_t* rgb) {
    std::string err;
    if (!dth::png_write(path, w, h, rgb, err)) { g_err = err; return DT_ERR_INVALID; }
    return DT_OK;
}

const char* dth_last_error(void) { return g_err.c_str(); }
void dth_internal_set_error(const char* msg) { g_err = msg ? msg : ""; }

}  // extern "C"
