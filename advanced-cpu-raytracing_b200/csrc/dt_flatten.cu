// Host-side flattener (no device code): reference-shaped scene description -> wide-BVH SoA buffers.
//   * BLAS: collapse of the reference's own BVH2 (Mesh::ConstructBVH, mesh.cpp:23-135) into BVH8
//   * TLAS: median-split tree over world-space shape boxes (the reference has none, raytracer.cpp:625-643)
//   * triangles pre-gathered in leaf order as (v0, v0-v1, v0-v2, canonical face id)
#include "dt_flatten.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>
#include <numeric>

#include "dt_collapse_core.h"

// BFS collapse; the per-node work (child selection, slot assignment, quantisation) is dt_collapse_node (dt_collapse_core.h),
// shared with the GPU flattener (dt_flatten_gpu.cu) so that both produce the same bytes.
bool dt_collapse_bvh8(const std::vector<DtB2Node>& b2, DtWideBvh& out, std::string& err) {
    out.nodes.clear(); out.prim_order.clear(); out.max_depth = 0;
    if (b2.empty()) { err = "empty tree"; return false; }
    struct Work { int b2node; uint32_t out_index; int depth; };
    std::vector<Work> queue;
    out.nodes.emplace_back();
    queue.push_back({0, 0u, 1});
    size_t qh = 0;
    while (qh < queue.size()) {
        Work w = queue[qh++];
        out.max_depth = std::max(out.max_depth, w.depth);
        DtNode8 node; int child_in_slot[8];
        const int rc = dt_collapse_node(b2.data(), w.b2node, node, child_in_slot);
        if (rc == DT_COLLAPSE_ERR_EXPONENT) { err = "BVH8 quantisation exponent overflow (scene extent beyond 2^100)"; return false; }
        if (rc == DT_COLLAPSE_ERR_LEAF) { err = "leaf with unsupported primitive count (the binary tree must be split down to single primitives)"; return false; }
        // children: internal ones get consecutive node indices in slot order; leaves get consecutive primitives
        node.child_base = (uint32_t)out.nodes.size();
        node.prim_base = (uint32_t)out.prim_order.size();
        for (int s = 0; s < 8; s++) {
            const int c = child_in_slot[s];
            if (c < 0) continue;
            if (node.lmask & (1u << s)) out.prim_order.push_back(b2[c].first);       // slot order == primitive order within the node
            else { queue.push_back({c, (uint32_t)out.nodes.size(), w.depth + 1}); out.nodes.emplace_back(); }
        }
        out.nodes[w.out_index] = node;
    }
    return true;
}

namespace {

// Make sure every leaf of a binary tree holds ONE primitive by splitting larger leaf ranges in halves
// (the reference build leaves a multi-face leaf when a midpoint split puts every centroid on one side,
// mesh.cpp:104-106).  prim_box(i, mn, mx) yields the box of primitive i.
template <class BoxFn>
void split_big_leaves(std::vector<DtB2Node>& b2, BoxFn prim_box) {
    for (size_t i = 0; i < b2.size(); i++) {
        if (b2[i].left >= 0 || b2[i].count <= 1) continue;
        uint32_t first = b2[i].first, count = b2[i].count;
        uint32_t lc = count / 2;
        DtB2Node l, r;
        l.left = l.right = r.left = r.right = -1;
        l.first = first; l.count = lc; r.first = first + lc; r.count = count - lc;
        for (DtB2Node* n : {&l, &r}) {
            for (int a = 0; a < 3; a++) { n->mn[a] = FLT_MAX; n->mx[a] = -FLT_MAX; }
            for (uint32_t k = 0; k < n->count; k++) {
                float mn[3], mx[3];
                prim_box(n->first + k, mn, mx);
                for (int a = 0; a < 3; a++) { n->mn[a] = std::min(n->mn[a], mn[a]); n->mx[a] = std::max(n->mx[a], mx[a]); }
            }
        }
        b2[i].left = (int)b2.size(); b2.push_back(l);
        b2[i].right = (int)b2.size(); b2.push_back(r);      // appended nodes are visited later by this same loop
    }
}

struct Box3 { float mn[3], mx[3]; };

void box_extend(Box3& b, const float* p) { for (int a = 0; a < 3; a++) { b.mn[a] = std::min(b.mn[a], p[a]); b.mx[a] = std::max(b.mx[a], p[a]); } }

// world-space box of a local box under a row-major double 4x4, with motion-blur sweep and safety margin
Box3 world_box(const float* lmn, const float* lmx, const double* T, const float* mbv, bool has_mb, bool sweep_local) {
    float mn[3] = {lmn[0], lmn[1], lmn[2]}, mx[3] = {lmx[0], lmx[1], lmx[2]};
    if (has_mb && sweep_local) for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], mn[a] - mbv[a]); mx[a] = std::max(mx[a], mx[a] - mbv[a]); }
    Box3 b; for (int a = 0; a < 3; a++) { b.mn[a] = FLT_MAX; b.mx[a] = -FLT_MAX; }
    for (int c = 0; c < 8; c++) {
        double p[3] = {(c & 1) ? mx[0] : mn[0], (c & 2) ? mx[1] : mn[1], (c & 4) ? mx[2] : mn[2]};
        float q[3];
        for (int r = 0; r < 3; r++) q[r] = (float)(T[r * 4 + 0] * p[0] + T[r * 4 + 1] * p[1] + T[r * 4 + 2] * p[2] + T[r * 4 + 3]);
        box_extend(b, q);
    }
    if (has_mb && !sweep_local) for (int a = 0; a < 3; a++) { float lo = b.mn[a], hi = b.mx[a]; b.mn[a] = std::min(lo, lo - mbv[a]); b.mx[a] = std::max(hi, hi - mbv[a]); }
    for (int a = 0; a < 3; a++) {
        float ext = b.mx[a] - b.mn[a];
        float m = 1e-4f * ext + 1e-5f * std::max(std::fabs(b.mn[a]), std::fabs(b.mx[a])) + 1e-7f;
        if (!(m == m)) m = 0.f;
        b.mn[a] -= m; b.mx[a] += m;
        if (!(b.mn[a] == b.mn[a]) || !(b.mx[a] == b.mx[a]) || std::isinf(b.mn[a]) || std::isinf(b.mx[a])) { b.mn[a] = -1e30f; b.mx[a] = 1e30f; }
    }
    return b;
}

void rows012(double* dst, const double* src16) { memcpy(dst, src16, sizeof(double) * 12); }

}  // namespace

bool dt_flatten_scene(const dt_scene_desc* d, DtHostScene& out, std::string& err, int gpu_min_faces) {
    if (!d) { err = "null scene description"; return false; }
    if (d->abi_version != DT_ABI_VERSION) { err = "dt_scene_desc.abi_version mismatch"; return false; }
    if (d->n_shapes <= 0) { err = "scene has no shapes"; return false; }
    if (d->n_mesh_shapes < 0 || d->n_mesh_shapes > d->n_shapes) { err = "bad n_mesh_shapes"; return false; }

    // ---- vertex / uv pools, deduplicated by host pointer (inline meshes all share Scene::vertex_data) ----
    std::map<const float*, uint32_t> vert_off, uv_off;
    out.meshes.resize((size_t)d->n_meshes);
    int blas_depth = 0;
    for (int mi = 0; mi < d->n_meshes; mi++) {
        const dt_mesh& m = d->meshes[mi];
        if (m.n_faces <= 0 || !m.faces || !m.bvh || m.n_bvh_nodes <= 0 || !m.vertices) { err = "mesh " + std::to_string(mi) + " is empty"; return false; }
        DtMeshDev md; memset(&md, 0, sizeof md);
        auto vit = vert_off.find(m.vertices);
        if (vit == vert_off.end()) {
            uint32_t off = (uint32_t)(out.verts.size() / 3);
            out.verts.insert(out.verts.end(), m.vertices, m.vertices + (size_t)m.n_vertices * 3);
            vit = vert_off.emplace(m.vertices, off).first;
        }
        md.vert_base = vit->second;
        if (m.n_uvs > 0 && m.uvs) {
            auto uit = uv_off.find(m.uvs);
            if (uit == uv_off.end()) {
                uint32_t off = (uint32_t)(out.uvs.size() / 2);
                out.uvs.insert(out.uvs.end(), m.uvs, m.uvs + (size_t)m.n_uvs * 2);
                uit = uv_off.emplace(m.uvs, off).first;
            }
            md.uv_base = uit->second;
        }
        md.normal_base = -1;
        if (m.vertex_normals) {                      // shadingMode="smooth" meshes (DT_FLAG_SMOOTH_SHADING): per mesh, never shared
            md.normal_base = (int32_t)(out.vnormals.size() / 3);
            out.vnormals.insert(out.vnormals.end(), m.vertex_normals, m.vertex_normals + (size_t)m.n_vertices * 3);
        }
        md.n_faces = m.n_faces; md.n_uvs = m.n_uvs;
        md.vertex_offset = m.vertex_offset; md.texture_offset = m.texture_offset;
        memcpy(md.bbox_min, m.bbox_min, 12); memcpy(md.bbox_max, m.bbox_max, 12);
        md.surface_area = m.surface_area;
        md.face_base = (uint32_t)out.faces.size();

        // canonical faces
        out.faces.reserve(out.faces.size() + (size_t)m.n_faces);
        for (int f = 0; f < m.n_faces; f++) {
            const dt_face& fc = m.faces[f];
            int ids[3] = {fc.v0_id - 1 + m.vertex_offset, fc.v1_id - 1 + m.vertex_offset, fc.v2_id - 1 + m.vertex_offset};
            for (int k = 0; k < 3; k++) if (ids[k] < 0 || ids[k] >= m.n_vertices) { err = "face vertex index out of range in mesh " + std::to_string(mi); return false; }
            if (m.n_uvs > 0) {
                int tids[3] = {fc.v0_id - 1 + m.texture_offset, fc.v1_id - 1 + m.texture_offset, fc.v2_id - 1 + m.texture_offset};
                for (int k = 0; k < 3; k++) if (tids[k] < 0 || tids[k] >= m.n_uvs) { err = "face uv index out of range in mesh " + std::to_string(mi) + " (TexCoordData must cover every vertex id, SURVEY Appendix A)"; return false; }
            }
            DtFaceDev fd; fd.v0 = ids[0]; fd.v1 = ids[1]; fd.v2 = ids[2];
            fd.nx = fc.n[0]; fd.ny = fc.n[1]; fd.nz = fc.n[2];
            fd.light_weight = (float)(fc.area / m.surface_area);
            fd.pad = 0;
            out.faces.push_back(fd);
        }

        if (m.n_faces >= gpu_min_faces) {                       // BLAS of this mesh: dt_flatten_mesh_gpu, after the upload of faces / verts
            out.gpu_meshes.push_back(mi);
            out.n_triangles += (uint64_t)m.n_faces;
            out.meshes[(size_t)mi] = md;
            continue;
        }

        // ---- reference BVH2 -> generic binary tree with subtree ranges (children have larger indices than
        // their parent in Mesh::RecursiveBVHBuild's allocation order, so one reverse sweep suffices) ----
        std::vector<DtB2Node> b2((size_t)m.n_bvh_nodes);
        for (int i = m.n_bvh_nodes - 1; i >= 0; i--) {
            const dt_bvh2_node& n = m.bvh[i];
            DtB2Node& b = b2[(size_t)i];
            memcpy(b.mn, n.bmin, 12); memcpy(b.mx, n.bmax, 12);
            if (n.left >= 0 && n.right >= 0) {
                if (n.left <= i || n.right <= i || n.left >= m.n_bvh_nodes || n.right >= m.n_bvh_nodes) { err = "BVH2 child index order violated"; return false; }
                b.left = n.left; b.right = n.right;
                b.first = b2[(size_t)n.left].first;
                b.count = b2[(size_t)n.left].count + b2[(size_t)n.right].count;
                if (b2[(size_t)n.right].first != b.first + b2[(size_t)n.left].count) { err = "BVH2 face ranges are not contiguous"; return false; }
            } else {
                b.left = b.right = -1; b.first = n.first_face; b.count = n.face_count;
                if (b.count == 0 || (uint64_t)b.first + (uint64_t)b.count > (uint64_t)m.n_faces) { err = "BVH2 leaf range out of bounds"; return false; }
            }
        }
        if (b2[0].first != 0 || b2[0].count != (uint32_t)m.n_faces) { err = "BVH2 root does not cover all faces"; return false; }
        // box of the reference BVH2 leaf holding each canonical face (exact accept-time confirmation, dt_traverse.cuh)
        std::vector<int> leaf_of((size_t)m.n_faces, -1);
        for (int i = 0; i < m.n_bvh_nodes; i++) {
            const dt_bvh2_node& n = m.bvh[i];
            if (n.left >= 0 && n.right >= 0) continue;
            for (uint32_t f = n.first_face; f < n.first_face + n.face_count && f < (uint32_t)m.n_faces; f++) leaf_of[f] = i;
        }
        const float* V = m.vertices;
        auto tri_box = [&](uint32_t f, float* mn, float* mx) {
            const DtFaceDev& fd = out.faces[md.face_base + f];
            for (int a = 0; a < 3; a++) {
                float x0 = V[(size_t)fd.v0 * 3 + a], x1 = V[(size_t)fd.v1 * 3 + a], x2 = V[(size_t)fd.v2 * 3 + a];
                mn[a] = std::min(x0, std::min(x1, x2)); mx[a] = std::max(x0, std::max(x1, x2));
            }
        };
        split_big_leaves(b2, tri_box);
        DtWideBvh wide;
        if (!dt_collapse_bvh8(b2, wide, err)) return false;
        blas_depth = std::max(blas_depth, wide.max_depth);
        md.node_root = (uint32_t)out.blas_nodes.size();
        uint32_t node_off = md.node_root, prim_off = (uint32_t)(out.tris.size() / 3);
        for (DtNode8 n : wide.nodes) { n.child_base += node_off; n.prim_base += prim_off; out.blas_nodes.push_back(n); }
        out.face_prim.resize(out.faces.size(), 0);
        for (uint32_t f : wide.prim_order) {
            out.face_prim[md.face_base + f] = (uint32_t)(out.tris.size() / 3);         // canonical face -> leaf-order slot
            const DtFaceDev& fd = out.faces[md.face_base + f];
            const float* a = &V[(size_t)fd.v0 * 3]; const float* b = &V[(size_t)fd.v1 * 3]; const float* c = &V[(size_t)fd.v2 * 3];
            // e1 = v0 - v1, e2 = v0 - v2: the float subtractions of mesh.cpp:208-210 (host compiled without FMA)
            float e1[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
            float e2[3] = {a[0] - c[0], a[1] - c[1], a[2] - c[2]};
            int fi = (int)f; float fbits; memcpy(&fbits, &fi, 4);
            out.tris.push_back(make_float4(a[0], a[1], a[2], e1[0]));
            out.tris.push_back(make_float4(e1[1], e1[2], e2[0], e2[1]));
            out.tris.push_back(make_float4(e2[2], fbits, 0.f, 0.f));
            const dt_bvh2_node& lf = m.bvh[leaf_of[f] >= 0 ? leaf_of[f] : 0];
            out.leaf_boxes.push_back(make_float4(lf.bmin[0], lf.bmin[1], lf.bmin[2], 0.f));
            out.leaf_boxes.push_back(make_float4(lf.bmax[0], lf.bmax[1], lf.bmax[2], 0.f));
        }
        out.n_triangles += (uint64_t)m.n_faces;
        out.meshes[(size_t)mi] = md;
    }

    // ---- shapes ----
    out.shapes.resize((size_t)d->n_shapes);
    std::vector<Box3> wboxes((size_t)d->n_shapes);
    for (int si = 0; si < d->n_shapes; si++) {
        const dt_shape& s = d->shapes[si];
        DtShapeDev sd; memset(&sd, 0, sizeof sd);
        sd.kind = s.kind; sd.id = s.id; sd.material = s.material;
        if (s.material < 1 || s.material > d->n_materials) { err = "shape material id out of range"; return false; }
        sd.tex_diffuse = s.tex_diffuse; sd.tex_specular = s.tex_specular; sd.tex_normal = s.tex_normal; sd.tex_bump = s.tex_bump; sd.tex_replace_all = s.tex_replace_all;
        for (int t : {s.tex_diffuse, s.tex_specular, s.tex_normal, s.tex_bump, s.tex_replace_all}) if (t >= d->n_textures) { err = "shape texture index out of range"; return false; }
        sd.has_motion_blur = s.has_motion_blur; memcpy(sd.motion_blur, s.motion_blur, 12);
        rows012(sd.inv, s.inverse_transform); rows012(sd.invT, s.inverse_transpose_transform); rows012(sd.fwd, s.transform);
        {
            static const double I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
            bool id = true;
            for (int k = 0; k < 12; k++) if (!(s.inverse_transform[k] == I16[k])) id = false;
            sd.inv_is_identity = id ? 1 : 0;
        }
        sd.skip_shadow = (si < d->n_mesh_shapes && d->materials[s.material - 1].type == DT_MAT_EMISSIVE) ? 1 : 0;
        if (s.kind == DT_SHAPE_MESH) {
            if (s.mesh < 0 || s.mesh >= d->n_meshes) { err = "shape mesh index out of range"; return false; }
            sd.mesh = s.mesh; sd.owner = si;
            const dt_mesh& m = d->meshes[s.mesh];
            memcpy(sd.bbox_min, m.bbox_min, 12); memcpy(sd.bbox_max, m.bbox_max, 12);
            wboxes[(size_t)si] = world_box(m.bbox_min, m.bbox_max, s.transform, s.motion_blur, s.has_motion_blur != 0, true);
        } else if (s.kind == DT_SHAPE_INSTANCE) {
            if (s.base_shape < 0 || s.base_shape >= d->n_mesh_shapes || d->shapes[s.base_shape].kind != DT_SHAPE_MESH) { err = "instance base_shape invalid"; return false; }
            sd.mesh = d->shapes[s.base_shape].mesh; sd.owner = s.base_shape;
            memcpy(sd.bbox_min, s.bbox_min, 12); memcpy(sd.bbox_max, s.bbox_max, 12);
            static const double I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
            wboxes[(size_t)si] = world_box(s.bbox_min, s.bbox_max, I, s.motion_blur, s.has_motion_blur != 0, false);
        } else if (s.kind == DT_SHAPE_SPHERE) {
            sd.mesh = -1; sd.owner = si;
            memcpy(sd.center, s.center, 12); sd.radius = s.radius;
            float r = std::fabs(s.radius);
            float lmn[3] = {s.center[0] - r, s.center[1] - r, s.center[2] - r}, lmx[3] = {s.center[0] + r, s.center[1] + r, s.center[2] + r};
            wboxes[(size_t)si] = world_box(lmn, lmx, s.transform, s.motion_blur, s.has_motion_blur != 0, true);
        } else { err = "unknown shape kind"; return false; }
        out.shapes[(size_t)si] = sd;
    }

    // ---- TLAS: top-down median split over shape boxes ----
    {
        std::vector<uint32_t> order((size_t)d->n_shapes);
        std::iota(order.begin(), order.end(), 0u);
        std::vector<DtB2Node> b2;
        struct Task { int node; uint32_t first, count; };
        std::vector<Task> st;
        b2.emplace_back();
        st.push_back({0, 0u, (uint32_t)d->n_shapes});
        while (!st.empty()) {
            Task t = st.back(); st.pop_back();
            DtB2Node nd; nd.left = nd.right = -1; nd.first = t.first; nd.count = t.count;
            float cmn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (int a = 0; a < 3; a++) { nd.mn[a] = FLT_MAX; nd.mx[a] = -FLT_MAX; }
            for (uint32_t k = 0; k < t.count; k++) {
                const Box3& b = wboxes[order[t.first + k]];
                for (int a = 0; a < 3; a++) {
                    nd.mn[a] = std::min(nd.mn[a], b.mn[a]); nd.mx[a] = std::max(nd.mx[a], b.mx[a]);
                    float c = 0.5f * (b.mn[a] + b.mx[a]);
                    cmn[a] = std::min(cmn[a], c); cmx[a] = std::max(cmx[a], c);
                }
            }
            if (t.count > 1) {
                int axis = 0; float best = cmx[0] - cmn[0];
                for (int a = 1; a < 3; a++) if (cmx[a] - cmn[a] > best) { best = cmx[a] - cmn[a]; axis = a; }
                uint32_t half = t.count / 2;
                std::nth_element(order.begin() + t.first, order.begin() + t.first + half, order.begin() + t.first + t.count,
                                 [&](uint32_t x, uint32_t y) { return wboxes[x].mn[axis] + wboxes[x].mx[axis] < wboxes[y].mn[axis] + wboxes[y].mx[axis]; });
                nd.left = (int)b2.size(); b2.emplace_back();
                nd.right = (int)b2.size(); b2.emplace_back();
                st.push_back({nd.left, t.first, half});
                st.push_back({nd.right, t.first + half, t.count - half});
            }
            b2[(size_t)t.node] = nd;
        }
        for (int a = 0; a < 3; a++) { out.world_min[a] = b2[0].mn[a]; out.world_max[a] = b2[0].mx[a]; }
        DtWideBvh wide;
        if (!dt_collapse_bvh8(b2, wide, err)) return false;
        out.tlas_nodes = wide.nodes;
        out.tlas_prims.resize(wide.prim_order.size());
        for (size_t k = 0; k < wide.prim_order.size(); k++) out.tlas_prims[k] = (int32_t)order[wide.prim_order[k]];
        out.tlas_depth = wide.max_depth; out.blas_depth = blas_depth;
        out.max_stack_need = wide.max_depth + blas_depth + 4;
    }
    if (out.max_stack_need > DT_STACK_SIZE) {
        err = "BVH too deep for the traversal stack (" + std::to_string(out.max_stack_need) + " > " + std::to_string(DT_STACK_SIZE) + ")";
        return false;
    }

    // ---- images ----
    out.images.resize((size_t)d->n_images);
    for (int i = 0; i < d->n_images; i++) {
        const dt_image& im = d->images[i];
        DtImageDev id; id.width = im.width; id.height = im.height; id.channels = im.is_hdr ? 3 : im.channels; id.is_hdr = im.is_hdr;
        if (im.width <= 0 || im.height <= 0 || !im.data) { err = "image " + std::to_string(i) + " has no pixels"; return false; }
        size_t n = (size_t)im.width * im.height * (size_t)id.channels;
        id.count = n;
        if (im.is_hdr) { id.offset = out.image_f32.size(); const float* p = (const float*)im.data; out.image_f32.insert(out.image_f32.end(), p, p + n); }
        else { id.offset = out.image_u8.size(); const uint8_t* p = (const uint8_t*)im.data; out.image_u8.insert(out.image_u8.end(), p, p + n); }
        out.images[(size_t)i] = id;
    }
    for (int i = 0; i < d->n_textures; i++) if (d->textures[i].kind == DT_TEX_IMAGE && (d->textures[i].image < 0 || d->textures[i].image >= d->n_images)) { err = "texture image index out of range"; return false; }
    for (int i = 0; i < d->n_env_lights; i++) if (d->env_lights[i].image < 0 || d->env_lights[i].image >= d->n_images || !d->images[d->env_lights[i].image].is_hdr) { err = "environment light needs an HDR image"; return false; }
    for (int i = 0; i < d->n_mesh_lights; i++) if (d->mesh_lights[i].shape < 0 || d->mesh_lights[i].shape >= d->n_mesh_shapes || d->shapes[d->mesh_lights[i].shape].kind != DT_SHAPE_MESH) { err = "mesh light shape invalid"; return false; }
    for (int i = 0; i < d->n_materials; i++) if (d->materials[i].brdf >= d->n_brdfs) { err = "material BRDF index out of range"; return false; }
    if (d->bg_texture >= d->n_textures) { err = "bg_texture out of range"; return false; }
    return true;
}
