// GPU flattener for large meshes: reference BVH2 (Mesh::bvh, mesh.cpp:23-135) -> BVH8 nodes, leaf-order triangles,
// reference leaf boxes and the canonical-face -> primitive map, written straight into the scene's device arrays.
// Same algorithm as the host flattener (dt_flatten.cu) and, through dt_collapse_core.h, the same per-node arithmetic, so
// both produce identical bytes (tests/test_gpu_parity.py compares checksums of the device arrays).  The collapse runs one
// tree level per step: per-node work in parallel, child / primitive offsets by exclusive scans in BFS order.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <string>
#include <vector>
#include "dt_collapse_core.h"
#include "dt_flatten_gpu.h"

namespace {

#define GCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string("GPU flattener: ") + #call + ": " + cudaGetErrorString(e_); return false; } } while (0)

enum { ERR_CHILD_ORDER = 1, ERR_LEAF_RANGE = 2, ERR_CAPACITY = 4, ERR_EXPONENT = 8, ERR_LEAF_COUNT = 16 };

__device__ __forceinline__ float min_first(float a, float b) { return b < a ? b : a; }     // std::min / std::max argument order (sign of zero)
__device__ __forceinline__ float max_first(float a, float b) { return a < b ? b : a; }

__device__ __forceinline__ void tri_box(const DtFaceDev* faces, const float* V, uint32_t f, float* mn, float* mx) {
    const DtFaceDev fd = faces[f];
    for (int a = 0; a < 3; a++) {
        const float x0 = V[(size_t)fd.v0 * 3 + a], x1 = V[(size_t)fd.v1 * 3 + a], x2 = V[(size_t)fd.v2 * 3 + a];
        mn[a] = min_first(x0, min_first(x1, x2)); mx[a] = max_first(x0, max_first(x1, x2));
    }
}

// dt_bvh2_node -> DtB2Node (inner nodes get count = 2: only "more than one primitive" matters to the collapse), the
// BVH2 leaf of every face, and the structural checks of the host flattener.
__global__ void k_b2_init(const dt_bvh2_node* in, int n_nodes, int n_faces, DtB2Node* b2, int* leaf_of, unsigned int* faces_in_leaves, int* error) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const dt_bvh2_node n = in[i];
    DtB2Node b;
    for (int a = 0; a < 3; a++) { b.mn[a] = n.bmin[a]; b.mx[a] = n.bmax[a]; }
    if (n.left >= 0 && n.right >= 0) {
        if (n.left <= i || n.right <= i || n.left >= n_nodes || n.right >= n_nodes) atomicOr(error, ERR_CHILD_ORDER);
        b.left = n.left; b.right = n.right; b.first = n.first_face; b.count = 2u;
    } else {
        b.left = b.right = -1; b.first = n.first_face; b.count = n.face_count;
        if (n.face_count == 0u || (unsigned long long)n.first_face + n.face_count > (unsigned long long)n_faces) atomicOr(error, ERR_LEAF_RANGE);
        else {
            for (uint32_t f = n.first_face; f < n.first_face + n.face_count; f++) leaf_of[f] = i;
            atomicAdd(faces_in_leaves, n.face_count);
        }
    }
    b2[i] = b;
}

// every face must lie in some leaf; together with "the leaf sizes sum to n_faces" (faces_in_leaves) this is "exactly one leaf
// per face", i.e. the leaf ranges tile [0, n_faces) without gaps or overlaps (the host flattener's contiguity / root-coverage checks)
__global__ void k_check_cover(const int* leaf_of, int n_faces, int* error) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n_faces && leaf_of[f] < 0) atomicOr(error, ERR_LEAF_RANGE);
}

// split_big_leaves (dt_flatten.cu): a leaf of k > 1 faces (all centroids on one side of the reference's split plane,
// mesh.cpp:104-106) becomes a balanced subtree of single-face leaves, halving the range.
__global__ void k_split_big_leaves(DtB2Node* b2, int n_orig, uint32_t capacity, unsigned int* n_b2, const DtFaceDev* faces, const float* V, int* error) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_orig) return;
    if (b2[i].left >= 0 || b2[i].count <= 1u) return;
    int stack[40]; int sp = 0;
    stack[sp++] = i;
    while (sp > 0) {
        const int x = stack[--sp];
        const uint32_t first = b2[x].first, count = b2[x].count;
        if (count <= 1u) continue;
        const uint32_t idx = atomicAdd(n_b2, 2u);
        if (idx + 2u > capacity) { atomicOr(error, ERR_CAPACITY); return; }
        const uint32_t lc = count / 2u;
        for (int c = 0; c < 2; c++) {
            DtB2Node nd; nd.left = nd.right = -1;
            nd.first = c == 0 ? first : first + lc; nd.count = c == 0 ? lc : count - lc;
            for (int a = 0; a < 3; a++) { nd.mn[a] = FLT_MAX; nd.mx[a] = -FLT_MAX; }
            for (uint32_t k = 0; k < nd.count; k++) {
                float mn[3], mx[3];
                tri_box(faces, V, nd.first + k, mn, mx);
                for (int a = 0; a < 3; a++) { nd.mn[a] = min_first(nd.mn[a], mn[a]); nd.mx[a] = max_first(nd.mx[a], mx[a]); }
            }
            b2[idx + c] = nd;
        }
        b2[x].left = (int)idx; b2[x].right = (int)idx + 1;
        if (sp + 2 <= 40) { stack[sp++] = (int)idx + 1; stack[sp++] = (int)idx; }
    }
}

__global__ void k_collapse_level(const DtB2Node* b2, const int* items, int n_items, DtNode8* nodes, int* child_ids, uint32_t* n_int, uint32_t* n_leaf, int* error) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_items) return;
    if (i == n_items) { n_int[i] = 0u; n_leaf[i] = 0u; return; }                   // so that the exclusive scans end with the totals
    DtNode8 node; int cis[8];
    const int rc = dt_collapse_node(b2, items[i], node, cis);
    if (rc == DT_COLLAPSE_ERR_EXPONENT) atomicOr(error, ERR_EXPONENT);
    if (rc == DT_COLLAPSE_ERR_LEAF) atomicOr(error, ERR_LEAF_COUNT);
    nodes[i] = node;
    for (int s = 0; s < 8; s++) child_ids[(size_t)i * 8 + s] = cis[s];
    n_int[i] = (uint32_t)__popc(node.imask); n_leaf[i] = (uint32_t)__popc(node.lmask);
}

// child_base = nodes allocated when the host's BFS reaches this node = end of this level + inner children of the nodes
// before it in the level; prim_base likewise; next level's work items and the primitive order in slot order.
__global__ void k_collapse_emit(const DtB2Node* b2, int n_items, DtNode8* nodes, const int* child_ids, const uint32_t* s_int, const uint32_t* s_leaf,
                                uint32_t next_level_begin, uint32_t prims_before, uint32_t node_off, uint32_t prim_off, int* items_next, uint32_t* prim_order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    DtNode8& node = nodes[i];
    uint32_t ci = s_int[i], pi = prims_before + s_leaf[i];
    node.child_base = node_off + next_level_begin + ci;
    node.prim_base = prim_off + pi;
    for (int s = 0; s < 8; s++) {
        const int c = child_ids[(size_t)i * 8 + s];
        if (c < 0) continue;
        if (node.lmask & (1u << s)) prim_order[pi++] = b2[c].first;
        else items_next[ci++] = c;
    }
}

// triangles in leaf order (v0, v0 - v1, v0 - v2: the float differences of mesh.cpp:208-210, canonical face id), the box of
// the reference BVH2 leaf that holds each of them, and the face -> primitive map
__global__ void k_emit_prims(const uint32_t* prim_order, uint32_t n, const DtFaceDev* faces, const float* V, const dt_bvh2_node* bvh, const int* leaf_of,
                             float4* tris, float4* leaf_boxes, uint32_t* face_prim, uint32_t prim_off) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t f = prim_order[k];
    const DtFaceDev fd = faces[f];
    const float* a = V + (size_t)fd.v0 * 3; const float* b = V + (size_t)fd.v1 * 3; const float* c = V + (size_t)fd.v2 * 3;
    const float e1[3] = {__fsub_rn(a[0], b[0]), __fsub_rn(a[1], b[1]), __fsub_rn(a[2], b[2])};
    const float e2[3] = {__fsub_rn(a[0], c[0]), __fsub_rn(a[1], c[1]), __fsub_rn(a[2], c[2])};
    const size_t p = (size_t)prim_off + k;
    face_prim[f] = (uint32_t)p;
    tris[p * 3] = make_float4(a[0], a[1], a[2], e1[0]);
    tris[p * 3 + 1] = make_float4(e1[1], e1[2], e2[0], e2[1]);
    tris[p * 3 + 2] = make_float4(e2[2], __int_as_float((int)f), 0.f, 0.f);
    const int l = leaf_of[f];
    const dt_bvh2_node lf = bvh[l >= 0 ? l : 0];
    leaf_boxes[p * 2] = make_float4(lf.bmin[0], lf.bmin[1], lf.bmin[2], 0.f);
    leaf_boxes[p * 2 + 1] = make_float4(lf.bmax[0], lf.bmax[1], lf.bmax[2], 0.f);
}

__global__ void k_checksum(const uint32_t* w, size_t n, unsigned long long* out) {
    unsigned long long s0 = 0, s1 = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long v = w[i];
        s0 += v; s1 += v * (unsigned long long)((i % 0xFFFFFFFBull) + 1ull);            // position-weighted: catches reorderings
    }
    atomicAdd(out, s0); atomicAdd(out + 1, s1);
}

struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFree(p); }
    template <class T> cudaError_t get(T** p, size_t n) { void* v = nullptr; cudaError_t e = cudaMalloc(&v, n * sizeof(T) + 16); if (e == cudaSuccess) { ptrs.push_back(v); *p = (T*)v; } return e; }
};

}  // namespace

bool dt_flatten_mesh_gpu(const dt_mesh& m, const DtFaceDev* d_faces, const float* d_verts, DtNode8* d_nodes, uint32_t node_off, uint32_t node_capacity,
                         float4* d_tris, float4* d_leaf_boxes, uint32_t* d_face_prim, uint32_t prim_off, uint32_t* n_nodes_out, int* depth_out, std::string& err) {
    const int n_faces = m.n_faces, n_in = m.n_bvh_nodes;
    if (n_faces <= 0 || n_in <= 0 || n_in > 2 * n_faces - 1) { err = "GPU flattener: mesh has no faces or too many BVH2 nodes"; return false; }
    const uint32_t b2_cap = 2u * (uint32_t)n_faces - 1u;
    Scratch S;
    // a tree built by dt_bvh2_build on this device is still resident: no upload
    dt_bvh2_node* resident = dt_resident_tree_take(m.bvh, (uint32_t)n_in);
    struct ResidentGuard { dt_bvh2_node* p; ~ResidentGuard() { if (p) cudaFree(p); } } resident_guard = {resident};
    dt_bvh2_node* bvh; DtB2Node* b2; int *leaf_of, *error, *items[2], *child_ids; unsigned int *counters; uint32_t *n_int, *n_leaf, *s_int, *s_leaf, *prim_order;
    if (resident) bvh = resident; else GCK(S.get(&bvh, (size_t)n_in));
    GCK(S.get(&b2, (size_t)b2_cap)); GCK(S.get(&leaf_of, (size_t)n_faces)); GCK(S.get(&error, 1)); GCK(S.get(&counters, 2));
    GCK(S.get(&items[0], (size_t)n_faces)); GCK(S.get(&items[1], (size_t)n_faces)); GCK(S.get(&child_ids, (size_t)n_faces * 8));
    GCK(S.get(&n_int, (size_t)n_faces + 1)); GCK(S.get(&n_leaf, (size_t)n_faces + 1)); GCK(S.get(&s_int, (size_t)n_faces + 1)); GCK(S.get(&s_leaf, (size_t)n_faces + 1));
    GCK(S.get(&prim_order, (size_t)n_faces));
    size_t scan_bytes = 0;
    GCK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, n_int, s_int, n_faces + 1));
    uint8_t* scan_tmp; GCK(S.get(&scan_tmp, scan_bytes));
    cudaStream_t st = nullptr;
    const int TB = 128;
    if (!resident) GCK(cudaMemcpyAsync(bvh, m.bvh, (size_t)n_in * sizeof(dt_bvh2_node), cudaMemcpyHostToDevice, st));
    GCK(cudaMemsetAsync(leaf_of, 0xFF, (size_t)n_faces * sizeof(int), st));
    GCK(cudaMemsetAsync(error, 0, sizeof(int), st));
    const unsigned int h_counters0[2] = {0u, (unsigned int)n_in};                   // faces covered by leaves, binary-tree node count
    GCK(cudaMemcpyAsync(counters, h_counters0, sizeof h_counters0, cudaMemcpyHostToDevice, st));
    k_b2_init<<<(n_in + TB - 1) / TB, TB, 0, st>>>(bvh, n_in, n_faces, b2, leaf_of, counters, error);
    k_check_cover<<<(n_faces + TB - 1) / TB, TB, 0, st>>>(leaf_of, n_faces, error);
    int h_error = 0; unsigned int h_counters[2] = {0, 0};
    // the structural checks are read back BEFORE any kernel dereferences a leaf range (k_split_big_leaves reads faces[first + k])
    GCK(cudaMemcpyAsync(&h_error, error, sizeof(int), cudaMemcpyDeviceToHost, st));
    GCK(cudaMemcpyAsync(h_counters, counters, sizeof h_counters, cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    if (h_error & ERR_CHILD_ORDER) { err = "BVH2 child index order violated"; return false; }
    if (h_error & ERR_LEAF_RANGE) { err = "BVH2 leaf range out of bounds or faces not covered by any leaf"; return false; }
    if (h_counters[0] != (unsigned int)n_faces) { err = "BVH2 leaves do not cover every face exactly once"; return false; }
    k_split_big_leaves<<<(n_in + TB - 1) / TB, TB, 0, st>>>(b2, n_in, b2_cap, counters + 1, d_faces, d_verts, error);
    GCK(cudaMemcpyAsync(&h_error, error, sizeof(int), cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    if (h_error & ERR_CAPACITY) { err = "BVH2 leaf ranges overlap (more than 2n-1 nodes after splitting)"; return false; }

    // level-synchronous collapse; nodes of this mesh are written at d_nodes[node_off ...]
    const int root_item = 0;
    GCK(cudaMemcpyAsync(items[0], &root_item, sizeof(int), cudaMemcpyHostToDevice, st));
    uint32_t level_begin = 0, n_items = 1, prims_before = 0;
    int cur = 0, depth = 0;
    while (n_items > 0) {
        depth++;
        if (level_begin + n_items > node_capacity) { err = "GPU flattener: node capacity exceeded"; return false; }
        DtNode8* level_nodes = d_nodes + node_off + level_begin;
        k_collapse_level<<<(n_items + 1 + TB - 1) / TB, TB, 0, st>>>(b2, items[cur], (int)n_items, level_nodes, child_ids, n_int, n_leaf, error);
        GCK(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, n_int, s_int, (int)n_items + 1, st));
        GCK(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, n_leaf, s_leaf, (int)n_items + 1, st));
        uint32_t tot_int = 0, tot_leaf = 0;
        GCK(cudaMemcpyAsync(&tot_int, s_int + n_items, 4, cudaMemcpyDeviceToHost, st));
        GCK(cudaMemcpyAsync(&tot_leaf, s_leaf + n_items, 4, cudaMemcpyDeviceToHost, st));
        GCK(cudaMemcpyAsync(&h_error, error, sizeof(int), cudaMemcpyDeviceToHost, st));
        GCK(cudaStreamSynchronize(st));
        if (h_error & ERR_EXPONENT) { err = "BVH8 quantisation exponent overflow (scene extent beyond 2^100)"; return false; }
        if (h_error & ERR_LEAF_COUNT) { err = "leaf with unsupported primitive count (the binary tree must be split down to single primitives)"; return false; }
        if (prims_before + tot_leaf > (uint32_t)n_faces || tot_int > (uint32_t)n_faces) { err = "GPU flattener: malformed tree (more leaves than faces)"; return false; }
        k_collapse_emit<<<(n_items + TB - 1) / TB, TB, 0, st>>>(b2, (int)n_items, level_nodes, child_ids, s_int, s_leaf, level_begin + n_items, prims_before,
                                                                 node_off, prim_off, items[1 - cur], prim_order);
        level_begin += n_items; prims_before += tot_leaf; n_items = tot_int; cur = 1 - cur;
    }
    if (prims_before != (uint32_t)n_faces) { err = "GPU flattener: the collapsed tree does not hold every face"; return false; }
    k_emit_prims<<<((uint32_t)n_faces + TB - 1) / TB, TB, 0, st>>>(prim_order, (uint32_t)n_faces, d_faces, d_verts, bvh, leaf_of, d_tris, d_leaf_boxes, d_face_prim, prim_off);
    GCK(cudaStreamSynchronize(st));
    GCK(cudaGetLastError());
    *n_nodes_out = level_begin; *depth_out = depth;
    return true;
}

bool dt_device_checksum(const void* p, size_t bytes, uint64_t out[2], std::string& err) {
    unsigned long long* d = nullptr;
    out[0] = out[1] = 0;
    GCK(cudaMalloc(&d, 16));
    GCK(cudaMemset(d, 0, 16));
    if (bytes >= 4) k_checksum<<<592, 256>>>((const uint32_t*)p, bytes / 4, d);
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { err = std::string("checksum: ") + cudaGetErrorString(e); return false; }
    out[0] = h[0]; out[1] = h[1];
    return true;
}
