// dt_bvh2_build: Mesh::ConstructBVH / RecursiveBVHBuild / RecomputeBoundingBox (mesh.cpp:23-156) on the GPU.
//
// The reference builds a binary tree by longest-axis spatial-midpoint splits of the node box and partitions the face array IN
// PLACE with a sequential two-pointer loop (mesh.cpp:92-102).  The resulting face order is the scene's canonical face numbering
// (the tie rule of the traversal kernels and the leaf boxes they confirm hits against depend on it), so a GPU build has to
// reproduce that permutation exactly, not just some tree.  The loop
//     i = first, j = last;  while (i <= j) { if (left(a[i])) i++; else { swap(a[i], a[j]); j--; } }
// has a closed form.  With m = number of "left" faces of a segment of n, L[p] the class of the face at local position p,
// bad-left = "right" faces inside [0, m) (ascending, b_1 .. b_B) and bad-right = "left" faces inside [m, n) (DESCENDING):
//     p <  m:  out[p] = L[p] ? a[p] : (the k-th bad-right, k = rank of p among the bad-lefts)
//     p >= m:  out[p] = (p == n-1 || L[p+1]) ? b_k with k = 1 + #L in (p, n)   (b_{B+1} := a[m])
//                                              : a[p+1]
// (every rejected face lands one position below the face it displaced; a displaced "left" face fills the hole of the bad-left
// that started the chain).  tests/test_cpu_oracle_host.py checks the closed form against the loop exhaustively for small n,
// tests/test_gpu_parity.py checks this builder bit for bit against the host build (dth_scene.cpp, the reference's algorithm).
// The same permutation is applied when a split is REJECTED afterwards (all faces on one side, mesh.cpp:104-106): the
// reference has already swapped by then.
//
// Level-synchronous: one pass per tree level over all n face positions (flag -> exclusive scan -> bad-left / bad-right index
// lists -> gather), child boxes by float atomics on order-preserving integer keys, and at the end the node numbering of the
// reference's recursion (children of the k-th split in depth-first pre-order get indices 2k+1, 2k+2).
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <math.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/dorktracer.h"

void dt_internal_set_error(const std::string& msg);      // dt_api.cu
int dt_internal_ensure_device();
void dt_resident_tree_put(const dt_bvh2_node* host, uint32_t n, dt_bvh2_node* dev);

namespace {

#define BCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { dt_internal_set_error(std::string("dt_bvh2_build: ") + #call + ": " + cudaGetErrorString(e_)); return DT_ERR_CUDA; } } while (0)
#define BFAIL(code, msg) do { dt_internal_set_error(msg); return code; } while (0)

struct BNode {
    uint32_t kmin[3], kmax[3];     // box as order-preserving keys (atomicMin / atomicMax)
    int32_t left, right;           // build numbering (allocation order of the level loop)
    uint32_t first, count;
    uint32_t m, bad;               // faces on the left side; "right" faces inside the left zone
    float split;
    int32_t axis;                  // -1: not partitioned at this level
    int32_t accepted;
    uint32_t n_internal;           // internal nodes in this subtree
    uint32_t rank;                 // pre-order rank among internal nodes
    uint32_t out_index;            // index in the reference's numbering
};

__host__ __device__ __forceinline__ uint32_t f2key(float f) { uint32_t b; memcpy(&b, &f, 4); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__host__ __device__ __forceinline__ float key2f(uint32_t k) { uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k; float f; memcpy(&f, &b, 4); return f; }

// mesh.cpp:60-84: longest axis of the node box (ties resolved by the nested comparisons), split = min + len * 0.5f
__global__ void k_decide(BNode* nodes, uint32_t begin, uint32_t end) {
    const uint32_t v = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= end) return;
    BNode& n = nodes[v];
    n.axis = -1; n.accepted = 0; n.left = n.right = -1;
    if (n.count < 2) return;
    const float mn[3] = {key2f(n.kmin[0]), key2f(n.kmin[1]), key2f(n.kmin[2])};
    const float mx[3] = {key2f(n.kmax[0]), key2f(n.kmax[1]), key2f(n.kmax[2])};
    const float lenX = __fsub_rn(mx[0], mn[0]), lenY = __fsub_rn(mx[1], mn[1]), lenZ = __fsub_rn(mx[2], mn[2]);
    int axis;
    if (lenX > lenY) axis = (lenX > lenZ) ? 0 : 2;
    else axis = (lenY > lenZ) ? 1 : 2;
    const float len = axis == 0 ? lenX : (axis == 1 ? lenY : lenZ);
    n.axis = axis;
    n.split = __fadd_rn(mn[axis], __fmul_rn(len, 0.5f));
}

__global__ void k_flag(const BNode* nodes, const int32_t* seg, const float* cx, const float* cy, const float* cz, uint32_t n, uint8_t* flag) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n) return;
    uint8_t f = 0;
    if (p < n) {
        const int32_t v = seg[p];
        if (v >= 0) {
            const int axis = nodes[v].axis;
            if (axis >= 0) f = ((axis == 0 ? cx[p] : (axis == 1 ? cy[p] : cz[p])) < nodes[v].split) ? 1 : 0;
        }
    }
    flag[p] = f;                     // flag[n] = 0 so that excl[n] is the grand total
}

__global__ void k_count(BNode* nodes, uint32_t begin, uint32_t end, const uint32_t* excl, uint32_t* n_nodes, uint32_t capacity, int* overflow) {
    const uint32_t v = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= end) return;
    BNode& n = nodes[v];
    if (n.axis < 0) return;
    const uint32_t base = excl[n.first];
    const uint32_t m = excl[n.first + n.count] - base;
    n.m = m;
    n.bad = m - (excl[n.first + m] - base);
    if (m == 0 || m == n.count) return;                    // mesh.cpp:104-106: stays a leaf (its faces are permuted all the same)
    const uint32_t idx = atomicAdd(n_nodes, 2u);
    if (idx + 2 > capacity) { *overflow = 1; return; }
    n.accepted = 1; n.left = (int32_t)idx; n.right = (int32_t)idx + 1;
    for (int c = 0; c < 2; c++) {
        BNode& ch = nodes[idx + c];
        for (int a = 0; a < 3; a++) { ch.kmin[a] = f2key(INFINITY); ch.kmax[a] = f2key(-INFINITY); }
        ch.left = ch.right = -1;
        ch.first = c == 0 ? n.first : n.first + m;
        ch.count = c == 0 ? m : n.count - m;
        ch.axis = -1; ch.accepted = 0; ch.m = ch.bad = 0; ch.n_internal = 0; ch.rank = 0; ch.out_index = 0; ch.split = 0.f;
    }
}

// index lists of the misplaced faces: bad-lefts ascending, bad-rights descending (see the header)
__global__ void k_bad_lists(const BNode* nodes, const int32_t* seg, const uint8_t* flag, const uint32_t* excl, uint32_t n, uint32_t* bad_left, uint32_t* bad_right) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int32_t v = seg[p];
    if (v < 0 || nodes[v].axis < 0) return;
    const BNode& nd = nodes[v];
    const uint32_t lp = p - nd.first, cnt_before = excl[p] - excl[nd.first];
    if (lp < nd.m) { if (!flag[p]) bad_left[nd.first + (lp - cnt_before)] = p; }
    else if (flag[p]) bad_right[nd.first + (nd.m - cnt_before) - 1] = p;
}

__global__ void k_permute(BNode* nodes, const int32_t* seg, const uint8_t* flag, const uint32_t* excl, uint32_t n, const uint32_t* bad_left, const uint32_t* bad_right,
                          const uint32_t* id_in, const float* cx_in, const float* cy_in, const float* cz_in,
                          uint32_t* id_out, float* cx_out, float* cy_out, float* cz_out, int32_t* seg_out, const float* fboxes) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < n;
    int32_t v = live ? seg[p] : -1;
    uint32_t src = p;
    int32_t child = -1;
    if (v >= 0 && nodes[v].axis >= 0) {
        const BNode& nd = nodes[v];
        const uint32_t lp = p - nd.first, base = excl[nd.first];
        if (lp < nd.m) src = flag[p] ? p : bad_right[nd.first + (lp - (excl[p] - base))];
        else if (lp == nd.count - 1 || flag[p + 1]) {
            const uint32_t k = 1u + (nd.m - (excl[p + 1] - base));               // 1-based index into the rejected stream b
            src = k <= nd.bad ? bad_left[nd.first + k - 1] : nd.first + nd.m;
        } else src = p + 1;
        if (nd.accepted) child = lp < nd.m ? nd.left : nd.right;
    }
    uint32_t id = 0;
    if (live) {
        id = id_in[src];
        id_out[p] = id; cx_out[p] = cx_in[src]; cy_out[p] = cy_in[src]; cz_out[p] = cz_in[src];
        seg_out[p] = child;                                                      // -1: this face's leaf is final
    }
    // RecomputeBoundingBox (mesh.cpp:137-156) of the two children: min / max over the face boxes, order-independent
    const unsigned mask = __activemask();
    const unsigned has = __ballot_sync(mask, child >= 0);
    if (has == 0u) return;
    uint32_t k[6];
    if (child >= 0) { const float* fb = fboxes + (size_t)id * 6; for (int a = 0; a < 6; a++) k[a] = f2key(fb[a]); }
    const int first_lane = __ffs(has) - 1;
    const int32_t c0 = __shfl_sync(mask, child, first_lane);
    if (__all_sync(mask, child == c0 || child < 0)) {                           // the common case: one child per warp
        for (int a = 0; a < 3; a++) {
            const uint32_t lo = __reduce_min_sync(mask, child >= 0 ? k[a] : 0xFFFFFFFFu);
            const uint32_t hi = __reduce_max_sync(mask, child >= 0 ? k[3 + a] : 0u);
            if ((int)(threadIdx.x & 31) == first_lane) { atomicMin(&nodes[c0].kmin[a], lo); atomicMax(&nodes[c0].kmax[a], hi); }
        }
    } else if (child >= 0) {
        for (int a = 0; a < 3; a++) { atomicMin(&nodes[child].kmin[a], k[a]); atomicMax(&nodes[child].kmax[a], k[3 + a]); }
    }
}

__global__ void k_subtree_counts(BNode* nodes, uint32_t begin, uint32_t end) {
    const uint32_t v = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= end) return;
    BNode& n = nodes[v];
    n.n_internal = n.left >= 0 ? 1u + nodes[n.left].n_internal + nodes[n.right].n_internal : 0u;
}
// mesh.cpp:107-121: both children are allocated before the recursion descends left, then right
__global__ void k_number(BNode* nodes, uint32_t begin, uint32_t end) {
    const uint32_t v = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= end) return;
    const BNode& n = nodes[v];
    if (n.left < 0) return;
    BNode& l = nodes[n.left]; BNode& r = nodes[n.right];
    l.out_index = 2u * n.rank + 1u; r.out_index = 2u * n.rank + 2u;
    l.rank = n.rank + 1u;
    r.rank = n.rank + 1u + l.n_internal;
}
__global__ void k_emit(const BNode* nodes, uint32_t n_nodes, dt_bvh2_node* out) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    const BNode& n = nodes[v];
    dt_bvh2_node o;
    for (int a = 0; a < 3; a++) { o.bmin[a] = key2f(n.kmin[a]); o.bmax[a] = key2f(n.kmax[a]); }
    o.left = n.left >= 0 ? (int32_t)nodes[n.left].out_index : -1;
    o.right = n.right >= 0 ? (int32_t)nodes[n.right].out_index : -1;
    o.first_face = n.first;
    o.face_count = n.left >= 0 ? 0u : n.count;                                  // mesh.cpp:117: an inner node's faceCount is reset
    out[n.out_index] = o;
}
__global__ void k_iota(uint32_t* id, int32_t* seg, const float* xyz, float* cx, float* cy, float* cz, uint32_t n) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) { id[p] = p; seg[p] = 0; cx[p] = xyz[3 * (size_t)p]; cy[p] = xyz[3 * (size_t)p + 1]; cz[p] = xyz[3 * (size_t)p + 2]; }
}

// inputs must be finite (a NaN centre would fall on neither side of a split): checked on the device, after the upload
__global__ void k_check_finite(const float* boxes, size_t n_boxes, const float* centers, size_t n_centers, int* bad) {
    bool ok = true;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_boxes + n_centers; i += (size_t)gridDim.x * blockDim.x) {
        const float v = i < n_boxes ? boxes[i] : centers[i - n_boxes];
        ok &= fabsf(v) <= 3.0e38f;
    }
    if (!ok) atomicOr(bad, 1);
}

struct DevBuf {
    std::vector<void*> ptrs;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~DevBuf() { for (void* p : ptrs) cudaFree(p); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
    template <class T> cudaError_t get(T** p, size_t n) { void* v = nullptr; cudaError_t e = cudaMalloc(&v, n * sizeof(T) + 16); if (e == cudaSuccess) { ptrs.push_back(v); *p = (T*)v; } return e; }
};

}  // namespace

// ---- resident trees: the most recent dt_bvh2_build results stay on the device until dt_scene_create consumes them ----
// Keyed by the HOST array the caller received (pointer + node count) and guarded by a fingerprint of three nodes, so a caller that
// edits or frees-and-reuses the array simply gets the ordinary upload.  At most DT_RESIDENT_MAX trees; the oldest is dropped.
#include <mutex>
namespace {
struct ResidentTree { const dt_bvh2_node* host; uint32_t n; dt_bvh2_node* dev; int device; dt_bvh2_node probe[3]; };
std::mutex g_res_mu;
std::vector<ResidentTree> g_res;
const size_t DT_RESIDENT_MAX = 4;
void res_probe(const dt_bvh2_node* h, uint32_t n, dt_bvh2_node out[3]) { out[0] = h[0]; out[1] = h[n / 2]; out[2] = h[n - 1]; }
}
void dt_resident_tree_put(const dt_bvh2_node* host, uint32_t n, dt_bvh2_node* dev) {
    ResidentTree t; t.host = host; t.n = n; t.dev = dev; t.device = 0;
    cudaGetDevice(&t.device);
    res_probe(host, n, t.probe);
    std::lock_guard<std::mutex> lk(g_res_mu);
    for (size_t k = 0; k < g_res.size(); k++) if (g_res[k].host == host) { cudaFree(g_res[k].dev); g_res.erase(g_res.begin() + (long)k); break; }
    if (g_res.size() >= DT_RESIDENT_MAX) { cudaFree(g_res.front().dev); g_res.erase(g_res.begin()); }
    g_res.push_back(t);
}
// the device copy of `host[0..n)` if dt_bvh2_build produced it on this device and the host array still holds it; the caller owns (cudaFree) the result
dt_bvh2_node* dt_resident_tree_take(const dt_bvh2_node* host, uint32_t n) {
    int device = 0;
    cudaGetDevice(&device);
    std::lock_guard<std::mutex> lk(g_res_mu);
    for (size_t k = 0; k < g_res.size(); k++) {
        ResidentTree& t = g_res[k];
        if (t.host != host || t.n != n || t.device != device) continue;
        dt_bvh2_node now[3];
        res_probe(host, n, now);
        dt_bvh2_node* dev = t.dev;
        const bool same = memcmp(now, t.probe, sizeof now) == 0;
        g_res.erase(g_res.begin() + (long)k);
        if (same) return dev;
        cudaFree(dev);
        return nullptr;
    }
    return nullptr;
}
void dt_resident_trees_clear() {
    std::lock_guard<std::mutex> lk(g_res_mu);
    for (auto& t : g_res) cudaFree(t.dev);
    g_res.clear();
}

extern "C" int dt_bvh2_build(int32_t n_faces, const float* centers, const float* face_boxes, const float* root_min, const float* root_max,
                             uint32_t* face_order, dt_bvh2_node* nodes_out, uint32_t node_capacity, uint32_t* n_nodes_out, float* ms_device) {
    if (n_faces <= 0 || !centers || !face_boxes || !root_min || !root_max || !face_order || !nodes_out || !n_nodes_out) BFAIL(DT_ERR_INVALID, "dt_bvh2_build: null or empty argument");
    const uint32_t n = (uint32_t)n_faces;
    const uint32_t capacity = 2u * n - 1u;
    if (node_capacity < capacity) BFAIL(DT_ERR_INVALID, "dt_bvh2_build: node array must hold 2 * n_faces - 1 nodes (mesh.cpp:29)");
    int rc = dt_internal_ensure_device();
    if (rc) return rc;
    DevBuf B;
    BNode* nodes; uint32_t *id[2], *excl, *bad_left, *bad_right, *d_n_nodes; float *c[2][3], *fboxes, *aos; int32_t* seg[2]; uint8_t* flag; int* overflow; dt_bvh2_node* d_out;
    BCK(B.get(&nodes, capacity)); BCK(B.get(&excl, n + 1)); BCK(B.get(&bad_left, n)); BCK(B.get(&bad_right, n)); BCK(B.get(&d_n_nodes, 1)); BCK(B.get(&overflow, 1));
    BCK(B.get(&fboxes, (size_t)n * 6)); BCK(B.get(&aos, (size_t)n * 3)); BCK(B.get(&flag, n + 1)); BCK(B.get(&d_out, capacity));
    for (int k = 0; k < 2; k++) { BCK(B.get(&id[k], n)); BCK(B.get(&seg[k], n)); for (int a = 0; a < 3; a++) BCK(B.get(&c[k][a], n)); }
    size_t scan_bytes = 0;
    BCK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flag, excl, (int)(n + 1)));
    void* scan_tmp; BCK(B.get((uint8_t**)&scan_tmp, scan_bytes));
    cudaStream_t st = nullptr;
    BCK(cudaEventCreate(&B.e0)); BCK(cudaEventCreate(&B.e1));
    const cudaEvent_t e0 = B.e0, e1 = B.e1;
    BCK(cudaMemcpy(fboxes, face_boxes, (size_t)n * 24, cudaMemcpyHostToDevice));
    BCK(cudaMemcpy(aos, centers, (size_t)n * 12, cudaMemcpyHostToDevice));
    BCK(cudaEventRecord(e0, st));
    const int TB = 256;
    {
        int h_bad = 0;
        BCK(cudaMemsetAsync(overflow, 0, 4, st));
        k_check_finite<<<1184, TB, 0, st>>>(fboxes, (size_t)n * 6, aos, (size_t)n * 3, overflow);
        BCK(cudaMemcpyAsync(&h_bad, overflow, 4, cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        if (h_bad) BFAIL(DT_ERR_INVALID, "dt_bvh2_build: non-finite face box or face centre");
    }
    const uint32_t gp = (n + TB) / TB;                                                          // covers n + 1 positions
    k_iota<<<gp, TB, 0, st>>>(id[0], seg[0], aos, c[0][0], c[0][1], c[0][2], n);
    BNode root; memset(&root, 0, sizeof root);
    for (int a = 0; a < 3; a++) { root.kmin[a] = f2key(root_min[a]); root.kmax[a] = f2key(root_max[a]); }      // Mesh::bbox, not recomputed (mesh.cpp:31-35)
    root.left = root.right = -1; root.first = 0; root.count = n; root.axis = -1;
    BCK(cudaMemcpyAsync(nodes, &root, sizeof root, cudaMemcpyHostToDevice, st));
    uint32_t h_n_nodes = 1; int h_overflow = 0;
    BCK(cudaMemcpyAsync(d_n_nodes, &h_n_nodes, 4, cudaMemcpyHostToDevice, st));
    BCK(cudaMemsetAsync(overflow, 0, 4, st));
    std::vector<uint32_t> level_begin; level_begin.push_back(0);
    uint32_t begin = 0, end = 1;
    int cur = 0;
    while (begin < end) {
        const uint32_t gn = (end - begin + TB - 1) / TB;
        k_decide<<<gn, TB, 0, st>>>(nodes, begin, end);
        k_flag<<<gp, TB, 0, st>>>(nodes, seg[cur], c[cur][0], c[cur][1], c[cur][2], n, flag);
        BCK(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, flag, excl, (int)(n + 1), st));
        k_count<<<gn, TB, 0, st>>>(nodes, begin, end, excl, d_n_nodes, capacity, overflow);
        k_bad_lists<<<gp, TB, 0, st>>>(nodes, seg[cur], flag, excl, n, bad_left, bad_right);
        k_permute<<<gp, TB, 0, st>>>(nodes, seg[cur], flag, excl, n, bad_left, bad_right, id[cur], c[cur][0], c[cur][1], c[cur][2],
                                     id[1 - cur], c[1 - cur][0], c[1 - cur][1], c[1 - cur][2], seg[1 - cur], fboxes);
        cur = 1 - cur;
        BCK(cudaMemcpyAsync(&h_n_nodes, d_n_nodes, 4, cudaMemcpyDeviceToHost, st));
        BCK(cudaMemcpyAsync(&h_overflow, overflow, 4, cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        if (h_overflow) BFAIL(DT_ERR_OVERFLOW, "dt_bvh2_build: node array overflow");
        begin = end; end = h_n_nodes;
        if (begin < end) level_begin.push_back(begin);
    }
    level_begin.push_back(h_n_nodes);
    const int n_levels = (int)level_begin.size() - 1;
    for (int l = n_levels - 1; l >= 0; l--) k_subtree_counts<<<(level_begin[l + 1] - level_begin[l] + TB - 1) / TB, TB, 0, st>>>(nodes, level_begin[l], level_begin[l + 1]);
    for (int l = 0; l < n_levels; l++) k_number<<<(level_begin[l + 1] - level_begin[l] + TB - 1) / TB, TB, 0, st>>>(nodes, level_begin[l], level_begin[l + 1]);
    k_emit<<<(h_n_nodes + TB - 1) / TB, TB, 0, st>>>(nodes, h_n_nodes, d_out);
    BCK(cudaEventRecord(e1, st));
    BCK(cudaMemcpyAsync(nodes_out, d_out, (size_t)h_n_nodes * sizeof(dt_bvh2_node), cudaMemcpyDeviceToHost, st));
    BCK(cudaMemcpyAsync(face_order, id[cur], (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    BCK(cudaStreamSynchronize(st));
    BCK(cudaGetLastError());
    if (ms_device) { float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); *ms_device = ms; }
    *n_nodes_out = h_n_nodes;
    // hand the device copy of the tree to the flattener (dt_resident_tree_take): dt_scene_create would otherwise upload the same
    // ~800 MB (config 5) that were just downloaded
    for (size_t k = 0; k < B.ptrs.size(); k++) if (B.ptrs[k] == (void*)d_out) { B.ptrs.erase(B.ptrs.begin() + (long)k); break; }
    dt_resident_tree_put(nodes_out, h_n_nodes, d_out);
    return DT_OK;
}
