// Host-side flattener: dt_scene_desc -> SoA device buffers (see dt_device.h for the layout).
#pragma once
#include <string>
#include <vector>
#include "dt_device.h"

struct DtWideBvh {
    std::vector<DtNode8> nodes;
    std::vector<uint32_t> prim_order;   // prim_order[k] = source primitive index of the k-th primitive in leaf order
    int max_depth = 0;
};

// Generic binary tree over a contiguous primitive range (the reference's BVH2 or our TLAS build).
struct DtB2Node {
    float mn[3], mx[3];
    int left, right;                    // -1 for leaves
    uint32_t first, count;              // primitive range of the whole subtree (contiguous)
};

// Collapse a binary tree (root = node 0, leaves of <= 3 primitives) into a BVH8 with 8-bit quantised child
// boxes rounded outward.  Returns false (err filled) if the tree is malformed.
bool dt_collapse_bvh8(const std::vector<DtB2Node>& b2, DtWideBvh& out, std::string& err);

struct DtHostScene {
    // everything that gets uploaded, in upload layout
    std::vector<DtNode8> tlas_nodes;
    std::vector<int32_t> tlas_prims;
    std::vector<DtNode8> blas_nodes;
    std::vector<float4> tris;
    std::vector<float4> leaf_boxes;
    std::vector<uint32_t> face_prim;      // canonical face (DtMeshDev::face_base + f) -> index into tris / leaf_boxes
    std::vector<DtShapeDev> shapes;
    std::vector<DtMeshDev> meshes;
    std::vector<DtFaceDev> faces;
    std::vector<float> verts;
    std::vector<float> uvs;
    std::vector<float> vnormals;
    std::vector<DtImageDev> images;
    std::vector<uint8_t> image_u8;
    std::vector<float> image_f32;
    int max_stack_need = 0;
    int tlas_depth = 0, blas_depth = 0;   // BVH8 depths (blas_depth: host-flattened meshes only)
    uint64_t n_triangles = 0;
    float world_min[3] = {0.f, 0.f, 0.f}, world_max[3] = {0.f, 0.f, 0.f};    // union of the shapes' world boxes (TLAS root)
    std::vector<int> gpu_meshes;          // meshes left to the GPU flattener (dt_flatten_gpu.cu): faces / verts / DtMeshDev are filled, the BLAS is not
};

// gpu_min_faces: meshes with at least this many faces are left to the GPU flattener (INT_MAX: flatten everything here).
bool dt_flatten_scene(const dt_scene_desc* d, DtHostScene& out, std::string& err, int gpu_min_faces = 0x7FFFFFFF);
