// GPU flattener entry points (dt_flatten_gpu.cu), used by dt_scene_create for meshes above DT_GPU_FLATTEN_MIN_FACES.
#pragma once
#include <stdint.h>
#include <string>
#include "dt_device.h"

// One mesh: m.bvh (host) -> BVH8 nodes at d_nodes[node_off ..) (child_base / prim_base already global), triangles and
// reference leaf boxes at primitive slots [prim_off, prim_off + n_faces), face -> primitive map.  d_faces / d_verts /
// d_face_prim point at THIS mesh's slice of the scene arrays (DtMeshDev::face_base / vert_base applied).
bool dt_flatten_mesh_gpu(const dt_mesh& m, const DtFaceDev* d_faces, const float* d_verts, DtNode8* d_nodes, uint32_t node_off, uint32_t node_capacity,
                         float4* d_tris, float4* d_leaf_boxes, uint32_t* d_face_prim, uint32_t prim_off, uint32_t* n_nodes_out, int* depth_out, std::string& err);

// (sum of 32-bit words, position-weighted sum) of a device buffer
bool dt_device_checksum(const void* p, size_t bytes, uint64_t out[2], std::string& err);

// dt_build.cu: device copies of the trees dt_bvh2_build returned most recently (see there); `take` transfers ownership (cudaFree)
dt_bvh2_node* dt_resident_tree_take(const dt_bvh2_node* host, uint32_t n);
void dt_resident_trees_clear();
