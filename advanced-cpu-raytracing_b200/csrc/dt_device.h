// Device-side scene layout (HBM) shared by the flattener (host code) and the kernels.
//
// Geometry is two-level:
//   TLAS  BVH8 (80-byte compressed nodes) over the scene's shapes in world space.  The reference has no
//         TLAS: Raytracer::IntersectObjects (raytracer.cpp:625-643) scans every shape linearly.  A TLAS leaf
//         only nominates candidate shapes; each candidate then runs the reference's exact per-shape tests
//         (double-precision ray transform, float slab tests) so results do not change.
//   BLAS  one BVH8 per Mesh, collapsed from the reference's own BVH2 (Mesh::bvh, mesh.cpp:23-135) so leaves
//         are contiguous canonical face ranges.  Child boxes are quantised to 8 bits rounded OUTWARD: the
//         candidate set is a superset of what BVH::IntersectBVH (bvh.cpp:5-31) would visit.
// Triangles are stored pre-gathered in BVH8 leaf order: 48 bytes = v0, v0-v1, v0-v2 (the exact float
// differences Mesh::IntersectFace forms, mesh.cpp:208-210) + canonical face index.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/dorktracer.h"

#define DT_STACK_SIZE 48          // traversal stack entries (uint2) per ray; build verifies the trees fit

// 80-byte compressed wide-BVH node, five 16-byte words (after Ylitie, Karras, Laine 2017; every leaf child is one primitive,
// so the per-slot meta bytes of that layout shrink to one leaf mask and the node test returns a plain 8-bit slot mask).
struct DtNode8 {
    float px, py, pz;             // quantisation origin
    uint8_t ex, ey, ez, imask;    // per-axis exponent (biased, as float exponent bits), internal-child mask
    uint32_t child_base;          // index of first internal child (children are stored compactly)
    uint32_t prim_base;           // index of first primitive referenced by this node's leaf children
    uint8_t lmask;                // leaf-child mask: slot s holds exactly ONE primitive, prim_base + popc(lmask & ((1 << s) - 1))
    uint8_t pad[7];
    uint8_t qlox[8], qloy[8], qloz[8], qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(DtNode8) == 80, "DtNode8 must be 80 bytes");

struct DtShapeDev {
    int32_t kind;                 // DT_SHAPE_*
    int32_t id;                   // Shape::id
    int32_t mesh;                 // geometry mesh index (own for MESH, base mesh's for INSTANCE), -1 sphere
    int32_t owner;                // shape index whose Mesh::IntersectFace semantics apply (self for MESH, base_shape for INSTANCE)
    int32_t material;             // 1-based
    int32_t tex_diffuse, tex_specular, tex_normal, tex_bump, tex_replace_all;
    int32_t has_motion_blur;
    int32_t skip_shadow;          // mesh-list shape with Emissive material: skipped by CastShadowRay (raytracer.cpp:590-593)
    float motion_blur[3];
    float radius;
    float center[3];
    int32_t inv_is_identity;      // inverseTransform == I exactly: the double transform is a no-op up to -0 -> +0
    float bbox_min[3], bbox_max[3];   // MESH: Mesh::bbox (local); INSTANCE: InstancedMesh::bbox (world)
    float pad1[2];
    double inv[12];               // rows 0..2 of inverseTransform
    double invT[12];              // rows 0..2 of inverseTransposeTransform
    double fwd[12];               // rows 0..2 of transform
};

struct DtMeshDev {
    uint32_t node_root;           // index of this mesh's BVH8 root in blas_nodes
    uint32_t face_base;           // index of canonical face 0 in faces[]
    uint32_t vert_base;           // index of vertex 0 in verts[] (float3 units)
    uint32_t uv_base;             // index of uv 0 in uvs[] (float2 units)
    int32_t n_faces, n_uvs;
    int32_t vertex_offset, texture_offset;
    float bbox_min[3], bbox_max[3];
    double surface_area;
    int32_t normal_base;          // index of vertex normal 0 in vnormals[] (float3 units), -1: the mesh has none (flat shading only)
    int32_t pad_;
};

struct DtFaceDev {                // canonical (post-build) face record, read only when shading the final hit
    int32_t v0, v1, v2;           // 0-based indices into the mesh's vertex array (offset already applied)
    float nx, ny, nz;
    float light_weight;           // (float)(face.area / mesh.surfaceArea): MeshLight::getSample's selectionWeight (meshLight.h:31)
    int32_t pad;
};

struct DtImageDev { int32_t width, height, channels, is_hdr; uint64_t offset; uint64_t count; };  // offset into image_u8 / image_f32

struct DtSceneDev {
    // acceleration
    const uint4* tlas_nodes;      // DtNode8 as 5 x uint4
    const int32_t* tlas_prims;    // shape indices in TLAS leaf order
    const uint4* blas_nodes;
    const float4* tris;           // 3 x float4 per triangle, BVH8 leaf order
    const float4* leaf_boxes;     // 2 x float4 per triangle (same order): bbox of the reference BVH2 leaf that holds it
    const uint32_t* face_prim;    // canonical face -> index into tris / leaf_boxes (rare path: dt_best_survives)
    // shading data
    const DtShapeDev* shapes;
    const DtMeshDev* meshes;
    const DtFaceDev* faces;
    const float* verts;           // xyz
    const float* uvs;             // uv
    const float* vnormals;        // xyz per vertex of the shadingMode="smooth" meshes (DT_FLAG_SMOOTH_SHADING)
    const dt_material* materials;
    const dt_brdf* brdfs;
    const dt_point_light* point_lights;
    const dt_area_light* area_lights;
    const dt_directional_light* directional_lights;
    const dt_spot_light* spot_lights;
    const dt_env_light* env_lights;
    const dt_mesh_light* mesh_lights;
    const dt_texture* textures;
    const DtImageDev* images;
    const uint8_t* image_u8;
    const float* image_f32;
    int32_t n_shapes, n_mesh_shapes, n_materials;
    int32_t n_point_lights, n_area_lights, n_directional_lights, n_spot_lights, n_env_lights, n_mesh_lights;
    int32_t bg_texture, max_recursion_depth;
    uint32_t one_bits;            // 0x3F800000, see dt_byte_m (dt_traverse.cuh)
    int32_t tlas_direct;          // > 0: the TLAS is one leaf-only node over this many (<= 4) shapes -> rays start with the shape list itself
    int32_t background_color[3];
    float shadow_ray_epsilon;
    float ambient_light[3];
    float sort_min[3], sort_scale[3];   // world bounds of the scene for the hit-cell sort: cell coordinate = (p - sort_min) * sort_scale in [0, 1)
};

// ---- wavefront queues (SoA in HBM) ----
// A closest-hit ray: 32 B in (o.xyz + motion-blur time, d.xyz + tmax), 32 B hit record out.
struct DtRayQueue {
    float4* o_time;               // origin.xyz, motionBlurTime
    float4* d_tmax;               // dir.xyz, tmax (INFINITY for closest-hit rays)
    float4* hit0;                 // t, beta, gamma, shape (int bits; -1 miss)
    int32_t* hit_face;            // canonical face index / -1
    // path state (not touched by traversal)
    uint32_t* pixel;              // pixel index | flags in top bits
    float4* weight_n;             // W.rgb (radiance weight of this tree node), refractiveIndexOfCurrentMedium
    float4* thr_beer;             // throughput.rgb (Russian roulette), beer threshold (0 = no Beer attenuation)
    int4* misc;                   // x: recursion depth left, y: parent material (1-based) for Beer, z: rng key, w: flags
    uint32_t* sort_key;           // material key for the sort stage
};

// A shadow ray: 32 B in + pixel + rgb contribution added iff unoccluded.
struct DtShadowQueue {
    float4* o_time;
    float4* d_tmax;               // tmax = distance to the light (INFINITY: directional)
    float4* contrib_pix;          // rgb contribution, pixel index (int bits)
    int2* defer;                  // x: GI-child slot in the next wave (-1 none), y: mesh light id (deferred NEE)
};

#define DT_PIX_MASK 0x0FFFFFFFu
#define DT_FLAG_PRIMARY 1         // misc.w bit0: eye position is the camera position (raytracer.cpp:47)
#define DT_FLAG_ENV_ON_MISS 2     // misc.w bit1: a miss reads the environment map (mirror / dielectric children)
#define DT_FLAG_REFRACT_ENV 4     // misc.w bit2: (unused marker for refracted child; env dir stored separately)
