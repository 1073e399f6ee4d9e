/*
 * dorktracer.h — C ABI of the B200-native render hot path ("libdorktracer.so").
 *
 * This is the drop-in boundary of SURVEY.md section 8(b).  The reference (dorukb/Advanced-CPU-Raytracing)
 * has no FFI of its own; the narrowest seam is
 *     Vec3f Raytracer::RenderPixel(int i, int j, Camera& cam)            src/raytracer.hpp:19, raytracer.cpp:33-36
 * driven per pixel/sample by renderThreadMain (src/main.cpp:26-130) and followed by
 *     cam.GetTonemappedImage(W, H, hdr, ldr)                              src/main.cpp:190, camera.cpp:88-91.
 * A per-pixel call cannot feed a GPU, so the boundary is one frame-level call (dt_render) that replaces
 * src/main.cpp:164-192 (thread spawn ... tonemap).  Every struct below is a flat restatement of a reference
 * class (cited per struct); the host fills them from its own Scene (see INTEGRATION.md for the stub a
 * maintainer adds to main.cpp) — plain pointers and sizes only, no C++/torch types.
 *
 * Ownership: the caller owns every pointer it passes; dt_scene_create deep-copies.  The library writes only
 * into the output buffers / stats it is handed.  Errors: int status (0 = DT_OK), never throws across the
 * boundary; dt_last_error() returns a thread-local message.  There is NO CPU fallback: without a CUDA
 * device every compute entry point fails with DT_ERR_NO_DEVICE.
 */
#ifndef DORKTRACER_H
#define DORKTRACER_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DT_ABI_VERSION 2

/* ---- status codes ---- */
enum {
    DT_OK = 0,
    DT_ERR_INVALID = -1,      /* bad argument / inconsistent description          */
    DT_ERR_NO_DEVICE = -2,    /* no CUDA device (there is no CPU fallback)         */
    DT_ERR_CUDA = -3,         /* CUDA runtime error (message in dt_last_error)     */
    DT_ERR_OVERFLOW = -4,     /* wavefront queue overflow that retry could not fix */
    DT_ERR_UNSUPPORTED = -5
};

/* ---- Material (src/material.hpp:8-50); enum order is the reference's ---- */
enum { DT_MAT_MIRROR = 0, DT_MAT_DIELECTRIC = 1, DT_MAT_CONDUCTOR = 2, DT_MAT_EMISSIVE = 3, DT_MAT_DEFAULT = 4 };

typedef struct dt_material {
    int32_t type;                       /* DT_MAT_*                                         */
    int32_t brdf;                       /* index into dt_scene_desc.brdfs, -1 = no BRDF     */
    float ambient[3];
    float diffuse[3];
    float specular[3];
    float mirror[3];
    float phong_exponent;
    float refractive_index;
    float absorption_coefficient[3];
    float conductor_absorption_index;
    float roughness;
    float radiance[3];                  /* Emissive only (parser.cpp:1484-1487)             */
} dt_material;

/* ---- BRDFs (src/brdf.h:10-25 and the five subclasses) ---- */
enum { DT_BRDF_PHONG = 0, DT_BRDF_BLINN_PHONG = 1, DT_BRDF_MODIFIED_PHONG = 2,
       DT_BRDF_MODIFIED_BLINN_PHONG = 3, DT_BRDF_TORRANCE_SPARROW = 4 };

typedef struct dt_brdf {
    int32_t kind;                       /* DT_BRDF_*                                        */
    float exponent;
    int32_t flag;                       /* isEnergyConserving ("normalized") / kdFresnel    */
} dt_brdf;

/* ---- lights ---- */
typedef struct dt_point_light {         /* src/pointLight.h:9-20 */
    float position[3];
    float intensity[3];
} dt_point_light;

typedef struct dt_area_light {          /* src/areaLight.h:8-52; u,v = GetOrthonormalBasis(normal) done by the host ctor */
    float position[3];
    float normal[3];
    float radiance[3];
    float extent;
    float u[3];
    float v[3];
} dt_area_light;

typedef struct dt_directional_light {   /* src/directionalLight.h:8-22; dir already makeUnit()ed by the ctor */
    float dir[3];
    float radiance[3];
} dt_directional_light;

typedef struct dt_spot_light {          /* src/spotLight.h:10-62 */
    float pos[3];
    float dir[3];                       /* makeUnit()ed by the ctor */
    float intensity[3];
    float coverage_angle;               /* degrees */
    float falloff_angle;                /* degrees */
    double cos_half_falloff;
    double cos_half_coverage;
} dt_spot_light;

typedef struct dt_env_light {           /* src/sphericalEnvironmentLight.h:9-70 */
    int32_t image;                      /* index into images (HDR, lat-long) */
} dt_env_light;

typedef struct dt_mesh_light {          /* src/meshLight.h:10-52 */
    int32_t shape;                      /* index into shapes[] of the LightMesh (a DT_SHAPE_MESH) */
    int32_t id;                         /* Shape::id, compared with hitLightMeshId (raytracer.cpp:781) */
    float radiance[3];
} dt_mesh_light;

/* ---- images / textures ---- */
typedef struct dt_image {               /* src/LDRImage.h (uint8, `channels` interleaved) / src/HDRImage.h (float RGB) */
    int32_t width, height;
    int32_t channels;                   /* LDR: as loaded (1..4); HDR: 3 */
    int32_t is_hdr;                     /* 0: data is uint8_t[w*h*channels]; 1: data is float[w*h*3] */
    const void* data;
} dt_image;

enum { DT_TEX_IMAGE = 0, DT_TEX_PERLIN = 1 };
/* DecalMode (src/texture.h:9-19) */
enum { DT_DECAL_REPLACE_KD = 0, DT_DECAL_BLEND_KD = 1, DT_DECAL_REPLACE_KS = 2, DT_DECAL_REPLACE_BG = 3,
       DT_DECAL_REPLACE_NORMAL = 4, DT_DECAL_BUMP_NORMAL = 5, DT_DECAL_REPLACE_ALL = 6 };
enum { DT_INTERP_NEAREST = 0, DT_INTERP_BILINEAR = 1 };
enum { DT_NOISE_LINEAR = 0, DT_NOISE_ABSVAL = 1 };

typedef struct dt_texture {             /* src/texture.h, imageTexture.h, perlinTexture.h */
    int32_t kind;                       /* DT_TEX_*                                                   */
    int32_t decal_mode;                 /* DT_DECAL_*  (operationMode Blend iff DT_DECAL_BLEND_KD)    */
    int32_t image;                      /* image textures: index into images                          */
    int32_t interpolation;              /* DT_INTERP_*                                                */
    float normalizer;                   /* image: Normalizer (default 255); perlin: 1                 */
    float sample_multiplier;            /* BumpFactor                                                  */
    float noise_scale;                  /* perlin                                                      */
    int32_t noise_conversion;           /* DT_NOISE_*                                                  */
} dt_texture;

/* ---- geometry ---- */
typedef struct dt_face {                /* src/shape.hpp:103-112 (center/bbox are build-time only) */
    int32_t v0_id, v1_id, v2_id;        /* 1-based, resolved through vertex_offset like Mesh::GetVertex */
    float n[3];                         /* face normal exactly as Scene::computeFaceNormal produced it  */
    double area;
} dt_face;

typedef struct dt_bvh2_node {           /* src/bvh.hpp:11-26 with the raw pointers turned into indices  */
    float bmin[3], bmax[3];
    int32_t left, right;                /* index into the same node array, -1 = nullptr                 */
    uint32_t first_face, face_count;
} dt_bvh2_node;

typedef struct dt_mesh {                /* src/mesh.hpp:11-50: geometry + the BVH2 built by Mesh::ConstructBVH */
    const float* vertices;              /* xyz triples (Mesh::vertices)                                  */
    int32_t n_vertices;
    const float* uvs;                   /* uv pairs (Mesh::uv); n_uvs == 0 <=> "mesh has no UVs"         */
    int32_t n_uvs;
    int32_t vertex_offset, texture_offset;
    const dt_face* faces;               /* in post-ConstructBVH order (the canonical face ids)           */
    int32_t n_faces;
    const dt_bvh2_node* bvh;            /* node 0 = root                                                 */
    int32_t n_bvh_nodes;
    float bbox_min[3], bbox_max[3];     /* Mesh::bbox                                                    */
    double surface_area;                /* Mesh::surfaceArea                                             */
    const float* vertex_normals;        /* SURVEY 8f-4, optional: xyz per vertex (n_vertices triples) of a shadingMode="smooth" mesh,
                                           NULL otherwise.  The reference parses nothing of the kind and shades every mesh flat; the
                                           normals are used only by renders that pass DT_FLAG_SMOOTH_SHADING                   */
} dt_mesh;

enum { DT_SHAPE_MESH = 0, DT_SHAPE_INSTANCE = 1, DT_SHAPE_SPHERE = 2 };

typedef struct dt_shape {               /* src/shape.hpp:14-69 + Mesh / InstancedMesh / Sphere members  */
    int32_t kind;                       /* DT_SHAPE_*                                                   */
    int32_t id;                         /* Shape::id                                                    */
    int32_t mesh;                       /* MESH: index into meshes[]; others: -1                        */
    int32_t base_shape;                 /* INSTANCE: index into shapes[] of InstancedMesh::baseMesh     */
    int32_t material;                   /* 1-based material id (materials[matId-1], raytracer.cpp:73)   */
    int32_t tex_diffuse, tex_specular, tex_normal, tex_bump, tex_replace_all;   /* -1 = nullptr        */
    int32_t has_motion_blur;
    float motion_blur[3];
    double transform[16];               /* row-major 4x4 (Shape::transform)                             */
    double inverse_transform[16];
    double inverse_transpose_transform[16];
    float bbox_min[3], bbox_max[3];     /* INSTANCE: InstancedMesh::bbox (world space)                  */
    float center[3];                    /* SPHERE: vertex_data[center_vertex_id-1]                      */
    float radius;                       /* SPHERE                                                       */
} dt_shape;

typedef struct dt_scene_desc {          /* src/scene.h:32-89 */
    int32_t abi_version;                /* DT_ABI_VERSION */
    int32_t background_color[3];
    int32_t bg_texture;                 /* index into textures, -1 = nullptr */
    int32_t max_recursion_depth;
    float shadow_ray_epsilon;
    float ambient_light[3];

    const dt_material* materials;            int32_t n_materials;
    const dt_brdf* brdfs;                    int32_t n_brdfs;
    const dt_point_light* point_lights;      int32_t n_point_lights;
    const dt_area_light* area_lights;        int32_t n_area_lights;
    const dt_directional_light* directional_lights; int32_t n_directional_lights;
    const dt_spot_light* spot_lights;        int32_t n_spot_lights;
    const dt_env_light* env_lights;          int32_t n_env_lights;
    const dt_mesh_light* mesh_lights;        int32_t n_mesh_lights;
    const dt_image* images;                  int32_t n_images;
    const dt_texture* textures;              int32_t n_textures;
    const dt_mesh* meshes;                   int32_t n_meshes;
    /* shapes[0 .. n_mesh_shapes) = scene.meshes in order (Mesh, LightMesh, MeshInstance, Triangle:
       parser.cpp:348-349,453,510); shapes[n_mesh_shapes .. n_shapes) = scene.spheres.  This is the scan
       order of Raytracer::IntersectObjects (raytracer.cpp:625-643) and therefore the tie-break order. */
    const dt_shape* shapes;                  int32_t n_shapes;
    int32_t n_mesh_shapes;
} dt_scene_desc;

typedef struct dt_camera_desc {         /* src/camera.hpp:12-50, rendererParams.h, tonemapper.h:11-26 */
    float position[3], gaze[3], up[3], right[3];
    float q[3];                         /* m_q = top-left corner of the image plane (camera.cpp:61-72) */
    float left, right_, bottom, top;    /* m_left, m_right, m_bottom, m_top                            */
    float near_dist;
    int32_t width, height;
    int32_t samples_per_pixel;          /* not a perfect square: the samples past floor(sqrt(n))^2 go through sample position (0,0)
                                           like the reference's never-written `samples` entries (main.cpp:47,63-81)                  */
    float focus_distance, aperture_size;
    int32_t path_tracing, importance_sampling, next_event_estimation, russian_roulette;
                                        /* russian_roulette: a path the roulette can never end (NaN or >= 1 throughput in closed
                                           geometry; the reference recurses until its stack overflows) is cut 32768 bounces below
                                           depth 0                                                                                  */
    int32_t has_tonemapper;
    float tm_key, tm_burn, tm_saturation, tm_gamma;
} dt_camera_desc;

typedef struct dt_render_params {
    uint64_t seed;                      /* counter-based RNG key (the reference's RNG is unseeded + racy)        */
    int32_t tile_rank, tile_world;      /* strips s (8 consecutive 8x4-pixel tiles of a tile row, row-major) with
                                           s % tile_world == tile_rank are rendered (1 GPU: 0,1) */
    int32_t max_wave_rays;              /* primary rays per wave, 0 = default                                      */
    int32_t flags;                      /* DT_FLAG_*                                                               */
} dt_render_params;

/* dt_render_params.flags.  Everything that changes what a render call does travels here (or in dt_scene_options), not in
 * environment variables; the DT_* environment variables the library still reads are measurement knobs for A/B runs on the GPU
 * box (grid sizes, stream priorities, debug timing; listed in profiles/README.md) and never change results. */
enum {
    DT_FLAG_SKIP_TONEMAP = 1,       /* leave hdr untouched, do not write ldr from the tonemapper                                  */
    DT_FLAG_NO_SORT = 2,            /* disable the sort-by-material stage between closest-hit and shade (A/B measurement)          */
    DT_FLAG_SERIAL_WAVES = 4,       /* measurement mode: one kernel at a time, host sync per wave, so that the per-stage CUDA-event
                                       times in dt_stats are those of each kernel running ALONE (default: waves enqueued back to
                                       back, shadow(k) overlapping closest(k+1), stage times overlap)                              */
    DT_FLAG_PEER_FRAME = 8,         /* multi-GPU: resolve ONLY the owned tiles, straight into the frame buffers of the rank that
                                       called dt_frame_export (this rank's own buffers if it did not dt_frame_import): the gather
                                       is fused into the resolve kernel as P2P stores over NVLink, no collective and no full-frame
                                       exchange                                                                                    */
    DT_FLAG_JITTER_AA = 16,         /* multi-sample cameras keep the sub-pixel sample position.  The reference intends
                                       this but RenderPixel(int,int,..) truncates it away (main.cpp:83), so the default (parity)
                                       sends every sample through the pixel centre                                                 */
    DT_FLAG_REF_ROW_BANDS = 32,     /* parity with the reference's 8 row bands (main.cpp:15,38-39): the bottom H mod 8 rows get no
                                       camera rays and stay black.  Default: every row is rendered                                 */
    DT_FLAG_TEST_TIGHT_QUEUES = 64, /* test hook: size the wavefront queues of the first attempt for a ray-tree fan-out of 1, so
                                       that fanning scenes overflow them and take the retry path (dt_stats.retries > 0)            */
    DT_FLAG_FORCE_SORT = 128,       /* sort by material even in scenes with fewer than three materials                             */
    DT_FLAG_HOST_WAVE_LOOP = 256,   /* one host round trip per wave instead of the device-resident wave loop (A/B, debugging)       */
    DT_FLAG_FRAME_GRAPH = 512,      /* bounded-depth frames: replay the enqueued frame as a CUDA graph instead of ~50 launches      */
    DT_FLAG_PEER_HDR = 1024,        /* with DT_FLAG_PEER_FRAME: gather the radiance frame too when the camera has no tonemapper      */
    DT_FLAG_SORT_MATERIAL_ONLY = 2048, /* ignore DT_SORT_SPATIAL (the opt-in hit-cell sort, an A/B knob): sort the hits by material only */
    DT_FLAG_KEEP_WEIGHTLESS_PATHS = 4096, /* path tracing with Russian roulette: by default a hit whose path weight W is EXACTLY (0,0,0) is
                                       not shaded -- every radiance term below it is W times something, i.e. an exact zero, and the
                                       reference's roulette never ends such chains (raytracer.cpp:137-147), so they are a fifth of the
                                       rays of config 5.  The image is unchanged (only a NaN produced after the underflow would no longer
                                       poison its pixel); dt_stats then counts the rays actually traced.  With this flag the paths are
                                       followed as the reference follows them (ray counts comparable with the oracle's; A/B).  The same
                                       default / flag pair governs shadow rays whose contribution W*c is exactly zero (light below the
                                       horizon, vanishing lobe): path-traced frames do not trace them                                 */
    DT_FLAG_SMOOTH_SHADING = 8192    /* meshes that carry dt_mesh.vertex_normals (shadingMode="smooth" in the XML, which the
                                       reference ignores) are shaded with the barycentric interpolation of their vertex normals
                                       in place of the face normal.  Off by default: parity with the reference means flat shading */
};

typedef struct dt_stats {
    uint64_t rays_closest;              /* closest-hit queries traced (== Raytracer::IntersectObjects calls; under Russian roulette
                                           only with DT_FLAG_KEEP_WEIGHTLESS_PATHS, see there)                                */
    uint64_t rays_shadow;               /* occlusion queries traced (== Raytracer::CastShadowRay calls, same remark)          */
    uint64_t nan_pixels;
    uint32_t waves;
    uint32_t kernel_launches;           /* launches of this library's kernels inside the call           */
    float ms_total;                     /* CUDA-event time of all device work of the call               */
    float ms_generate, ms_traverse_closest, ms_traverse_shadow, ms_shade, ms_sort, ms_resolve, ms_tonemap;
    uint32_t launches_traverse_closest; /* number of closest-hit traversal launches                      */
    uint32_t retries;                   /* queue-overflow retries with a smaller wave                    */
    /* "Bulk" waves = waves whose closest-hit pass held at least half of max_wave_rays; filled by the host-synchronised loop
       (DT_FLAG_SERIAL_WAVES / DT_FLAG_HOST_WAVE_LOOP) only.  The launches a roofline figure is taken from: the thousands of
       tail waves of a Russian-roulette frame hold a handful of rays each and are all launch latency. */
    uint32_t bulk_waves;
    float ms_bulk_closest, ms_bulk_shadow;
    uint32_t pad_;
    uint64_t rays_bulk_closest, rays_bulk_shadow;
} dt_stats;

typedef struct dt_scene dt_scene;       /* opaque; owns all device memory of one GPU */

/* Select the CUDA device used by subsequently created scenes (one process per GPU); <0 on error. */
int dt_gpu_init(int device);
int dt_device_count(void);

int dt_scene_create(const dt_scene_desc* desc, dt_scene** out);
/* Same with options.  gpu_flatten_min_faces: meshes of at least this many faces get their BVH8 from the GPU flattener
 * (0 = default 32768, < 0 = never: host flattener only; both produce identical bytes, see dt_scene_accel_checksum). */
typedef struct dt_scene_options {
    int32_t gpu_flatten_min_faces;
    int32_t reserved[7];                /* must be zero */
} dt_scene_options;
int dt_scene_create_opts(const dt_scene_desc* desc, const dt_scene_options* opts, dt_scene** out);
void dt_scene_destroy(dt_scene* scene);

/* Render one camera.  ldr_rgb: W*H*3 bytes, row-major RGB, top row first (main.cpp:109,146); hdr_rgb: W*H*3
 * floats or NULL (main.cpp:148-152).  When cam->has_tonemapper the tonemapped image is written to ldr_rgb
 * (replaces main.cpp:190), else the LDR clamp of main.cpp:118-125.  With tile_world > 1 only this rank's
 * tiles are written (others are left zero). */
int dt_render(dt_scene* scene, const dt_camera_desc* cam, const dt_render_params* params,
              uint8_t* ldr_rgb, float* hdr_rgb, dt_stats* stats);

/* Same, but the resolved radiance stays in device memory: *hdr_dev (float W*H*3) is owned by the scene and
 * valid until the next render call.  Used by the multi-GPU gather (NCCL) without a host round trip. */
int dt_render_device(dt_scene* scene, const dt_camera_desc* cam, const dt_render_params* params,
                     float** hdr_dev, dt_stats* stats);

/* Final stage on device memory: LDR clamp (main.cpp:118-125) or Reinhard tonemap (tonemapper.h:28-119) of a
 * full-frame radiance buffer hdr_dev (device pointer, W*H*3 floats) into host ldr_rgb. */
int dt_finish_device(dt_scene* scene, const dt_camera_desc* cam, const float* hdr_dev,
                     uint8_t* ldr_rgb, dt_stats* stats);

/* Multi-GPU gather over peer memory (one process per GPU on one box).  The destination rank exports its
 * frame buffers as CUDA IPC handles (plain bytes: ship them with any host-side transport), the other ranks import them;
 * renders issued with DT_FLAG_PEER_FRAME then store their tiles directly into the destination's memory.  The caller
 * orders "all ranks finished rendering" before "destination reads the frame" (one barrier); dt_frame_finish then
 * tonemaps (if the camera has a tonemapper) and copies the complete LDR frame to the host (ldr_rgb == NULL: the finished
 * frame stays in device memory; used to time the device work of a frame without the D2H copy). */
typedef struct dt_frame_handle {
    unsigned char hdr[64];              /* cudaIpcMemHandle_t of the float W*H*3 radiance frame */
    unsigned char ldr[64];              /* cudaIpcMemHandle_t of the uint8 W*H*3 LDR frame      */
    int32_t width, height;
} dt_frame_handle;
int dt_frame_export(dt_scene* scene, int32_t width, int32_t height, dt_frame_handle* out);
int dt_frame_import(dt_scene* scene, const dt_frame_handle* in);
int dt_frame_release(dt_scene* scene);
int dt_frame_finish(dt_scene* scene, const dt_camera_desc* cam, uint8_t* ldr_rgb, dt_stats* stats);

/* One process, several GPUs ("the library drives all GPUs internally").  The reference's main() is one process
 * that splits the rows of a frame over its threads (main.cpp:38-39,164-185); dt_multi splits its strips over the first
 * n_devices GPUs of the box (<= 0: all visible): scene replicated, one host thread per GPU, every GPU's resolve kernel stores
 * its strips into device 0's frame through peer access over NVLink, device 0 tonemaps / copies the frame out.  Same outputs
 * and semantics as dt_render (tile_rank / tile_world of params are ignored). */
typedef struct dt_multi dt_multi;
int dt_multi_create(const dt_scene_desc* desc, int n_devices, dt_multi** out);
int dt_multi_render(dt_multi* m, const dt_camera_desc* cam, const dt_render_params* params,
                    uint8_t* ldr_rgb, float* hdr_rgb, dt_stats* stats);
int dt_multi_device_count(const dt_multi* m);
void dt_multi_destroy(dt_multi* m);

/* Parity/debug: primary-ray closest hits.  shape = index into desc.shapes (-1 miss), face = canonical
 * (post-build) face index of the (base) mesh or -1 for spheres, t = hit distance (INFINITY on miss). */
int dt_primary_hits(dt_scene* scene, const dt_camera_desc* cam, int32_t* shape, int32_t* face, float* t);

/* Generic ray queries (host arrays, n rays): origins/dirs are xyz triples.  Closest: as IntersectObjects.
 * Occlusion: as CastShadowRay with light distance tmax[i] (INFINITY for directional). */
int dt_trace_closest(dt_scene* scene, const float* origins, const float* dirs, int64_t n,
                     int32_t* shape, int32_t* face, float* t);
int dt_trace_occluded(dt_scene* scene, const float* origins, const float* dirs, const float* tmax, int64_t n,
                      uint8_t* occluded);

/* Stand-alone tonemapper (tonemapper.h:28-119) on host buffers. */
int dt_tonemap(const float* hdr_rgb, int32_t width, int32_t height, float key, float burn, float saturation,
               float gamma, uint8_t* ldr_rgb);

/* Mesh::ConstructBVH / RecursiveBVHBuild / RecomputeBoundingBox (mesh.cpp:23-156) on the GPU: the
 * reference's longest-axis spatial-midpoint build INCLUDING the face permutation of its in-place two-pointer partition
 * (mesh.cpp:92-102), which defines the canonical face ids, and its node numbering (mesh.cpp:107-121).  Host arrays:
 * centers = Face::center xyz per face, face_boxes = Face::bbox (min xyz, max xyz) per face, root_min / root_max = Mesh::bbox
 * (the root box is not recomputed, mesh.cpp:31-35).  Out: face_order[new] = old face index, nodes[0 .. *n_nodes)
 * (node_capacity >= 2 * n_faces - 1, mesh.cpp:29), device time of the build without the copies in *ms_device (may be NULL).
 * Identical to the reference's build except for the sign of a zero box coordinate (min / max are order-independent here).
 * Rejects non-finite inputs. */
int dt_bvh2_build(int32_t n_faces, const float* centers, const float* face_boxes, const float* root_min, const float* root_max,
                  uint32_t* face_order, dt_bvh2_node* nodes, uint32_t node_capacity, uint32_t* n_nodes, float* ms_device);

/* Parity/debug: checksums (sum of 32-bit words, position-weighted sum) of the scene's device-resident acceleration arrays
 * -- BVH8 nodes, leaf-order triangles, reference leaf boxes, face -> primitive map -- then the node and primitive counts.
 * The host flattener and the GPU flattener (DT_GPU_FLATTEN_MIN_FACES) must agree on all ten. */
int dt_scene_accel_checksum(dt_scene* scene, uint64_t out[10]);

/* The CUDA stream (cudaStream_t) all device work of this scene is enqueued on, so callers can bracket calls
 * with their own CUDA events / order collectives after a render without a device-wide sync. */
void* dt_scene_stream(dt_scene* scene);

const char* dt_last_error(void);
const char* dt_version(void);


#ifdef __cplusplus
}
#endif
#endif /* DORKTRACER_H */
