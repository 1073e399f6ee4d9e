/*
 * dorktracer_host.h — C ABI of the host-side mirror ("libdthost.so") of the reference's scene layer.
 *
 * The reference keeps its own tinyxml2/happly parser and Scene/Camera/Material/light classes
 * (src/parser.cpp, src/scene.h); a maintainer wires those to dorktracer.h with the stub in INTEGRATION.md.
 * /root/reference does not travel to the GPU box and its sources may not be copied, so this library is an
 * independent C++ restatement of exactly that layer — Scene::loadFromXml (parser.cpp:26-577), the face /
 * transform / bbox helpers (parser.cpp:579-826), parse{Cameras,Lights,BRDFs,Materials,Meshes}
 * (parser.cpp:828-1636), Camera::Setup* (camera.cpp:5-72) and Mesh::ConstructBVH (mesh.cpp:23-156) — whose
 * only output is the flat dt_scene_desc / dt_camera_desc the hot path consumes.  It does no rendering.
 */
#ifndef DORKTRACER_HOST_H
#define DORKTRACER_HOST_H

#include "dorktracer.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dth_scene dth_scene;     /* opaque: owns every array the returned descs point into */

/* Scene::loadFromXml (parser.cpp:26).  Relative plyFile / image paths are tried against the current
 * directory first (as the reference does) and then against the XML file's directory. */
int dth_scene_load_xml(const char* xml_path, dth_scene** out);
void dth_scene_free(dth_scene* scene);

/* Optional replacement for the host's own Mesh::ConstructBVH (mesh.cpp:23-156) during dth_scene_load_xml: the signature of
 * dt_bvh2_build (dorktracer.h).  Meshes with fewer than `min_faces` faces are still built on the host.  NULL restores the
 * host build.  A builder error fails the load (no silent fallback). */
typedef int (*dth_bvh_builder)(int32_t n_faces, const float* centers, const float* face_boxes, const float* root_min, const float* root_max,
                               uint32_t* face_order, dt_bvh2_node* nodes, uint32_t node_capacity, uint32_t* n_nodes, float* ms_device);
void dth_set_bvh_builder(dth_bvh_builder builder, int32_t min_faces);
/* seconds spent in BVH construction (host or builder, including the face permutation) during the last load on this thread */
double dth_last_bvh_build_seconds(void);

const dt_scene_desc* dth_scene_desc(const dth_scene* scene);
int dth_scene_num_cameras(const dth_scene* scene);
const dt_camera_desc* dth_scene_camera(const dth_scene* scene, int index);
const char* dth_scene_camera_image_name(const dth_scene* scene, int index);

/* Replace the pixel data of image `index` (for image formats this library does not decode itself:
 * the caller decodes and hands over raw pixels; the data is copied). */
int dth_scene_set_image(dth_scene* scene, int index, int width, int height, int channels, int is_hdr,
                        const void* data);
/* File name of image `index` as written in the XML, and whether the loader could decode it itself. */
const char* dth_scene_image_path(const dth_scene* scene, int index);
int dth_scene_image_loaded(const dth_scene* scene, int index);

/* Camera::SetupLookAt / SetupDefault (camera.cpp:5-72) for callers that build cameras programmatically. */
int dth_camera_look_at(const float pos[3], const float gaze_point[3], const float up[3], float near_dist,
                       float fov_y, int width, int height, dt_camera_desc* out);
int dth_camera_default(const float pos[3], const float gaze_dir[3], const float up[3], const float near_plane[4],
                       float near_dist, int width, int height, dt_camera_desc* out);

/* Minimal image writers for the stand-alone driver (main.cpp:191-195 uses stb_image_write). */
int dth_write_png(const char* path, int width, int height, const uint8_t* rgb);

/* Band-parallel PNG encoder (rows deflated on n_threads threads into one zlib stream; 0 = all cores, at most 16) and the flat
 * RGBE .hdr writer (what stbi_write_hdr emits for main.cpp:191). */
int dth_write_png_parallel(const char* path, int width, int height, const uint8_t* rgb, int n_threads);
int dth_write_hdr(const char* path, int width, int height, const float* rgb);

/* Asynchronous output: the reference encodes PNG / HDR on the main thread after every camera, inside its timed
 * region (main.cpp:186-195).  A dth_writer owns `n_files_in_parallel` worker threads (1..8); submit copies the pixels and returns
 * at once, so the caller renders the next camera while the previous image is encoded (each PNG on `encode_threads` threads,
 * 0 = all cores, at most 16).  dth_writer_wait blocks until every submitted file is on disk, returns the first error (text in
 * dth_last_error) and optionally the seconds the workers spent encoding.  dth_writer_destroy waits, then joins the workers. */
typedef struct dth_writer dth_writer;
dth_writer* dth_writer_create(int n_files_in_parallel, int encode_threads);
int dth_writer_submit_png(dth_writer* writer, const char* path, int width, int height, const uint8_t* rgb);
int dth_writer_submit_hdr(dth_writer* writer, const char* path, int width, int height, const float* rgb);
int dth_writer_wait(dth_writer* writer, double* busy_seconds);
void dth_writer_destroy(dth_writer* writer);

const char* dth_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
