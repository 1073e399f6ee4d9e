"""ctypes mirror of include/dorktracer.h and include/dorktracer_host.h.

Python is plumbing only (tests, bench, smoke): every struct is the flat C struct of the header, the
libraries are loaded from the package directory (built in-tree by the Makefile / __graft_entry__.build()).
There is no Python or CPU fallback for the render path: `load_dorktracer()` raises if the CUDA library is
missing.
"""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPO_DIR = os.path.dirname(PKG_DIR)

c_f3 = C.c_float * 3
c_f4 = C.c_float * 4
c_d16 = C.c_double * 16
c_i3 = C.c_int32 * 3

DT_OK = 0
DT_ERR_INVALID, DT_ERR_NO_DEVICE, DT_ERR_CUDA, DT_ERR_OVERFLOW, DT_ERR_UNSUPPORTED = -1, -2, -3, -4, -5
DT_SHAPE_MESH, DT_SHAPE_INSTANCE, DT_SHAPE_SPHERE = 0, 1, 2
DT_MAT_MIRROR, DT_MAT_DIELECTRIC, DT_MAT_CONDUCTOR, DT_MAT_EMISSIVE, DT_MAT_DEFAULT = range(5)
DT_FLAG_SKIP_TONEMAP = 1
DT_FLAG_NO_SORT = 2
DT_FLAG_SERIAL_WAVES = 4
DT_FLAG_PEER_FRAME = 8
DT_FLAG_JITTER_AA = 16
DT_FLAG_REF_ROW_BANDS = 32
DT_FLAG_TEST_TIGHT_QUEUES = 64
DT_FLAG_FORCE_SORT = 128
DT_FLAG_HOST_WAVE_LOOP = 256
DT_FLAG_FRAME_GRAPH = 512
DT_FLAG_PEER_HDR = 1024
DT_FLAG_SORT_MATERIAL_ONLY = 2048
DT_FLAG_KEEP_WEIGHTLESS_PATHS = 4096
DT_FLAG_SMOOTH_SHADING = 8192


class dt_scene_options(C.Structure):
    _fields_ = [("gpu_flatten_min_faces", C.c_int32), ("reserved", C.c_int32 * 7)]


class dt_material(C.Structure):
    _fields_ = [("type", C.c_int32), ("brdf", C.c_int32), ("ambient", c_f3), ("diffuse", c_f3), ("specular", c_f3),
                ("mirror", c_f3), ("phong_exponent", C.c_float), ("refractive_index", C.c_float),
                ("absorption_coefficient", c_f3), ("conductor_absorption_index", C.c_float),
                ("roughness", C.c_float), ("radiance", c_f3)]


class dt_brdf(C.Structure):
    _fields_ = [("kind", C.c_int32), ("exponent", C.c_float), ("flag", C.c_int32)]


class dt_point_light(C.Structure):
    _fields_ = [("position", c_f3), ("intensity", c_f3)]


class dt_area_light(C.Structure):
    _fields_ = [("position", c_f3), ("normal", c_f3), ("radiance", c_f3), ("extent", C.c_float), ("u", c_f3), ("v", c_f3)]


class dt_directional_light(C.Structure):
    _fields_ = [("dir", c_f3), ("radiance", c_f3)]


class dt_spot_light(C.Structure):
    _fields_ = [("pos", c_f3), ("dir", c_f3), ("intensity", c_f3), ("coverage_angle", C.c_float),
                ("falloff_angle", C.c_float), ("cos_half_falloff", C.c_double), ("cos_half_coverage", C.c_double)]


class dt_env_light(C.Structure):
    _fields_ = [("image", C.c_int32)]


class dt_mesh_light(C.Structure):
    _fields_ = [("shape", C.c_int32), ("id", C.c_int32), ("radiance", c_f3)]


class dt_image(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32), ("is_hdr", C.c_int32),
                ("data", C.c_void_p)]


class dt_texture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("decal_mode", C.c_int32), ("image", C.c_int32), ("interpolation", C.c_int32),
                ("normalizer", C.c_float), ("sample_multiplier", C.c_float), ("noise_scale", C.c_float),
                ("noise_conversion", C.c_int32)]


class dt_face(C.Structure):
    _fields_ = [("v0_id", C.c_int32), ("v1_id", C.c_int32), ("v2_id", C.c_int32), ("n", c_f3), ("area", C.c_double)]


class dt_bvh2_node(C.Structure):
    _fields_ = [("bmin", c_f3), ("bmax", c_f3), ("left", C.c_int32), ("right", C.c_int32),
                ("first_face", C.c_uint32), ("face_count", C.c_uint32)]


class dt_mesh(C.Structure):
    _fields_ = [("vertices", C.POINTER(C.c_float)), ("n_vertices", C.c_int32),
                ("uvs", C.POINTER(C.c_float)), ("n_uvs", C.c_int32),
                ("vertex_offset", C.c_int32), ("texture_offset", C.c_int32),
                ("faces", C.POINTER(dt_face)), ("n_faces", C.c_int32),
                ("bvh", C.POINTER(dt_bvh2_node)), ("n_bvh_nodes", C.c_int32),
                ("bbox_min", c_f3), ("bbox_max", c_f3), ("surface_area", C.c_double),
                ("vertex_normals", C.POINTER(C.c_float))]


class dt_shape(C.Structure):
    _fields_ = [("kind", C.c_int32), ("id", C.c_int32), ("mesh", C.c_int32), ("base_shape", C.c_int32),
                ("material", C.c_int32),
                ("tex_diffuse", C.c_int32), ("tex_specular", C.c_int32), ("tex_normal", C.c_int32),
                ("tex_bump", C.c_int32), ("tex_replace_all", C.c_int32),
                ("has_motion_blur", C.c_int32), ("motion_blur", c_f3),
                ("transform", c_d16), ("inverse_transform", c_d16), ("inverse_transpose_transform", c_d16),
                ("bbox_min", c_f3), ("bbox_max", c_f3), ("center", c_f3), ("radius", C.c_float)]


class dt_scene_desc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("background_color", c_i3), ("bg_texture", C.c_int32),
                ("max_recursion_depth", C.c_int32), ("shadow_ray_epsilon", C.c_float), ("ambient_light", c_f3),
                ("materials", C.POINTER(dt_material)), ("n_materials", C.c_int32),
                ("brdfs", C.POINTER(dt_brdf)), ("n_brdfs", C.c_int32),
                ("point_lights", C.POINTER(dt_point_light)), ("n_point_lights", C.c_int32),
                ("area_lights", C.POINTER(dt_area_light)), ("n_area_lights", C.c_int32),
                ("directional_lights", C.POINTER(dt_directional_light)), ("n_directional_lights", C.c_int32),
                ("spot_lights", C.POINTER(dt_spot_light)), ("n_spot_lights", C.c_int32),
                ("env_lights", C.POINTER(dt_env_light)), ("n_env_lights", C.c_int32),
                ("mesh_lights", C.POINTER(dt_mesh_light)), ("n_mesh_lights", C.c_int32),
                ("images", C.POINTER(dt_image)), ("n_images", C.c_int32),
                ("textures", C.POINTER(dt_texture)), ("n_textures", C.c_int32),
                ("meshes", C.POINTER(dt_mesh)), ("n_meshes", C.c_int32),
                ("shapes", C.POINTER(dt_shape)), ("n_shapes", C.c_int32),
                ("n_mesh_shapes", C.c_int32)]


class dt_camera_desc(C.Structure):
    _fields_ = [("position", c_f3), ("gaze", c_f3), ("up", c_f3), ("right", c_f3), ("q", c_f3),
                ("left", C.c_float), ("right_", C.c_float), ("bottom", C.c_float), ("top", C.c_float),
                ("near_dist", C.c_float), ("width", C.c_int32), ("height", C.c_int32),
                ("samples_per_pixel", C.c_int32), ("focus_distance", C.c_float), ("aperture_size", C.c_float),
                ("path_tracing", C.c_int32), ("importance_sampling", C.c_int32),
                ("next_event_estimation", C.c_int32), ("russian_roulette", C.c_int32),
                ("has_tonemapper", C.c_int32), ("tm_key", C.c_float), ("tm_burn", C.c_float),
                ("tm_saturation", C.c_float), ("tm_gamma", C.c_float)]


class dt_render_params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("tile_rank", C.c_int32), ("tile_world", C.c_int32),
                ("max_wave_rays", C.c_int32), ("flags", C.c_int32)]


class dt_stats(C.Structure):
    _fields_ = [("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("nan_pixels", C.c_uint64),
                ("waves", C.c_uint32), ("kernel_launches", C.c_uint32), ("ms_total", C.c_float),
                ("ms_generate", C.c_float), ("ms_traverse_closest", C.c_float), ("ms_traverse_shadow", C.c_float),
                ("ms_shade", C.c_float), ("ms_sort", C.c_float), ("ms_resolve", C.c_float), ("ms_tonemap", C.c_float),
                ("launches_traverse_closest", C.c_uint32), ("retries", C.c_uint32),
                ("bulk_waves", C.c_uint32), ("ms_bulk_closest", C.c_float), ("ms_bulk_shadow", C.c_float), ("pad_", C.c_uint32),
                ("rays_bulk_closest", C.c_uint64), ("rays_bulk_shadow", C.c_uint64)]


# Every symbol include/dorktracer.h declares (tests check that the library exports all of them).
class dt_frame_handle(C.Structure):
    _fields_ = [("hdr", C.c_ubyte * 64), ("ldr", C.c_ubyte * 64), ("width", C.c_int32), ("height", C.c_int32)]


DORKTRACER_SYMBOLS = [
    "dt_gpu_init", "dt_device_count", "dt_scene_create", "dt_scene_create_opts", "dt_scene_destroy", "dt_render", "dt_render_device",
    "dt_finish_device", "dt_primary_hits", "dt_trace_closest", "dt_trace_occluded", "dt_tonemap",
    "dt_scene_stream", "dt_last_error", "dt_version",
    "dt_multi_create", "dt_multi_render", "dt_multi_device_count", "dt_multi_destroy",
    "dt_frame_export", "dt_frame_import", "dt_frame_release", "dt_frame_finish", "dt_bvh2_build", "dt_scene_accel_checksum",
]
DTHOST_SYMBOLS = [
    "dth_scene_load_xml", "dth_scene_free", "dth_scene_desc", "dth_scene_num_cameras", "dth_scene_camera",
    "dth_scene_camera_image_name", "dth_scene_set_image", "dth_scene_image_path", "dth_scene_image_loaded",
    "dth_camera_look_at", "dth_camera_default", "dth_write_png", "dth_last_error",
    "dth_set_bvh_builder", "dth_last_bvh_build_seconds",
    "dth_write_png_parallel", "dth_write_hdr",
    "dth_writer_create", "dth_writer_submit_png", "dth_writer_submit_hdr", "dth_writer_wait", "dth_writer_destroy",
]

_libs = {}


def _load(path, what):
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError("%s not built: %s is missing (run `make` or __graft_entry__.build())" % (what, path))
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    _libs[path] = lib
    return lib


def load_dthost():
    lib = _load(os.path.join(PKG_DIR, "libdthost.so"), "host library")
    vp = C.c_void_p
    lib.dth_scene_load_xml.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.dth_scene_load_xml.restype = C.c_int
    lib.dth_scene_free.argtypes = [vp]
    lib.dth_scene_free.restype = None
    lib.dth_scene_desc.argtypes = [vp]
    lib.dth_scene_desc.restype = C.POINTER(dt_scene_desc)
    lib.dth_scene_num_cameras.argtypes = [vp]
    lib.dth_scene_num_cameras.restype = C.c_int
    lib.dth_scene_camera.argtypes = [vp, C.c_int]
    lib.dth_scene_camera.restype = C.POINTER(dt_camera_desc)
    lib.dth_scene_camera_image_name.argtypes = [vp, C.c_int]
    lib.dth_scene_camera_image_name.restype = C.c_char_p
    lib.dth_scene_set_image.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.dth_scene_set_image.restype = C.c_int
    lib.dth_scene_image_path.argtypes = [vp, C.c_int]
    lib.dth_scene_image_path.restype = C.c_char_p
    lib.dth_scene_image_loaded.argtypes = [vp, C.c_int]
    lib.dth_scene_image_loaded.restype = C.c_int
    lib.dth_camera_look_at.argtypes = [c_f3, c_f3, c_f3, C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(dt_camera_desc)]
    lib.dth_camera_look_at.restype = C.c_int
    lib.dth_camera_default.argtypes = [c_f3, c_f3, c_f3, c_f4, C.c_float, C.c_int, C.c_int, C.POINTER(dt_camera_desc)]
    lib.dth_camera_default.restype = C.c_int
    lib.dth_write_png.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
    lib.dth_write_png.restype = C.c_int
    lib.dth_last_error.argtypes = []
    lib.dth_last_error.restype = C.c_char_p
    lib.dth_set_bvh_builder.argtypes = [C.c_void_p, C.c_int32]
    lib.dth_set_bvh_builder.restype = None
    lib.dth_last_bvh_build_seconds.argtypes = []
    lib.dth_last_bvh_build_seconds.restype = C.c_double
    lib.dth_write_png_parallel.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.dth_write_png_parallel.restype = C.c_int
    lib.dth_write_hdr.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
    lib.dth_write_hdr.restype = C.c_int
    lib.dth_writer_create.argtypes = [C.c_int, C.c_int]
    lib.dth_writer_create.restype = vp
    lib.dth_writer_submit_png.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_void_p]
    lib.dth_writer_submit_png.restype = C.c_int
    lib.dth_writer_submit_hdr.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_void_p]
    lib.dth_writer_submit_hdr.restype = C.c_int
    lib.dth_writer_wait.argtypes = [vp, C.POINTER(C.c_double)]
    lib.dth_writer_wait.restype = C.c_int
    lib.dth_writer_destroy.argtypes = [vp]
    lib.dth_writer_destroy.restype = None
    return lib


def load_dorktracer():
    """The CUDA hot path.  Fails loudly when the extension is missing: there is no fallback."""
    lib = _load(os.path.join(PKG_DIR, "libdorktracer.so"), "CUDA library (libdorktracer.so)")
    vp = C.c_void_p
    lib.dt_gpu_init.argtypes = [C.c_int]
    lib.dt_gpu_init.restype = C.c_int
    lib.dt_device_count.argtypes = []
    lib.dt_device_count.restype = C.c_int
    lib.dt_scene_create.argtypes = [C.POINTER(dt_scene_desc), C.POINTER(vp)]
    lib.dt_scene_create.restype = C.c_int
    lib.dt_scene_create_opts.argtypes = [C.POINTER(dt_scene_desc), C.POINTER(dt_scene_options), C.POINTER(vp)]
    lib.dt_scene_create_opts.restype = C.c_int
    lib.dt_scene_destroy.argtypes = [vp]
    lib.dt_scene_destroy.restype = None
    lib.dt_render.argtypes = [vp, C.POINTER(dt_camera_desc), C.POINTER(dt_render_params), C.c_void_p, C.c_void_p, C.POINTER(dt_stats)]
    lib.dt_render.restype = C.c_int
    lib.dt_render_device.argtypes = [vp, C.POINTER(dt_camera_desc), C.POINTER(dt_render_params), C.POINTER(C.c_void_p), C.POINTER(dt_stats)]
    lib.dt_render_device.restype = C.c_int
    lib.dt_finish_device.argtypes = [vp, C.POINTER(dt_camera_desc), C.c_void_p, C.c_void_p, C.POINTER(dt_stats)]
    lib.dt_finish_device.restype = C.c_int
    lib.dt_primary_hits.argtypes = [vp, C.POINTER(dt_camera_desc), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dt_primary_hits.restype = C.c_int
    lib.dt_trace_closest.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dt_trace_closest.restype = C.c_int
    lib.dt_trace_occluded.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.dt_trace_occluded.restype = C.c_int
    lib.dt_tonemap.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
    lib.dt_tonemap.restype = C.c_int
    lib.dt_multi_create.argtypes = [C.POINTER(dt_scene_desc), C.c_int, C.POINTER(vp)]
    lib.dt_multi_create.restype = C.c_int
    lib.dt_multi_render.argtypes = [vp, C.POINTER(dt_camera_desc), C.POINTER(dt_render_params), C.c_void_p, C.c_void_p, C.POINTER(dt_stats)]
    lib.dt_multi_render.restype = C.c_int
    lib.dt_multi_device_count.argtypes = [vp]
    lib.dt_multi_device_count.restype = C.c_int
    lib.dt_multi_destroy.argtypes = [vp]
    lib.dt_multi_destroy.restype = None
    lib.dt_frame_export.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(dt_frame_handle)]
    lib.dt_frame_export.restype = C.c_int
    lib.dt_frame_import.argtypes = [vp, C.POINTER(dt_frame_handle)]
    lib.dt_frame_import.restype = C.c_int
    lib.dt_frame_release.argtypes = [vp]
    lib.dt_frame_release.restype = C.c_int
    lib.dt_frame_finish.argtypes = [vp, C.POINTER(dt_camera_desc), C.c_void_p, C.POINTER(dt_stats)]
    lib.dt_frame_finish.restype = C.c_int
    lib.dt_bvh2_build.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                  C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
    lib.dt_bvh2_build.restype = C.c_int
    lib.dt_scene_accel_checksum.argtypes = [vp, C.POINTER(C.c_uint64 * 10)]
    lib.dt_scene_accel_checksum.restype = C.c_int
    lib.dt_scene_stream.argtypes = [vp]
    lib.dt_scene_stream.restype = C.c_void_p
    lib.dt_last_error.argtypes = []
    lib.dt_last_error.restype = C.c_char_p
    lib.dt_version.argtypes = []
    lib.dt_version.restype = C.c_char_p
    return lib
