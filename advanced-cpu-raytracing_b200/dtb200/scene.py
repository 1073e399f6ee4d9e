"""Thin Python handles over the two C ABIs (host mirror + CUDA hot path).  Plumbing for tests/bench only."""
import ctypes as C
import os

import numpy as np

from . import capi


class HostScene:
    """dth_scene handle: Scene::loadFromXml equivalent (include/dorktracer_host.h)."""

    def __init__(self, xml_path, gpu_build=False, gpu_build_min_faces=4096):
        """gpu_build: Mesh::ConstructBVH runs on the GPU (dt_bvh2_build) for meshes of at least gpu_build_min_faces faces."""
        self.lib = capi.load_dthost()
        self.handle = C.c_void_p()
        if gpu_build:
            dt = capi.load_dorktracer()
            self.lib.dth_set_bvh_builder(C.cast(dt.dt_bvh2_build, C.c_void_p), int(gpu_build_min_faces))
        try:
            rc = self.lib.dth_scene_load_xml(os.fsencode(xml_path), C.byref(self.handle))
        finally:
            if gpu_build:
                self.lib.dth_set_bvh_builder(None, 0)
        if rc != 0:
            msg = self.lib.dth_last_error().decode()
            if gpu_build:
                msg += " / " + capi.load_dorktracer().dt_last_error().decode()
            raise RuntimeError("dth_scene_load_xml(%s) failed: %s" % (xml_path, msg))
        self.bvh_build_seconds = float(self.lib.dth_last_bvh_build_seconds())
        self.xml_path = xml_path
        self._fill_missing_images()

    def _fill_missing_images(self):
        """Images the C++ loader cannot decode (jpg, ...) are decoded with PIL and handed over as raw pixels."""
        d = self.desc
        for i in range(d.n_images):
            if self.lib.dth_scene_image_loaded(self.handle, i):
                continue
            name = self.lib.dth_scene_image_path(self.handle, i).decode()
            base = os.path.dirname(os.path.abspath(self.xml_path))
            cands = [os.path.join("inputs", name), os.path.join(base, "inputs", name), os.path.join(base, name)]
            path = next((c for c in cands if os.path.exists(c)), None)
            if path is None:
                raise RuntimeError("image %s not found" % name)
            from PIL import Image
            im = Image.open(path)
            if im.mode not in ("L", "LA", "RGB", "RGBA"):
                im = im.convert("RGB")
            a = np.ascontiguousarray(np.array(im, dtype=np.uint8))
            ch = 1 if a.ndim == 2 else a.shape[2]
            rc = self.lib.dth_scene_set_image(self.handle, i, a.shape[1], a.shape[0], ch, 0, a.ctypes.data_as(C.c_void_p))
            if rc != 0:
                raise RuntimeError(self.lib.dth_last_error().decode())

    @property
    def desc(self):
        return self.lib.dth_scene_desc(self.handle).contents

    @property
    def desc_ptr(self):
        return self.lib.dth_scene_desc(self.handle)

    @property
    def num_cameras(self):
        return self.lib.dth_scene_num_cameras(self.handle)

    def camera(self, i=0):
        p = self.lib.dth_scene_camera(self.handle, i)
        if not p:
            raise IndexError(i)
        cam = capi.dt_camera_desc()
        C.memmove(C.byref(cam), p, C.sizeof(cam))
        return cam

    def image_name(self, i=0):
        return self.lib.dth_scene_camera_image_name(self.handle, i).decode()

    def n_triangles(self):
        d = self.desc
        return sum(d.meshes[i].n_faces for i in range(d.n_meshes))

    def close(self):
        if self.handle:
            self.lib.dth_scene_free(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuScene:
    """dt_scene handle on the CUDA hot path (include/dorktracer.h).  No fallback: raises without the library/GPU."""

    def __init__(self, host_scene, device=None, gpu_flatten_min_faces=0):
        """gpu_flatten_min_faces: dt_scene_options (0 = library default, < 0 = host flattener only)."""
        self.lib = capi.load_dorktracer()
        if device is not None:
            rc = self.lib.dt_gpu_init(int(device))
            if rc < 0:
                raise RuntimeError("dt_gpu_init failed: %s" % self.lib.dt_last_error().decode())
        self.host = host_scene
        self.handle = C.c_void_p()
        opts = capi.dt_scene_options()
        opts.gpu_flatten_min_faces = int(gpu_flatten_min_faces)
        rc = self.lib.dt_scene_create_opts(host_scene.desc_ptr, C.byref(opts), C.byref(self.handle))
        if rc != 0:
            raise RuntimeError("dt_scene_create failed (%d): %s" % (rc, self.lib.dt_last_error().decode()))

    def _err(self, what, rc):
        raise RuntimeError("%s failed (%d): %s" % (what, rc, self.lib.dt_last_error().decode()))

    def render(self, cam, seed=1234, tile_rank=0, tile_world=1, flags=0, max_wave_rays=0, ldr=None, hdr=None, want_hdr=True):
        W, H = cam.width, cam.height
        if ldr is None:
            ldr = np.zeros((H, W, 3), np.uint8)
        if hdr is None and want_hdr:
            hdr = np.zeros((H, W, 3), np.float32)
        params = capi.dt_render_params(seed, tile_rank, tile_world, max_wave_rays, flags)
        stats = capi.dt_stats()
        rc = self.lib.dt_render(self.handle, C.byref(cam), C.byref(params), ldr.ctypes.data_as(C.c_void_p),
                                hdr.ctypes.data_as(C.c_void_p) if hdr is not None else None, C.byref(stats))
        if rc != 0:
            self._err("dt_render", rc)
        return ldr, hdr, stats

    def render_device(self, cam, seed=1234, tile_rank=0, tile_world=1, flags=0, max_wave_rays=0):
        params = capi.dt_render_params(seed, tile_rank, tile_world, max_wave_rays, flags)
        stats = capi.dt_stats()
        ptr = C.c_void_p()
        rc = self.lib.dt_render_device(self.handle, C.byref(cam), C.byref(params), C.byref(ptr), C.byref(stats))
        if rc != 0:
            self._err("dt_render_device", rc)
        return ptr.value, stats

    def finish_device(self, cam, hdr_dev_ptr, ldr=None):
        if ldr is None:
            ldr = np.zeros((cam.height, cam.width, 3), np.uint8)
        stats = capi.dt_stats()
        rc = self.lib.dt_finish_device(self.handle, C.byref(cam), C.c_void_p(hdr_dev_ptr), ldr.ctypes.data_as(C.c_void_p), C.byref(stats))
        if rc != 0:
            self._err("dt_finish_device", rc)
        return ldr, stats

    # ---- multi-GPU gather over peer memory (dt_frame_*): handles are plain bytes, ship them with any transport
    def frame_export(self, width, height):
        h = capi.dt_frame_handle()
        rc = self.lib.dt_frame_export(self.handle, int(width), int(height), C.byref(h))
        if rc != 0:
            self._err("dt_frame_export", rc)
        return bytes(bytearray(h))

    def frame_import(self, handle_bytes):
        h = capi.dt_frame_handle.from_buffer_copy(handle_bytes)
        rc = self.lib.dt_frame_import(self.handle, C.byref(h))
        if rc != 0:
            self._err("dt_frame_import", rc)

    def frame_release(self):
        self.lib.dt_frame_release(self.handle)

    def frame_finish(self, cam, ldr=None, on_device=False):
        """on_device: tonemap / clamp only, the finished LDR frame stays in device memory (no D2H)."""
        if ldr is None and not on_device:
            ldr = np.zeros((cam.height, cam.width, 3), np.uint8)
        stats = capi.dt_stats()
        rc = self.lib.dt_frame_finish(self.handle, C.byref(cam), None if on_device else ldr.ctypes.data_as(C.c_void_p), C.byref(stats))
        if rc != 0:
            self._err("dt_frame_finish", rc)
        return ldr, stats

    def accel_checksum(self):
        """Ten integers identifying the device-resident acceleration arrays (dt_scene_accel_checksum)."""
        out = (C.c_uint64 * 10)()
        rc = self.lib.dt_scene_accel_checksum(self.handle, C.byref(out))
        if rc != 0:
            self._err("dt_scene_accel_checksum", rc)
        return tuple(int(x) for x in out)

    def primary_hits(self, cam):
        n = cam.width * cam.height
        shape = np.empty(n, np.int32); face = np.empty(n, np.int32); t = np.empty(n, np.float32)
        rc = self.lib.dt_primary_hits(self.handle, C.byref(cam), shape.ctypes.data_as(C.c_void_p),
                                      face.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
        if rc != 0:
            self._err("dt_primary_hits", rc)
        return shape, face, t

    def trace_closest(self, origins, dirs):
        origins = np.ascontiguousarray(origins, np.float32); dirs = np.ascontiguousarray(dirs, np.float32)
        n = origins.shape[0]
        shape = np.empty(n, np.int32); face = np.empty(n, np.int32); t = np.empty(n, np.float32)
        rc = self.lib.dt_trace_closest(self.handle, origins.ctypes.data_as(C.c_void_p), dirs.ctypes.data_as(C.c_void_p), n,
                                       shape.ctypes.data_as(C.c_void_p), face.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
        if rc != 0:
            self._err("dt_trace_closest", rc)
        return shape, face, t

    def trace_occluded(self, origins, dirs, tmax):
        origins = np.ascontiguousarray(origins, np.float32); dirs = np.ascontiguousarray(dirs, np.float32)
        tmax = np.ascontiguousarray(tmax, np.float32)
        n = origins.shape[0]
        occ = np.empty(n, np.uint8)
        rc = self.lib.dt_trace_occluded(self.handle, origins.ctypes.data_as(C.c_void_p), dirs.ctypes.data_as(C.c_void_p),
                                        tmax.ctypes.data_as(C.c_void_p), n, occ.ctypes.data_as(C.c_void_p))
        if rc != 0:
            self._err("dt_trace_occluded", rc)
        return occ

    @property
    def stream_ptr(self):
        return self.lib.dt_scene_stream(self.handle)

    def close(self):
        if self.handle:
            self.lib.dt_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuMulti:
    """dt_multi handle: one process, the first n GPUs of the box (include/dorktracer.h)."""

    def __init__(self, host_scene, n_devices=0):
        self.lib = capi.load_dorktracer()
        self.host = host_scene
        self.handle = C.c_void_p()
        rc = self.lib.dt_multi_create(host_scene.desc_ptr, int(n_devices), C.byref(self.handle))
        if rc != 0:
            raise RuntimeError("dt_multi_create failed (%d): %s" % (rc, self.lib.dt_last_error().decode()))

    @property
    def n_devices(self):
        return self.lib.dt_multi_device_count(self.handle)

    def render(self, cam, seed=1234, flags=0, max_wave_rays=0, want_hdr=True):
        W, H = cam.width, cam.height
        ldr = np.zeros((H, W, 3), np.uint8)
        hdr = np.zeros((H, W, 3), np.float32) if want_hdr else None
        params = capi.dt_render_params(seed, 0, 1, max_wave_rays, flags)
        stats = capi.dt_stats()
        rc = self.lib.dt_multi_render(self.handle, C.byref(cam), C.byref(params), ldr.ctypes.data_as(C.c_void_p),
                                      hdr.ctypes.data_as(C.c_void_p) if hdr is not None else None, C.byref(stats))
        if rc != 0:
            raise RuntimeError("dt_multi_render failed (%d): %s" % (rc, self.lib.dt_last_error().decode()))
        return ldr, hdr, stats

    def close(self):
        if self.handle:
            self.lib.dt_multi_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gpu_tonemap(hdr, key, burn, saturation, gamma):
    lib = capi.load_dorktracer()
    hdr = np.ascontiguousarray(hdr, np.float32)
    H, W = hdr.shape[:2]
    ldr = np.zeros((H, W, 3), np.uint8)
    rc = lib.dt_tonemap(hdr.ctypes.data_as(C.c_void_p), W, H, key, burn, saturation, gamma, ldr.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("dt_tonemap failed (%d): %s" % (rc, lib.dt_last_error().decode()))
    return ldr
