"""Checker-side helpers: the C oracle (oracle/libdtoracle.so) and the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY — nothing under advanced-cpu-raytracing_b200/ imports this.
"""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "advanced-cpu-raytracing_b200"))

from dtb200 import capi  # noqa: E402

DTORACLE_SYMBOLS = ["dto_render", "dto_render_reference_rng", "dto_debug_reference_rng", "dto_primary_hits", "dto_tonemap", "dto_trace_closest", "dto_trace_occluded", "dto_set_render_flags"]


def load_dtoracle():
    """CPU restatement of the reference algorithm — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline)."""
    lib = capi._load(os.path.join(REPO, "oracle", "libdtoracle.so"), "oracle library")
    lib.dto_render.argtypes = [C.POINTER(capi.dt_scene_desc), C.POINTER(capi.dt_camera_desc), C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(capi.dt_stats)]
    lib.dto_render.restype = C.c_int
    lib.dto_render_reference_rng.argtypes = [C.POINTER(capi.dt_scene_desc), C.POINTER(capi.dt_camera_desc), C.c_void_p, C.c_void_p, C.POINTER(capi.dt_stats)]
    lib.dto_render_reference_rng.restype = C.c_int
    lib.dto_debug_reference_rng.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    lib.dto_debug_reference_rng.restype = C.c_int
    lib.dto_primary_hits.argtypes = [C.POINTER(capi.dt_scene_desc), C.POINTER(capi.dt_camera_desc), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dto_primary_hits.restype = C.c_int
    lib.dto_tonemap.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
    lib.dto_tonemap.restype = C.c_int
    lib.dto_trace_closest.argtypes = [C.POINTER(capi.dt_scene_desc), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dto_trace_closest.restype = C.c_int
    lib.dto_trace_occluded.argtypes = [C.POINTER(capi.dt_scene_desc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.dto_trace_occluded.restype = C.c_int
    lib.dto_set_render_flags.argtypes = [C.c_int]
    lib.dto_set_render_flags.restype = None
    return lib


REF_DIR = os.path.join(REPO, "oracle", "_ref")
REF_BIN = os.path.join(REF_DIR, "raytracer")
REF_PROBE = os.path.join(REF_DIR, "raytracer_probe")
REF_DROPIN = os.path.join(REF_DIR, "raytracer_dropin")      # the reference's host code with its render loop replaced by libdorktracer.so


def have_ref():
    return os.path.exists(REF_BIN) and os.path.exists(REF_PROBE)


def oracle_render(host_scene, cam, seed=1234, threads=None, want_hdr=True, flags=0):
    lib = load_dtoracle()
    lib.dto_set_render_flags(flags)          # DT_FLAG_SMOOTH_SHADING is the only flag that changes the oracle's image
    W, H = cam.width, cam.height
    ldr = np.zeros((H, W, 3), np.uint8)
    hdr = np.zeros((H, W, 3), np.float32) if want_hdr else None
    stats = capi.dt_stats()
    rc = lib.dto_render(host_scene.desc_ptr, C.byref(cam), seed, threads or (os.cpu_count() or 1),
                        ldr.ctypes.data_as(C.c_void_p), hdr.ctypes.data_as(C.c_void_p) if hdr is not None else None, C.byref(stats))
    lib.dto_set_render_flags(0)
    if rc != 0:
        raise RuntimeError("dto_render failed %d" % rc)
    return ldr, hdr, stats


DTO_FLAG_ORIGIN_LEAK = 1 << 30      # oracle only: reproduce the origin leak of motion-blurred instances (instancedMesh.cpp:22-29)


def oracle_render_reference_rng(host_scene, cam, flags=0):
    """The oracle replaying the reference's own generators on one thread (bit-comparable with DT_THREADS=1 raytracer_probe)."""
    lib = load_dtoracle()
    lib.dto_set_render_flags(flags)
    W, H = cam.width, cam.height
    ldr = np.zeros((H, W, 3), np.uint8)
    hdr = np.zeros((H, W, 3), np.float32)
    stats = capi.dt_stats()
    rc = lib.dto_render_reference_rng(host_scene.desc_ptr, C.byref(cam), ldr.ctypes.data_as(C.c_void_p), hdr.ctypes.data_as(C.c_void_p), C.byref(stats))
    lib.dto_set_render_flags(0)
    if rc != 0:
        raise RuntimeError("dto_render_reference_rng failed %d" % rc)
    return ldr, hdr, stats


def oracle_primary_hits(host_scene, cam):
    lib = load_dtoracle()
    n = cam.width * cam.height
    shape = np.empty(n, np.int32); face = np.empty(n, np.int32); t = np.empty(n, np.float32)
    rc = lib.dto_primary_hits(host_scene.desc_ptr, C.byref(cam), shape.ctypes.data_as(C.c_void_p),
                              face.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("dto_primary_hits failed %d" % rc)
    return shape, face, t


def oracle_trace_closest(host_scene, origins, dirs):
    lib = load_dtoracle()
    origins = np.ascontiguousarray(origins, np.float32); dirs = np.ascontiguousarray(dirs, np.float32)
    n = origins.shape[0]
    shape = np.empty(n, np.int32); face = np.empty(n, np.int32); t = np.empty(n, np.float32)
    rc = lib.dto_trace_closest(host_scene.desc_ptr, origins.ctypes.data_as(C.c_void_p), dirs.ctypes.data_as(C.c_void_p), n,
                               shape.ctypes.data_as(C.c_void_p), face.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return shape, face, t


def oracle_trace_occluded(host_scene, origins, dirs, tmax):
    lib = load_dtoracle()
    origins = np.ascontiguousarray(origins, np.float32); dirs = np.ascontiguousarray(dirs, np.float32)
    tmax = np.ascontiguousarray(tmax, np.float32)
    n = origins.shape[0]
    occ = np.empty(n, np.uint8)
    rc = lib.dto_trace_occluded(host_scene.desc_ptr, origins.ctypes.data_as(C.c_void_p), dirs.ctypes.data_as(C.c_void_p),
                                tmax.ctypes.data_as(C.c_void_p), n, occ.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return occ


def oracle_tonemap(hdr, key, burn, saturation, gamma):
    lib = load_dtoracle()
    hdr = np.ascontiguousarray(hdr, np.float32)
    H, W = hdr.shape[:2]
    ldr = np.zeros((H, W, 3), np.uint8)
    rc = lib.dto_tonemap(hdr.ctypes.data_as(C.c_void_p), W, H, key, burn, saturation, gamma, ldr.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return ldr


def have_dropin():
    return os.path.exists(REF_DROPIN)


def run_reference(xml_path, probe=True, threads=None, cwd=None, timeout=3600, width_height=None, exe=None, extra_env=None):
    """Run the compiled reference on an XML scene.  Returns dict(png, hits, hdr, seconds, closest, shadow).

    The reference resolves plyFile / inputs/<image> against the cwd and writes <ImageName> into it, so it is
    run inside `cwd` (default: the XML's directory if writable, else a temp dir with symlinks)."""
    from PIL import Image
    xml_path = os.path.abspath(xml_path)
    src_dir = os.path.dirname(xml_path)
    tmp = tempfile.mkdtemp(prefix="dt_refrun_")
    # mirror the scene directory (symlinks) so relative assets resolve and outputs land in tmp
    for f in os.listdir(src_dir):
        os.symlink(os.path.join(src_dir, f), os.path.join(tmp, f))
    env = dict(os.environ)
    if threads:
        env["DT_THREADS"] = str(threads)
    hits_p = os.path.join(tmp, "_hits.bin"); hdr_p = os.path.join(tmp, "_hdr.bin")
    if probe:
        env["DT_DUMP_HITS"] = hits_p
        env["DT_DUMP_HDR"] = hdr_p
    if exe is None:
        exe = REF_PROBE if probe else REF_BIN
    if extra_env:
        env.update(extra_env)
    before = set(os.listdir(tmp))
    p = subprocess.run([exe, os.path.basename(xml_path)], cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
    out = p.stdout.decode(errors="replace")
    if p.returncode != 0:
        raise RuntimeError("reference failed (%d): %s" % (p.returncode, out[-2000:]))
    res = {"stdout": out[-4000:]}
    m = re.search(r"Rendering took: ([0-9.eE+-]+)s", out)
    res["seconds"] = float(m.group(1)) if m else None
    m = re.search(r"DT_RAYS closest=(\d+) shadow=(\d+)", out)
    if m:
        res["closest"] = int(m.group(1)); res["shadow"] = int(m.group(2))
    pngs = sorted(f for f in set(os.listdir(tmp)) - before if f.endswith(".png"))
    res["pngs"] = {f: np.array(Image.open(os.path.join(tmp, f)).convert("RGB")) for f in pngs}       # every camera's image, by name
    if pngs:
        res["png"] = np.array(Image.open(os.path.join(tmp, pngs[0])).convert("RGB"))
        H, W = res["png"].shape[:2]
        if probe and os.path.exists(hits_p):
            raw = np.fromfile(hits_p, dtype=np.int32).reshape(-1, 3)
            res["hit_shape"] = raw[:, 0].copy(); res["hit_face"] = raw[:, 1].copy()
            res["hit_t"] = raw[:, 2].copy().view(np.float32)
        if probe and os.path.exists(hdr_p):
            res["hdr"] = np.fromfile(hdr_p, dtype=np.float32).reshape(H, W, 3)
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return res


def ldr_mismatch_fraction(a, b, tol=1):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32)).max(axis=2)
    return float((d > tol).mean()), int(d.max())


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    if mse == 0:
        return 99.0
    return float(10 * np.log10(255.0 ** 2 / mse))


def mc_compare(hdr, ref_hdr, cam, clip=20.0):
    """Two Monte-Carlo radiance frames as estimates of the same image.  The reference itself produces the occasional NaN / inf
    pixel (0/0 in a BRDF lobe at grazing angles; it only prints "nan color!!", raytracer.cpp:128-131) and fireflies of 1e3-1e4
    (unweighted GI, 1/d^2 lights), so: pixels non-finite on either side are excluded and counted, the mean is compared plainly and
    clipped at `clip`, and both frames go through the SAME tonemapper (oracle restatement of tonemapper.h, the camera's parameters)
    with non-finite values zeroed before the LDR PSNR / RMSE is taken."""
    ok = np.isfinite(hdr).all(axis=2) & np.isfinite(ref_hdr).all(axis=2)
    a, b = hdr[ok].astype(np.float64), ref_hdr[ok].astype(np.float64)
    sa, sb = np.where(np.isfinite(hdr), hdr, 0).astype(np.float32), np.where(np.isfinite(ref_hdr), ref_hdr, 0).astype(np.float32)
    la = oracle_tonemap(sa, cam.tm_key, cam.tm_burn, cam.tm_saturation, cam.tm_gamma)
    lb = oracle_tonemap(sb, cam.tm_key, cam.tm_burn, cam.tm_saturation, cam.tm_gamma)
    mse = float(np.mean((la.astype(np.float64) - lb.astype(np.float64)) ** 2))
    # the same without the 2 % of the pixels that differ most: fireflies saturate a few random pixels of EITHER frame
    per_pix = np.sort(((la.astype(np.float64) - lb.astype(np.float64)) ** 2).mean(axis=2).ravel())
    mse_trim = float(per_pix[:max(1, int(per_pix.size * 0.98))].mean())
    return {"finite": float(ok.mean()), "psnr_trim2": 99.0 if mse_trim == 0 else float(10 * np.log10(255.0 ** 2 / mse_trim)), "mean_rel": abs(a.mean() - b.mean()) / b.mean(),
            "clip_rel": abs(np.minimum(a, clip).mean() - np.minimum(b, clip).mean()) / np.minimum(b, clip).mean(),
            "psnr": psnr(la, lb), "rmse": mse ** 0.5}
