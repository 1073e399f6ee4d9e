#!/usr/bin/env python3
"""Generate the committed Monte-Carlo fixtures tests/golden/mc_*.npz from the COMPILED REFERENCE (oracle/_ref).

Run in the build container only (needs /root/reference and oracle/_ref built by oracle/build_ref.py):
    python tests/golden/make_golden_mc.py [name ...]

The reference seeds every generator from a default-constructed mt19937 or from an unseeded rand(), so with ONE render thread
(DT_THREADS=1, patch P3 of oracle/build_ref.py) a Monte-Carlo render is deterministic.  Per scene one .npz:
    kwargs     JSON of the scene-generator call (dtb200.scenegen.gen_config4 / gen_config5 / tests/scenes_util.blur_dof_scene)
    xml        the generated scene file (the test regenerates it and checks that the generator has not drifted)
    pin_hdr    float32 HxWx3 radiance of `DT_THREADS=1 raytracer_probe` at 4 spp           } the EXACT pin: the oracle's
    pin_rays   [closest, shadow] ray counts of that run                                     } reference-RNG mode must match bit for bit
    hi_spp     samples per pixel of the high-spp render (4096; 16384 for mc_c5shape)
    hi_hdr     float32 radiance of an 8-thread hi_spp reference render (racy generators: a statistical truth, the
               "high-spp reference render" of BASELINE.json's north_star)
    hi_ldr     the tonemapped LDR image the reference wrote for it
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "advanced-cpu-raytracing_b200"))
from oracle_util import run_reference  # noqa: E402
from dtb200 import scenegen  # noqa: E402
import scenes_util  # noqa: E402

W, H = 48, 32
SCENES = {
    "mc_all": ("config4", dict(lights=["area", "mesh", "env"], depth=2)),
    "mc_area": ("config4", dict(lights=["area"], depth=2, area_light_y=7.5)),     # light 2.5 under the ceiling: finite variance (see gen_config4)
    "mc_mesh": ("config4", dict(lights=["mesh"], depth=2)),
    "mc_env": ("config4", dict(lights=["env"], depth=2)),
    "mc_c5shape": ("config5", dict(nlon=96, nlat=48, depth=2)),          # config-4 scene + a 9 024-triangle displaced sphere, lifted spheres
    "mc_rrbrdf": ("config4", dict(lights=["area", "mesh"], depth=2, area_light_y=7.5, mirror_brdf=True)),   # mirror WITH a BRDF: bounds the RR-throughput deviation
    "mc_blur": ("blur", dict()),                                         # thin lens, motion blur, rough mirror (no path tracing)
}


HI_SPP = {"mc_c5shape": 16384}        # default 4096; the mesh scene's estimator is the heaviest-tailed


def generate(kind, kw, out_dir, spp):
    kw = dict(kw)
    if "lights" in kw:
        kw["lights"] = tuple(kw["lights"])
    if kind == "config4":
        return scenegen.gen_config4(out_dir, width=W, height=H, spp=spp, **kw)
    if kind == "config5":
        return scenegen.gen_config5(out_dir, width=W, height=H, spp=spp, **kw)
    return scenes_util.blur_dof_scene(os.path.join(out_dir, "blur.xml"), width=W, height=H, spp=spp)


def main(names):
    for name in names:
        kind, kw = SCENES[name]
        tmp = tempfile.mkdtemp(prefix="dt_mc_")
        os.makedirs(tmp, exist_ok=True)
        p = generate(kind, kw, tmp, 4)
        pin = run_reference(p, probe=True, threads=1)
        out = {"kwargs": np.frombuffer(json.dumps({"kind": kind, "kw": kw}).encode(), dtype=np.uint8),
               "xml": np.frombuffer(open(p, "rb").read(), dtype=np.uint8),
               "pin_hdr": pin["hdr"].astype(np.float32), "pin_rays": np.array([pin["closest"], pin["shadow"]], dtype=np.int64)}
        if kind != "blur":
            tmp2 = tempfile.mkdtemp(prefix="dt_mc_hi_")
            hi_spp = HI_SPP.get(name, 4096)
            p2 = generate(kind, kw, tmp2, hi_spp)
            hi = run_reference(p2, probe=True, threads=8, timeout=7200)
            out["hi_spp"] = np.array([hi_spp], dtype=np.int64)
            out["hi_hdr"] = hi["hdr"].astype(np.float32)
            out["hi_ldr"] = hi["png"]
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB", "pin rays", out["pin_rays"].tolist(), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(SCENES))
