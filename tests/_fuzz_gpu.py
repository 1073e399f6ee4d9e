"""One-off sweep on the GPU box (not a test): random deterministic scenes beyond the committed seeds, GPU against the oracle.
usage: python tests/_fuzz_gpu.py FIRST LAST"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R + "/tests"); sys.path.insert(0, R + "/advanced-cpu-raytracing_b200")
import numpy as np
from dtb200.scene import GpuScene, HostScene
from oracle_util import ldr_mismatch_fraction, oracle_primary_hits, oracle_render
from scenes_util import random_scene

first, last = int(sys.argv[1]), int(sys.argv[2])
bad = []
for seed in range(first, last):
    p = random_scene("/tmp/rnds", seed, textures=seed % 2 == 1, extras=seed % 4 >= 2)
    hs = HostScene(p)
    cam = hs.camera(0)
    gs = GpuScene(hs)
    s, f, t = gs.primary_hits(cam)
    rs, rf, rt = oracle_primary_hits(hs, cam)
    nh = int(((s != rs) | (f != rf) | (t.view(np.uint32) != rt.view(np.uint32))).sum())
    ldr, hdr, st = gs.render(cam)
    gs.close()
    oldr, ohdr, ost = oracle_render(hs, cam)
    frac, mx = ldr_mismatch_fraction(ldr, oldr, 1)
    rays = (int(st.rays_closest), int(st.rays_shadow)) == (int(ost.rays_closest), int(ost.rays_shadow))
    if nh or frac > 1e-3 or not rays:
        bad.append(seed)
        print(seed, "hits", nh, "ldr", frac, mx, "rays", (int(st.rays_closest), int(st.rays_shadow)), (int(ost.rays_closest), int(ost.rays_shadow)), flush=True)
print("swept", first, last, "bad", bad)
