"""One-off sweep on the GPU box (not a test): random STOCHASTIC scenes (scenes_util.random_scene(mc=True)), every seed including
the ones the reference cannot render; the GPU path must return a frame for each, the same ray tree for the same seed.
usage: python tests/_fuzz_gpu_mc.py FIRST LAST [SPP]"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R + "/tests"); sys.path.insert(0, R + "/advanced-cpu-raytracing_b200")
import numpy as np
from dtb200.scene import GpuScene, HostScene
from scenes_util import random_scene

first, last = int(sys.argv[1]), int(sys.argv[2])
spp = int(sys.argv[3]) if len(sys.argv) > 3 else None
slow, bad = [], []
for seed in range(first, last):
    p = random_scene("/tmp/rndmc", seed, width=112, height=80, textures=(seed % 3 == 1), extras=(seed % 2 == 1), mc=True, spp=spp)
    hs = HostScene(p)
    cam = hs.camera(0)
    gs = GpuScene(hs)
    t0 = time.time()
    try:
        ldr, hdr, st = gs.render(cam, seed=3)
        ldr2, hdr2, st2 = gs.render(cam, seed=3)
    except Exception as e:
        bad.append(seed); print(seed, "ERROR", str(e)[:200], flush=True); gs.close(); continue
    dt = time.time() - t0
    gs.close()
    same = (int(st.rays_closest), int(st.rays_shadow)) == (int(st2.rays_closest), int(st2.rays_shadow))
    nanfrac = float(np.isnan(hdr).any(axis=2).mean())
    if dt > 4.0: slow.append(seed)
    if not same: bad.append(seed)
    print(seed, "spp", cam.samples_per_pixel, "pt" if cam.path_tracing else "", "rr" if cam.russian_roulette else "", "rays", int(st.rays_closest), int(st.rays_shadow), "same" if same else "DIFFERENT TREE",
          "retries", int(st.retries), "nan %.2f" % nanfrac, "%.2fs" % dt, flush=True)
print("swept", first, last, "slow", slow, "bad", bad)
