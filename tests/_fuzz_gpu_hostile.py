"""One-off sweep on the GPU box (not a test): random deterministic scenes with one hostile edit each (degenerate triangle, zero /
negative sphere radius, zero scale on an axis, objects 1e5 away, light / camera inside a sphere, up parallel to gaze), GPU against
the oracle (which is bit-exact against the compiled reference on all forty).  usage: python tests/_fuzz_gpu_hostile.py"""
import os, re, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R + "/tests"); sys.path.insert(0, R + "/advanced-cpu-raytracing_b200")
import numpy as np
from dtb200.scene import GpuScene, HostScene
from oracle_util import ldr_mismatch_fraction, oracle_primary_hits, oracle_render
from scenes_util import random_scene

rng = np.random.RandomState(11)


def hostile(x, kind):
    if kind == 0:
        m = list(re.finditer(r"(\d+) (\d+) (\d+)\n", x)); mm = m[rng.randint(len(m))]
        return x[:mm.start()] + "%s %s %s\n" % (mm.group(1), mm.group(1), mm.group(3)) + x[mm.end():]
    if kind == 1:
        return re.sub(r"<Radius>[^<]*</Radius>", "<Radius>0</Radius>", x, count=1)
    if kind == 2:
        return re.sub(r'<Scaling id="1">[^<]*</Scaling>', '<Scaling id="1">1 0 1</Scaling>', x)
    if kind == 3:
        return re.sub(r'<Translation id="1">[^<]*</Translation>', '<Translation id="1">100000 0 0</Translation>', x)
    if kind == 4:
        return re.sub(r"<Radius>([^<]*)</Radius>", r"<Radius>-\1</Radius>", x, count=1)
    v = re.search(r"<VertexData>(.*?)</VertexData>", x, re.S).group(1).strip().split("\n")
    c = re.search(r"<Center>(\d+)</Center>", x).group(1)
    if kind == 5:
        return re.sub(r'(<PointLight id="1"><Position>)[^<]*', r"\g<1>" + v[int(c) - 1], x)
    if kind == 6:
        return re.sub(r"(<Camera[^>]*><Position>)[^<]*", r"\g<1>" + v[int(c) - 1], x)
    return re.sub(r"<Up>[^<]*</Up>", "<Up>0 0 -1</Up>", x)


bad = []
for seed in range(40):
    kind = seed % 8
    p0 = random_scene("/tmp/rndh", seed, textures=seed % 2 == 1, extras=seed % 4 >= 2)
    p = "/tmp/rndh/h%d.xml" % seed
    open(p, "w").write(hostile(open(p0).read(), kind))
    hs = HostScene(p)
    cam = hs.camera(0)
    try:
        gs = GpuScene(hs)
    except Exception as e:
        print(seed, kind, "scene rejected:", str(e)[-120:], flush=True); continue
    s, f, t = gs.primary_hits(cam)
    rs, rf, rt = oracle_primary_hits(hs, cam)
    nh = int(((s != rs) | (f != rf) | (t.view(np.uint32) != rt.view(np.uint32))).sum())
    ldr, hdr, st = gs.render(cam)
    gs.close()
    oldr, ohdr, ost = oracle_render(hs, cam)
    frac, mx = ldr_mismatch_fraction(ldr, oldr, 1)
    rays = (int(st.rays_closest), int(st.rays_shadow)) == (int(ost.rays_closest), int(ost.rays_shadow))
    if nh or frac > 1e-3 or not rays:
        bad.append((seed, kind))
        print(seed, kind, "hits", nh, "ldr", frac, mx, "rays", (int(st.rays_closest), int(st.rays_shadow)), (int(ost.rays_closest), int(ost.rays_shadow)), flush=True)
print("hostile sweep: bad", bad)
