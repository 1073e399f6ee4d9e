import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R + '/tests'); sys.path.insert(0, R + '/advanced-cpu-raytracing_b200')
import numpy as np
from dtb200.scene import GpuScene
from scenes_util import golden_scene
name = sys.argv[1] if len(sys.argv) > 1 else 'cornellbox_recursive_conductors'
hs, g = golden_scene(name); cam = hs.camera(0); gs = GpuScene(hs)
s, f, t = gs.primary_hits(cam)
rs, rf, rt = g["hit_shape"].astype(np.int32), g["hit_face"], g["hit_t"]
bad = np.nonzero((s != rs) | (f != rf) | (t.view(np.uint32) != rt.view(np.uint32)))[0]
print('bad rays', bad[:10], 'W', cam.width, 'H', cam.height)
for i in bad[:5]:
    print('ray', i, 'x', i % cam.width, 'y', i // cam.width, 'got', s[i], f[i], repr(t[i]), 'want', rs[i], rf[i], repr(rt[i]))
    for j in (i - 1, i + 1, i - cam.width, i + cam.width):
        print('   nb', j, s[j], f[j], repr(t[j]), '| ref', rs[j], rf[j], repr(rt[j]))
d = hs.desc
print('n_shapes', d.n_shapes, 'n_meshes', d.n_meshes, [d.meshes[k].n_faces for k in range(d.n_meshes)])
