import sys, os, time, ctypes as C
R = '/root/repo'
sys.path.insert(0, R + '/tests'); sys.path.insert(0, R + '/advanced-cpu-raytracing_b200')
import numpy as np
from dtb200.scene import HostScene, GpuScene
from dtb200 import scenegen, capi
p = scenegen.gen_config5('/tmp/gen/c5', spp=1)
capi.load_dorktracer().dt_gpu_init(0)
t0 = time.time(); host = HostScene(p); t1 = time.time(); gpu = HostScene(p, gpu_build=True); t2 = time.time()
def arrays(hs):
    out = []
    d = hs.desc
    for i in range(d.n_meshes):
        m = d.meshes[i]
        faces = np.frombuffer(C.string_at(m.faces, m.n_faces * C.sizeof(capi.dt_face)), np.uint8)
        nodes = np.frombuffer(C.string_at(m.bvh, m.n_bvh_nodes * C.sizeof(capi.dt_bvh2_node)), np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("left", "<i4"), ("right", "<i4"), ("first", "<u4"), ("count", "<u4")]))
        out.append((faces, nodes))
    return out
ok = True
for (fa, na), (fb, nb) in zip(arrays(host), arrays(gpu)):
    ok &= fa.shape == fb.shape and bool((fa == fb).all()) and na.shape == nb.shape
    for k in ("left", "right", "first", "count", "bmin", "bmax"):
        ok &= bool((na[k] == nb[k]).all())
    print('mesh: %d faces, %d BVH2 nodes' % (fa.size // C.sizeof(capi.dt_face), na.size))
print('config 5 (%d triangles): GPU-built BVH2 identical to the host build: %s | load host %.2f s (build %.2f) / GPU %.2f s (build %.2f)' % (host.n_triangles(), ok, t1 - t0, host.bvh_build_seconds, t2 - t1, gpu.bvh_build_seconds))
os.environ['DT_GPU_FLATTEN_MIN_FACES'] = '-1'; a = GpuScene(host); ca = a.accel_checksum(); a.close()
os.environ['DT_GPU_FLATTEN_MIN_FACES'] = '1'; b = GpuScene(gpu); cb = b.accel_checksum(); b.close()
print('BVH8 / triangle / leaf-box / face-map checksums, host flattener vs GPU flattener on the GPU-built tree: %s (%d nodes, %d primitives)' % (ca == cb, cb[8], cb[9]))
