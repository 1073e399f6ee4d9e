"""Scene set-up times with the host and the GPU builders (GPU box only; not a test).
usage: python tests/_setup_times.py c2|c5"""
import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R + '/tests'); sys.path.insert(0, R + '/advanced-cpu-raytracing_b200')
from dtb200.scene import GpuScene, HostScene
from dtb200 import scenegen, capi
cfg = sys.argv[1] if len(sys.argv) > 1 else 'c2'
p = scenegen.gen_config2('/tmp/gen/c2') if cfg == 'c2' else scenegen.gen_config5('/tmp/gen/c5', spp=1)
capi.load_dorktracer().dt_gpu_init(0)
import ctypes as C
w = GpuScene(HostScene(scenegen.gen_config2('/tmp/gen/warm', nlon=40, nlat=19, width=64, height=64))); w.close()   # CUDA context + module load
for rep in range(2):
    for gpu_build in (False, True):
        t0 = time.perf_counter(); hs = HostScene(p, gpu_build=gpu_build); t1 = time.perf_counter()
        for mf in (-1, 32768):
            os.environ['DT_GPU_FLATTEN_MIN_FACES'] = str(mf)
            t2 = time.perf_counter(); gs = GpuScene(hs); t3 = time.perf_counter()
            print('%s %d tris | BVH2 build on %s: load %.3f s (build part %.3f s) | flatten on %s: dt_scene_create %.3f s | checksum %x' % (
                cfg, hs.n_triangles(), 'GPU ' if gpu_build else 'host', t1 - t0, hs.bvh_build_seconds, 'GPU ' if mf > 0 else 'host', t3 - t2, gs.accel_checksum()[1]), flush=True)
            gs.close()
