"""Traversal work counters per ray (GPU box, debug library variant `stats` built with -DDT_TRAV_STATS; not a test).
usage: python tests/_trav_stats.py c2|c3|c4"""
import sys, os, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R + '/tests'); sys.path.insert(0, R + '/advanced-cpu-raytracing_b200')
import numpy as np
from dtb200 import capi
real = os.path.join(capi.PKG_DIR, 'libdorktracer.so')
capi._libs[real] = C.CDLL(os.path.join(capi.PKG_DIR, 'libdorktracer_stats.so'), mode=C.RTLD_GLOBAL)
from dtb200.scene import GpuScene, HostScene
from dtb200 import scenegen
lib = capi.load_dorktracer()
lib.dt_debug_stats.argtypes = [C.c_void_p, C.c_int]
def stats(reset=1):
    a = (C.c_ulonglong * 8)(); lib.dt_debug_stats(a, reset); return np.array(list(a), dtype=np.float64)
cfg = sys.argv[1] if len(sys.argv) > 1 else 'c2'
p = {'c2': lambda: scenegen.gen_config2('/tmp/gen/c2'), 'c3': lambda: scenegen.gen_config3('/tmp/gen/c3', spp=1), 'c4': lambda: scenegen.gen_config4('/tmp/gen/c4', spp=1)}[cfg]()
hs = HostScene(p); cam = hs.camera(0)
gs = GpuScene(hs)
stats()
def show(tag, s, extra=''):
    print('%s rays %d: nodes/ray %.2f tri tests/ray %.2f shape visits/ray %.2f blas entries/ray %.2f leaf confirms/ray %.2f steps/ray %.2f %s' % (
        tag, s[6], s[0] / s[6], s[1] / s[6], s[2] / s[6], s[3] / s[6], s[4] / s[6], s[5] / s[6], extra), flush=True)
gs.primary_hits(cam); show(cfg + ' PRIMARY', stats())
ldr, hdr, st = gs.render(cam); show(cfg + ' FRAME', stats(), '(closest %d shadow %d)' % (st.rays_closest, st.rays_shadow))
