"""Render N frames of config 2 (GPU box only; driver for ncu captures, not a test)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R + '/tests'); sys.path.insert(0, R + '/advanced-cpu-raytracing_b200')
import ctypes as C
from dtb200 import capi
if os.environ.get('DT_AB_LIB'):
    capi._libs[os.path.join(capi.PKG_DIR, 'libdorktracer.so')] = C.CDLL(os.path.join(capi.PKG_DIR, 'libdorktracer_%s.so' % os.environ['DT_AB_LIB']), mode=C.RTLD_GLOBAL)
from dtb200.scene import GpuScene, HostScene
from dtb200 import scenegen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
p = scenegen.gen_config2('/tmp/gen/c2'); hs = HostScene(p); cam = hs.camera(0)
gs = GpuScene(hs)
for _ in range(n):
    ldr, hdr, st = gs.render(cam, want_hdr=False)
print('ms_total %.3f rays %d' % (st.ms_total, st.rays_closest + st.rays_shadow))
gs.close()
