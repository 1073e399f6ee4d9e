"""Render one frame of config 3/4/5 (reduced spp) for ncu captures (GPU box only; not a test). argv: c3|c4|c5"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R + '/tests'); sys.path.insert(0, R + '/advanced-cpu-raytracing_b200')
import ctypes as C
from dtb200 import capi
if os.environ.get('DT_AB_LIB'):
    capi._libs[os.path.join(capi.PKG_DIR, 'libdorktracer.so')] = C.CDLL(os.path.join(capi.PKG_DIR, 'libdorktracer_%s.so' % os.environ['DT_AB_LIB']), mode=C.RTLD_GLOBAL)
from dtb200.scene import GpuScene, HostScene
from dtb200 import scenegen
cfg = sys.argv[1]
p = {'c3': lambda: scenegen.gen_config3('/tmp/gen/c3', spp=4), 'c4': lambda: scenegen.gen_config4('/tmp/gen/c4', spp=4), 'c5': lambda: scenegen.gen_config5('/tmp/gen/c5', spp=int(os.environ.get('DT_AB_SPP', '1')))}[cfg]()
hs = HostScene(p); cam = hs.camera(0)
gs = GpuScene(hs)
ldr, hdr, st = gs.render(cam, want_hdr=False, max_wave_rays=int(os.environ.get('DT_AB_WAVE', '0')))
print(cfg, 'ms_total %.1f waves %d rays %d' % (st.ms_total, st.waves, st.rays_closest + st.rays_shadow))
gs.close()
