import sys, os, ctypes as C
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R+'/tests'); sys.path.insert(0,R+'/advanced-cpu-raytracing_b200')
import numpy as np
from dtb200 import capi
# load the stats build in place of the product library
path=os.path.join(capi.PKG_DIR,'libdorktracer_stats.so')
real=os.path.join(capi.PKG_DIR,'libdorktracer.so')
capi._libs[real]=C.CDLL(path, mode=C.RTLD_GLOBAL)
from dtb200.scene import GpuScene, HostScene
from dtb200 import scenegen
lib=capi.load_dorktracer()
lib.dt_debug_stats.argtypes=[C.c_void_p,C.c_int]
def stats(reset=1):
    a=(C.c_ulonglong*8)(); lib.dt_debug_stats(a,reset); return np.array(list(a),dtype=np.float64)
p=scenegen.gen_config2('/tmp/gen/c2'); hs=HostScene(p); cam=hs.camera(0)
gs=GpuScene(hs)
stats()
gs.primary_hits(cam); s=stats()
print('PRIMARY rays %d: nodes/ray %.1f tri tests/ray %.1f shapes/ray %.2f blas entries/ray %.2f leaf confirms/ray %.2f steps/ray %.1f'%(s[6],s[0]/s[6],s[1]/s[6],s[2]/s[6],s[3]/s[6],s[4]/s[6],s[5]/s[6]))
ldr,hdr,st=gs.render(cam); s=stats()
print('FRAME rays %d (closest %d shadow %d): nodes/ray %.1f tri tests/ray %.1f shapes/ray %.2f blas entries/ray %.2f leaf confirms/ray %.2f steps/ray %.1f'%(s[6],st.rays_closest,st.rays_shadow,s[0]/s[6],s[1]/s[6],s[2]/s[6],s[3]/s[6],s[4]/s[6],s[5]/s[6]))
