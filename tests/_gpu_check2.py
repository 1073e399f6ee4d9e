import sys, time, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R+'/tests'); sys.path.insert(0,R+'/advanced-cpu-raytracing_b200')
import numpy as np
from scenes_util import *
from oracle_util import *
from dtb200.scene import GpuScene, HostScene
from dtb200 import scenegen
def report(tag, hs, cam, gs, ref_ldr=None, ref_hits=None, reps=1):
    sh,fa,tt=gs.primary_hits(cam)
    if ref_hits is None: ref_hits=oracle_primary_hits(hs,cam)
    rs,rf,rt=ref_hits
    print(tag,'prim hits: shape mism',(sh!=rs).sum(),'face mism',(fa!=rf).sum(),'t bits mism',(tt.view(np.uint32)!=rt.view(np.uint32)).sum(),'of',sh.size, flush=True)
    for r in range(reps):
        ldr,hdr,st=gs.render(cam)
        print('   render: ms_total %.3f gen %.3f closest %.3f shade %.3f shadow %.3f resolve %.3f | rays %d+%d  waves %d launches %d  -> %.1f Mrays/s'%(st.ms_total,st.ms_generate,st.ms_traverse_closest,st.ms_shade,st.ms_traverse_shadow,st.ms_resolve,st.rays_closest,st.rays_shadow,st.waves,st.kernel_launches,(st.rays_closest+st.rays_shadow)/st.ms_total/1e3), flush=True)
    if ref_ldr is None:
        t0=time.time(); ref_ldr,_,ost=oracle_render(hs,cam); print('   oracle render %.1fs rays %d+%d'%(time.time()-t0,ost.rays_closest,ost.rays_shadow))
    fr,mx=ldr_mismatch_fraction(ldr,ref_ldr,1); fr0,_=ldr_mismatch_fraction(ldr,ref_ldr,0)
    print('   LDR vs oracle/ref: frac>1 = %.3g max %d frac>0 %.3g'%(fr,mx,fr0), flush=True)
    return ldr
for name in ['cornellbox_recursive_conductors','spheres_mirror','scienceTree_diamond']:
    hs,g=golden_scene(name); cam=hs.camera(0); gs=GpuScene(hs)
    report(name,hs,cam,gs,g['ref_ldr'],(g['hit_shape'].astype(np.int32),g['hit_face'],g['hit_t']),reps=2)
d='/tmp/gen'
p=scenegen.gen_config2(d+'/c2s', nlon=80, nlat=41, width=320, height=184); hs=HostScene(p); report('c2-small',hs,hs.camera(0),GpuScene(hs))
p=scenegen.gen_config3(d+'/c3s', grid=6, base_nlon=24, base_nlat=13, width=320, height=184, spp=4); hs=HostScene(p); report('c3-small',hs,hs.camera(0),GpuScene(hs))
p=scenegen.gen_config3(d+'/c3m', grid=32, base_nlon=64, base_nlat=33, width=640, height=360, spp=4); hs=HostScene(p); report('c3-medium',hs,hs.camera(0),GpuScene(hs))
t0=time.time(); p=scenegen.gen_config2(d+'/c2'); hs=HostScene(p); t1=time.time(); gs=GpuScene(hs); t2=time.time()
print('c2 full: gen+load %.1fs, scene_create %.1fs, tris %d'%(t1-t0,t2-t1,hs.n_triangles()))
ldr=report('c2-full',hs,hs.camera(0),gs,reps=3)
from PIL import Image
Image.fromarray(ldr).save(R+'/gpurun_out/c2_full.png')
