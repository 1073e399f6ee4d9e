import sys, time, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R+'/tests'); sys.path.insert(0,R+'/advanced-cpu-raytracing_b200')
import numpy as np
from dtb200.scene import GpuScene, HostScene
from dtb200 import scenegen
from oracle_util import oracle_render, psnr, oracle_primary_hits
t0=time.time(); p=scenegen.gen_config5('/tmp/gen/c5', spp=4); t1=time.time()
hs=HostScene(p); t2=time.time(); gs=GpuScene(hs); t3=time.time()
print('c5: gen %.1fs host load+BVH2 %.1fs scene_create %.1fs tris %d'%(t1-t0,t2-t1,t3-t2,hs.n_triangles()),flush=True)
cam=hs.camera(0)
for rep in range(2):
    ldr,hdr,st=gs.render(cam)
    print('c5 4K 4spp: ms_total %.1f closest %.1f shade %.1f shadow %.1f tonemap %.2f | rays %d+%d waves %d retries %d nan %d -> %.0f Mrays/s'%(st.ms_total,st.ms_traverse_closest,st.ms_shade,st.ms_traverse_shadow,st.ms_tonemap,st.rays_closest,st.rays_shadow,st.waves,st.retries,st.nan_pixels,(st.rays_closest+st.rays_shadow)/st.ms_total/1e3),flush=True)
from PIL import Image
Image.fromarray(ldr[::2,::2]).save(R+'/gpurun_out/c5_4spp_half.png')
# parity at reduced size: primary hits bit-exact vs oracle, MC statistics
cam.width,cam.height,cam.samples_per_pixel=240,136,64
sh,fa,tt=gs.primary_hits(cam); t0=time.time(); os_,of_,ot_=oracle_primary_hits(hs,cam); print('oracle primary hits %.1fs'%(time.time()-t0))
print('c5 240x136 primary hits mismatches:',int(((sh!=os_)|(fa!=of_)|(tt.view(np.uint32)!=ot_.view(np.uint32))).sum()),'of',sh.size,flush=True)
ldr,hdr,st=gs.render(cam,seed=3); t0=time.time(); oldr,ohdr,ost=oracle_render(hs,cam,seed=9); print('oracle render %.1fs'%(time.time()-t0))
print('c5 240x136x64spp: mean radiance gpu %.4f oracle %.4f  PSNR(ldr) %.1f dB'%(hdr.mean(),ohdr.mean(),psnr(ldr,oldr)),flush=True)
