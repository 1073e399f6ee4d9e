import sys, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R+'/tests'); sys.path.insert(0,R+'/advanced-cpu-raytracing_b200')
import numpy as np
from scenes_util import *
from oracle_util import *
from dtb200.scene import GpuScene
for name in PINS+DIELECTRIC:
    hs,g=golden_scene(name); cam=hs.camera(0)
    gs=GpuScene(hs)
    sh,fa,tt=gs.primary_hits(cam)
    rs=g['hit_shape'].astype(np.int32); rf=g['hit_face']; rt=g['hit_t']
    print(name,'prim hits: shape mism',(sh!=rs).sum(),'face mism',(fa!=rf).sum(),'t bits mism',(tt.view(np.uint32)!=rt.view(np.uint32)).sum(), flush=True)
    ldr,hdr,st=gs.render(cam)
    fr,mx=ldr_mismatch_fraction(ldr,g['ref_ldr'],1)
    fr0,_=ldr_mismatch_fraction(ldr,g['ref_ldr'],0)
    print('   LDR vs ref: frac>1 =',fr,'max',mx,'frac>0',fr0,' rays',st.rays_closest,st.rays_shadow,'ref',g['rays'],'ms',st.ms_total,'waves',st.waves, flush=True)
    if 'golden' in g.files:
        print('   LDR vs course golden frac>1:',ldr_mismatch_fraction(ldr,g['golden'],1), flush=True)
