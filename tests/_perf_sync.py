import sys, time, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R+'/tests'); sys.path.insert(0,R+'/advanced-cpu-raytracing_b200')
import numpy as np
from dtb200.scene import GpuScene, HostScene
from dtb200 import scenegen
p=scenegen.gen_config2('/tmp/gen/c2'); hs=HostScene(p); cam=hs.camera(0)
ref=None
for sync,mode,thr in [(1,2,16),(0,2,16),(0,3,16),(0,0,16),(0,1,16)]:
    os.environ['DT_SYNC_WAVES']=str(sync); os.environ['DT_TRAVERSE_MODE']=str(mode); os.environ['DT_REFILL_THRESHOLD']=str(thr)
    gs=GpuScene(hs)
    for _ in range(3): gs.render(cam)
    acc=np.zeros(6); n=10; wall=0
    for _ in range(n):
        t0=time.perf_counter(); ldr,hdr,st=gs.render(cam,want_hdr=False); wall+=time.perf_counter()-t0
        acc+=np.array([st.ms_total,st.ms_generate,st.ms_traverse_closest,st.ms_shade,st.ms_traverse_shadow,st.ms_resolve])
    acc/=n
    if ref is None: ref=ldr
    d=np.abs(ldr.astype(int)-ref.astype(int)).max()
    print('sync %d mode %d: total %.3f (wall %.3f) closest %.3f shade %.3f shadow %.3f | rays %d+%d | %.0f Mrays/s | max LDR diff vs first %d'%(sync,mode,acc[0],1e3*wall/n,acc[2],acc[3],acc[4],st.rays_closest,st.rays_shadow,(st.rays_closest+st.rays_shadow)/acc[0]/1e3,d),flush=True)
    gs.close()
