import sys, time, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R+'/tests'); sys.path.insert(0,R+'/advanced-cpu-raytracing_b200')
import numpy as np
from dtb200.scene import GpuScene, HostScene
from dtb200 import scenegen
p=scenegen.gen_config2('/tmp/gen/c2'); hs=HostScene(p); cam=hs.camera(0)
ref=None
configs=[(0,20),(1,20),(2,16),(2,24),(3,8),(3,16),(3,20),(3,24),(3,28),(3,32)]
if len(sys.argv)>1: configs=[tuple(int(x) for x in a.split(',')) for a in sys.argv[1:]]
for mode,thr in configs:
    os.environ['DT_TRAVERSE_MODE']=str(mode); os.environ['DT_REFILL_THRESHOLD']=str(thr)
    gs=GpuScene(hs)
    for _ in range(3): gs.render(cam)
    acc=np.zeros(6); n=8
    for _ in range(n):
        ldr,hdr,st=gs.render(cam)
        acc+=np.array([st.ms_total,st.ms_generate,st.ms_traverse_closest,st.ms_shade,st.ms_traverse_shadow,st.ms_resolve])
    acc/=n
    if ref is None: ref=ldr
    same=(ldr==ref).all()
    print('mode %d thr %2d: total %.3f closest %.3f shade %.3f shadow %.3f | %.0f Mrays/s | same image: %s'%(mode,thr,acc[0],acc[2],acc[3],acc[4],(st.rays_closest+st.rays_shadow)/acc[0]/1e3,same),flush=True)
    gs.close()
