"""Per-launch intervals of one config-2 frame (DT_DEBUG_TIMING=2; GPU box only; not a test)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R + '/tests'); sys.path.insert(0, R + '/advanced-cpu-raytracing_b200')
os.environ.setdefault('DT_DEBUG_TIMING', '0')
from dtb200.scene import GpuScene, HostScene
from dtb200 import scenegen
p = scenegen.gen_config2('/tmp/gen/c2'); hs = HostScene(p); cam = hs.camera(0)
gs = GpuScene(hs)
for _ in range(4): gs.render(cam, want_hdr=False)
gs.close()
