"""Per-warp timeline of the traversal launches of one config-2 frame (GPU box, debug library variant `tl`)."""
import sys, os, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R + '/tests'); sys.path.insert(0, R + '/advanced-cpu-raytracing_b200')
import numpy as np
from dtb200 import capi
real = os.path.join(capi.PKG_DIR, 'libdorktracer.so')
capi._libs[real] = C.CDLL(os.path.join(capi.PKG_DIR, 'libdorktracer_tl.so'), mode=C.RTLD_GLOBAL)
from dtb200.scene import GpuScene, HostScene
from dtb200 import scenegen
lib = capi.load_dorktracer()
lib.dt_debug_timeline.argtypes = [C.c_void_p, C.c_int]; lib.dt_debug_timeline.restype = C.c_int
hs = HostScene(scenegen.gen_config2('/tmp/gen/c2')); cam = hs.camera(0)
gs = GpuScene(hs)
for _ in range(3): gs.render(cam, want_hdr=False)
lib.dt_debug_timeline(None, 0)
hist = (C.c_uint * 128)(); lib.dt_debug_steps_hist(hist, 1)
ldr, hdr, st = gs.render(cam, want_hdr=False)
buf = np.zeros((1 << 20, 4), dtype=np.uint64)
n = lib.dt_debug_timeline(buf.ctypes.data, 1 << 20)
rec = buf[:n].astype(np.int64)
print('frame ms_total %.3f, %d warp records' % (st.ms_total, n))
t_frame0 = rec[:, 1].min()
keys = sorted(set(rec[:, 0].tolist()), key=lambda k: rec[rec[:, 0] == k][:, 1].min())
print('kind      rays   start_us  dur_us | warp exit percentiles (us after launch start): p50 p90 p99 max | drained p50 | busy-warp fraction')
for k in keys:
    r = rec[rec[:, 0] == k]
    t0 = r[:, 1].min(); t1 = r[:, 3].max()
    ex = np.sort(r[:, 3] - t0) / 1e3
    dr = r[:, 2][r[:, 2] > 0]
    busy = (r[:, 3] - r[:, 1]).sum() / max(1, len(r) * (t1 - t0))
    print('%-7s %8d %9.1f %7.1f | %7.1f %7.1f %7.1f %7.1f | %7.1f | %.2f (%d warps)' % (
        'shadow' if (k >> 32) else 'closest', k & 0xFFFFFFFF, (t0 - t_frame0) / 1e3, (t1 - t0) / 1e3,
        ex[len(ex) // 2], ex[int(len(ex) * 0.9)], ex[int(len(ex) * 0.99)], ex[-1],
        (np.median(dr) - t0) / 1e3 if len(dr) else -1, busy, len(r)))
lib.dt_debug_steps_hist(hist, 1)
h = np.array(list(hist)).reshape(2, 64)
for a, nm in ((0, 'closest'), (1, 'shadow')):
    tot = h[a].sum(); cum = np.cumsum(h[a]) / max(1, tot)
    print(nm, 'rays', tot, 'steps/8 histogram (bucket:count):', ' '.join('%d:%d' % (i, c) for i, c in enumerate(h[a]) if c))
gs.close()
