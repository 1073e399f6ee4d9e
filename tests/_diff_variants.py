"""Compare the LDR output of library variants against the oracle on config 2 (GPU box only; not a test)."""
import sys, os, subprocess, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, ctypes as C
R = %r
sys.path.insert(0, R + '/tests'); sys.path.insert(0, R + '/advanced-cpu-raytracing_b200')
import numpy as np
from dtb200 import capi
name = os.environ.get('DT_AB_LIB', '')
if name:
    real = os.path.join(capi.PKG_DIR, 'libdorktracer.so')
    capi._libs[real] = C.CDLL(os.path.join(capi.PKG_DIR, 'libdorktracer_%%s.so' %% name), mode=C.RTLD_GLOBAL)
from dtb200.scene import GpuScene, HostScene
from dtb200 import scenegen
p = scenegen.gen_config2('/tmp/gen/c2'); hs = HostScene(p); cam = hs.camera(0)
gs = GpuScene(hs)
ldr, hdr, st = gs.render(cam, want_hdr=True)
np.save(sys.argv[1] + '_ldr.npy', ldr); np.save(sys.argv[1] + '_hdr.npy', hdr)
sh, fc, t = gs.primary_hits(cam)
np.save(sys.argv[1] + '_face.npy', fc); np.save(sys.argv[1] + '_t.npy', t)
gs.close()
''' % R
names = sys.argv[1:] or ['', 'mb6']
for n in names:
    env = dict(os.environ); env['DT_AB_LIB'] = n
    subprocess.run([sys.executable, '-c', CHILD, '/tmp/var_' + (n or 'default')], env=env, check=True)
sys.path.insert(0, R + '/tests'); sys.path.insert(0, R + '/advanced-cpu-raytracing_b200')
from dtb200.scene import HostScene
from dtb200 import scenegen
from oracle_util import oracle_render
hs = HostScene(scenegen.gen_config2('/tmp/gen/c2')); cam = hs.camera(0)
oldr, ohdr, ost = oracle_render(hs, cam, threads=os.cpu_count() or 8, want_hdr=True)
base = None
for n in names:
    k = n or 'default'
    ldr = np.load('/tmp/var_%s_ldr.npy' % k).astype(int); hdr = np.load('/tmp/var_%s_hdr.npy' % k)
    fc = np.load('/tmp/var_%s_face.npy' % k); t = np.load('/tmp/var_%s_t.npy' % k)
    d = np.abs(ldr - oldr.astype(int)).reshape(-1, 3).max(axis=1)
    msg = '%s: vs oracle: %d pixels differ (max %d, >1: %d)' % (k, (d > 0).sum(), d.max(), (d > 1).sum())
    if base is None: base = (ldr, hdr, fc, t)
    else:
        dd = np.abs(ldr - base[0]).reshape(-1, 3).max(axis=1)
        idx = np.nonzero(dd)[0]
        msg += ' | vs %s: %d pixels differ (max %d) first idx %s; primary face diffs %d t diffs %d; hdr max abs diff %.4g' % (
            names[0] or 'default', len(idx), dd.max(), idx[:8].tolist(), (fc != base[2]).sum(), (t != base[3]).sum(), np.abs(hdr - base[1]).max())
    print(msg, flush=True)
