"""A/B timing of library variants / env knobs on config 2 (GPU box only; not a test).
usage: python tests/_perf_ab.py "lib=NAME,ENV=VAL,..." ...   (lib defaults to the product library)"""
import sys, os, subprocess, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, time, ctypes as C, hashlib
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
sc = int(os.environ.get('DT_AB_SCALE', '1')); cam.width *= sc; cam.height *= sc
gs = GpuScene(hs)
for _ in range(3): gs.render(cam)
n = int(os.environ.get('DT_AB_N', '10')); acc = np.zeros(5); wall = 0
for _ in range(n):
    t0 = time.perf_counter(); ldr, hdr, st = gs.render(cam, want_hdr=False); wall += time.perf_counter() - t0
    acc += np.array([st.ms_total, st.ms_traverse_closest, st.ms_shade, st.ms_traverse_shadow, st.kernel_launches])
acc /= n
print('total %%.3f wall %%.3f closest %%.3f shade %%.3f shadow %%.3f launches %%d | rays %%d+%%d | %%.0f Mrays/s | md5 %%s' %% (
    acc[0], 1e3 * wall / n, acc[1], acc[2], acc[3], acc[4], st.rays_closest, st.rays_shadow,
    (st.rays_closest + st.rays_shadow) / acc[0] / 1e3, hashlib.md5(ldr.tobytes()).hexdigest()[:8]), flush=True)
gs.close()
''' % R
for spec in sys.argv[1:] or ['']:
    env = dict(os.environ)
    for kv in filter(None, spec.split(',')):
        k, v = kv.split('=')
        env['DT_AB_LIB' if k == 'lib' else k] = v
    out = subprocess.run([sys.executable, '-c', CHILD], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600).stdout.decode()
    lines = out.strip().splitlines()
    dbg = [l for l in lines if l.startswith('[dt]')]
    print('[%s] %s %s' % (spec, lines[-1] if lines else 'NO OUTPUT', dbg[-1] if dbg else ''), flush=True)
