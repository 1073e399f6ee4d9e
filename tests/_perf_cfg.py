"""Timing of configs 3/4/5 at (reduced) bench sizes with env knobs (GPU box only; not a test).
usage: python tests/_perf_cfg.py c4 "DT_SORT=0" "DT_SORT=2" ..."""
import sys, os, subprocess
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, time, hashlib
R = %r
sys.path.insert(0, R + '/tests'); sys.path.insert(0, R + '/advanced-cpu-raytracing_b200')
import numpy as np
import ctypes as C
from dtb200 import capi
name = os.environ.get('DT_AB_LIB', '')
if name:
    real = os.path.join(capi.PKG_DIR, 'libdorktracer.so')
    capi._libs[real] = C.CDLL(os.path.join(capi.PKG_DIR, 'libdorktracer_%%s.so' %% name), mode=C.RTLD_GLOBAL)
from dtb200.scene import GpuScene, HostScene
from dtb200 import scenegen
cfg = sys.argv[1]
t0 = time.time()
if cfg == 'c3': p = scenegen.gen_config3('/tmp/gen/c3')
elif cfg == 'c4': p = scenegen.gen_config4('/tmp/gen/c4', spp=int(os.environ.get('DT_AB_SPP', '16')))
elif cfg == 'c5': p = scenegen.gen_config5('/tmp/gen/c5', spp=int(os.environ.get('DT_AB_SPP', '4')), sphere_lift=float(os.environ.get('DT_AB_LIFT', '0.25')))
else: raise SystemExit('unknown config')
hs = HostScene(p); cam = hs.camera(0); t1 = time.time()
gs = GpuScene(hs); t2 = time.time()
WAVE = int(os.environ.get('DT_AB_WAVE', '0')); FLAGS = int(os.environ.get('DT_AB_FLAGS', '0'))
if os.environ.get('DT_AB_WARM', '1') == '1': gs.render(cam, want_hdr=False, max_wave_rays=WAVE, flags=FLAGS)
n = int(os.environ.get('DT_AB_N', '3')); acc = np.zeros(7)
for _ in range(n):
    ldr, hdr, st = gs.render(cam, want_hdr=False, max_wave_rays=WAVE, flags=FLAGS)
    acc += np.array([st.ms_total, st.ms_traverse_closest, st.ms_shade, st.ms_traverse_shadow, st.ms_sort, st.waves, st.kernel_launches])
acc /= n
print('%%s %%dx%%d spp %%d tris %%d (load %%.1fs create %%.1fs): total %%.1f closest %%.1f shade %%.1f shadow %%.1f sort %%.1f ms | waves %%d launches %%d | rays %%d+%%d | %%.0f Mrays/s | md5 %%s' %% (
    cfg, cam.width, cam.height, cam.samples_per_pixel, hs.n_triangles(), t1 - t0, t2 - t1, acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6],
    st.rays_closest, st.rays_shadow, (st.rays_closest + st.rays_shadow) / acc[0] / 1e3, hashlib.md5(ldr.tobytes()).hexdigest()[:8]), flush=True)
gs.close()
''' % R
cfg = sys.argv[1]
for spec in sys.argv[2:] or ['']:
    env = dict(os.environ)
    for kv in filter(None, spec.split(',')):
        k, v = kv.split('=')
        env[k] = v
    out = subprocess.run([sys.executable, '-c', CHILD, cfg], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=3000).stdout.decode()
    lines = out.strip().splitlines()
    for l in lines[:-1]:
        if l.startswith('[dt'): print('    ' + l)
    print('[%s] %s' % (spec, lines[-1] if lines else 'NO OUTPUT'), flush=True)
