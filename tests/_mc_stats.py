"""GPU estimator against the high-spp reference renders of tests/golden/mc_*.npz (GPU box only; not a test): prints the
numbers the bounds of test_monte_carlo_against_the_high_spp_reference_render are taken from."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R + '/tests'); sys.path.insert(0, R + '/advanced-cpu-raytracing_b200')
import numpy as np
from dtb200.scene import GpuScene, HostScene
from oracle_util import mc_compare
from test_cpu_monte_carlo_pin import mc_scene
for name in (sys.argv[1:] or ["mc_all", "mc_area", "mc_mesh", "mc_env", "mc_c5shape", "mc_rrbrdf"]):
    for spp in (256, 4096, 16384, 65536):
        p, g = mc_scene(name, '/tmp/mcs_%s_%d' % (name, spp), spp)
        hs = HostScene(p); cam = hs.camera(0); gs = GpuScene(hs)
        for seed in (11, 12, 13):
            ldr, hdr, st = gs.render(cam, seed=seed)
            m = mc_compare(hdr, g['hi_hdr'], cam)
            print('%-10s spp %6d seed %d | finite %.4f | mean rel %.4f | clip20 rel %.4f | psnr %.2f dB (rmse %.2f), %.2f dB trimmed | nan_pixels %d | %.0f ms' % (
                name, spp, seed, m['finite'], m['mean_rel'], m['clip_rel'], m['psnr'], m['rmse'], m['psnr_trim2'], st.nan_pixels, st.ms_total), flush=True)
        gs.close()
