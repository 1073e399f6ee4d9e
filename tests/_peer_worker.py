"""Rank-1 worker of test_peer_frame_gather_two_processes: imports rank 0's frame buffers (CUDA IPC) and renders its
tiles straight into them (DT_FLAG_PEER_FRAME).  argv: scene-name handle-file width height"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "advanced-cpu-raytracing_b200"))
from dtb200 import capi
from dtb200.scene import GpuScene
from scenes_util import golden_scene

name, hfile, w, h = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
hs, _ = golden_scene(name)
cam = hs.camera(0)
cam.width, cam.height = w, h
gs = GpuScene(hs, device=0)
gs.frame_import(open(hfile, "rb").read())
_, st = gs.render_device(cam, tile_rank=1, tile_world=2, flags=capi.DT_FLAG_PEER_FRAME)
gs.frame_release()
print("worker ok", int(st.rays_closest), int(st.rays_shadow))
