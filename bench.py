#!/usr/bin/env python3
"""bench.py — headline benchmark of the render hot path (contract: see the task's bench section).

Metric (BASELINE.json): Mrays/s (primary + secondary + shadow) for the whole job, frame ms.
A *step* is one frame of the workload through the hot path.

  --config 5 (default, the configuration BASELINE.json quotes the metric on): the synthetic procedural 10 M-triangle scene
        (9 991 932 triangles: config-4 box + displaced-sphere mesh), path tracing with next-event estimation + importance
        sampling + Russian roulette, area light + mesh light + spherical HDR environment light, Torrance-Sparrow / modified
        Blinn-Phong BRDFs, photographic tonemapper, 3840x2160 — at a STATED REDUCED sample count (--spp, default 64 instead of
        1024: a 1024-spp frame is 55 s on one B200, a 64-spp frame 3.5 s).  N > 1: THE SAME FRAME is sharded over the ranks (strips of eight
        8x4-pixel tiles, round-robin; what the reference does with row bands over its threads, main.cpp:38-39) -> strong
        scaling.  Every rank's resolve kernel stores its strips of radiance straight into rank 0's frame over NVLink (CUDA IPC
        peer memory, no collective), one barrier, then rank 0 tonemaps the whole frame.
  --config 2: ~1 M-triangle mesh, mirror + dielectric recursion depth 6, 2 point lights, 1920x1080, 1 spp (the round-1
        headline; weak scaling: sqrt(N) x the resolution per axis).  Its N=1 numbers also ride along in `other_workloads`.

  value      device time only: scene + camera resident, CUDA events on the library's stream (render + resolve + tonemap,
             + the barrier for N > 1), L2 flushed between iterations.  Rays = rays TRACED: path-traced frames do not follow paths
             whose weight is exactly zero nor trace shadow rays whose contribution is exactly zero (same image; dorktracer.h,
             DT_FLAG_KEEP_WEIGHTLESS_PATHS); the size of the reference's ray tree for the same frame (one untimed frame with that
             flag) and the rate in that accounting are reported in details.ray_accounting
  e2e        the public C-ABI call dt_render() with a pinned HOST LDR buffer: camera / parameters in, D2H of the finished
             frame inside the timed region, wall clock between device synchronisations
  roofline   the dominant kernel (the traversal kernel with the larger share of the frame).  Primary bound = SM issue slots
             (the binding roof: warp instructions per ray from the committed ncu capture x rays / live CUDA-event time of the
             kernel's launches, against 4 issue slots x #SM x SM clock); secondary = HBM (SURVEY.md 8d algorithmic bytes per
             ray x rays / the same live time, against MEASURED_PEAKS.json).  Live times come from a separate
             DT_FLAG_SERIAL_WAVES pass inside this run (each kernel alone on the GPU).
  cpu_baseline / --impl reference   the compiled reference (oracle/_ref/raytracer) on this box's host cores, SAME scene, on a
             bounded sample of the workload (reduced resolution and samples: the reference manages ~1-3 Mrays/s here).
"""
import argparse
import ctypes as C
import json
import math
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "advanced-cpu-raytracing_b200"))

METRIC = "Mrays/s (primary+secondary+shadow), whole job"
UNIT = "Mrays/s"

# SURVEY.md 8(d) algorithmic bytes per ray: 32 in + 32 out + 80 B x nodes on one root-to-leaf chain + 4 x 48 B triangles;
# shadow rays the same minus the 32 B hit record plus a 4 B flag.
WORKLOADS = {
    5: {"b_closest": 896.0, "b_shadow": 868.0, "width": 3840, "height": 2160, "spp": 64, "scaling": "strong",
        "ref_sample": (480, 272, 4)},      # reference arm: same scene, 480x272, 4 spp (about 12 M rays per step)
    2: {"b_closest": 736.0, "b_shadow": 708.0, "width": 1920, "height": 1080, "spp": 1, "scaling": "weak",
        "ref_sample": (1920, 1080, 1)},
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def make_workload(cfg, tmp, width, height, spp, n_cameras=1):
    """Generates the scene files (seeded, procedural) and returns the XML path."""
    from dtb200 import scenegen
    if cfg == 2:
        return scenegen.gen_config2(os.path.join(tmp, "c2"), nlon=1000, nlat=499, width=width, height=height, depth=6)
    return scenegen.gen_config5(os.path.join(tmp, "c5"), width=width, height=height, spp=spp, n_cameras=n_cameras)


def workload_config(cfg, spp, n_tris=None):
    """The `config` object of the JSON line: static description of the workload, IDENTICAL in both arms."""
    w = WORKLOADS[cfg]
    if cfg == 5:
        return {"workload": "config5: synthetic procedural 10M-triangle scene (9991932 triangles), path tracing (NEE + importance sampling + Russian roulette), "
                            "area + mesh + spherical-environment lights, Torrance-Sparrow / modified Blinn-Phong BRDFs, photographic tonemap, 3840x2160, "
                            "%d spp (BASELINE.json configs[4] names 1024 spp; reduced and stated so that a step is seconds, not minutes)" % spp,
                "resolution": [w["width"], w["height"]], "spp": spp, "triangles": 9991932,
                "sharding": "the same frame at every N: strips of 64x4 pixels dealt round-robin over the GPUs (strong scaling), scene replicated",
                "l2": "flushed between timed iterations (256 MiB write); the 680 MB acceleration structure exceeds the 126 MB L2 anyway"}
    return {"workload": "config2: 996002-triangle procedural mesh + ground + dielectric sphere, mirror/dielectric depth 6, 2 point lights, 1920x1080, 1 spp",
            "resolution": [w["width"], w["height"]], "spp": 1, "triangles": 996002,
            "sharding": "N > 1: sqrt(N) x the resolution per axis, strips of 64x4 pixels dealt round-robin (weak scaling), scene replicated",
            "l2": "flushed between timed iterations (256 MiB write)"}


def scaled_resolution(cfg, n):
    w = WORKLOADS[cfg]
    if w["scaling"] == "strong" or n == 1:
        return w["width"], w["height"]
    s = math.sqrt(n)
    return int(round(w["width"] * s / 8.0)) * 8, int(round(w["height"] * s / 8.0)) * 8


def best_threads(height, cores):
    best = 1
    for t in range(1, max(1, cores) + 1):
        if height % t == 0:
            best = t
    return best


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU DURING the timed region (NVML, ~10 ms period; nvidia-smi
    subprocess as a fallback)."""
    HW_SLOWDOWN, SW_THERMAL, HW_THERMAL, SW_POWER = 0x8, 0x20, 0x40, 0x4

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def _sample_nvml(self):
        n = self.nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        for name, bit in (("hw_slowdown", self.HW_SLOWDOWN), ("hw_thermal_slowdown", self.HW_THERMAL), ("sw_thermal_slowdown", self.SW_THERMAL), ("sw_power_cap", self.SW_POWER)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
        f = [x.strip() for x in out.split(",")]
        if len(f) >= 6:
            self.samples.append(float(f[0]))
            self.max_mhz = float(f[1])
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._halt.wait(0.01 if self.nvml else 0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s),
                "source": "nvml" if self.nvml else "nvidia-smi"}


# ------------------------------------------------------------------ CPU arms (the reference itself, or the oracle port)
def run_reference(xml_path, threads, probe=False):
    """One run of the compiled reference (oracle/_ref) over ALL cameras of the XML.  Returns (seconds, closest, shadow):
    seconds = the binary's own 'Rendering took' (all cameras, excluding scene load), ray counts from the probe build."""
    ref_dir = os.path.join(REPO, "oracle", "_ref")
    exe = os.path.join(ref_dir, "raytracer_probe" if probe else "raytracer")
    env = dict(os.environ, DT_THREADS=str(threads))
    cwd = os.path.dirname(os.path.abspath(xml_path))
    with tempfile.TemporaryFile() as out_f:
        p = subprocess.run([exe, os.path.basename(xml_path)], cwd=cwd, env=env, stdout=out_f, stderr=subprocess.STDOUT, timeout=3600)
        out_f.seek(max(0, out_f.tell() - 4000))
        out = out_f.read().decode(errors="replace")
    if p.returncode != 0:
        raise RuntimeError("reference run failed: " + out[-500:])
    sec = float(re.search(r"Rendering took: ([0-9.eE+-]+)s", out).group(1))
    m = re.search(r"DT_RAYS closest=(\d+) shadow=(\d+)", out)
    return sec, (int(m.group(1)) if m else None), (int(m.group(2)) if m else None)


def have_ref():
    d = os.path.join(REPO, "oracle", "_ref")
    return os.path.exists(os.path.join(d, "raytracer")) and os.path.exists(os.path.join(d, "raytracer_probe"))


def oracle_port_frame(xml_path, threads):
    """Fallback CPU arm when oracle/_ref is absent: the C restatement (oracle/libdtoracle.so)."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from dtb200.scene import HostScene
    from oracle_util import oracle_render
    hs = HostScene(xml_path)
    cam = hs.camera(0)
    t0 = time.perf_counter()
    _, _, st = oracle_render(hs, cam, threads=threads, want_hdr=False)
    return time.perf_counter() - t0, int(st.rays_closest), int(st.rays_shadow)


_ray_count_cache = {}


def cpu_sample(cfg, tmp, frames):
    """Times `frames` steps of the CPU arm on the bounded sample of workload `cfg`.  Returns (seconds per step, rays per step,
    threads, kind, description of the sample)."""
    sw, sh, sspp = WORKLOADS[cfg]["ref_sample"]
    cores = os.cpu_count() or 1
    threads = best_threads(sh, cores)
    what = ("full %dx%d frame" % (sw, sh)) if cfg == 2 else ("the same 9991932-triangle scene at %dx%d, %d spp (reduced from 3840x2160)" % (sw, sh, sspp))
    if have_ref():
        xml1 = make_workload(cfg, os.path.join(tmp, "probe"), sw, sh, sspp)
        if cfg not in _ray_count_cache:
            _ray_count_cache[cfg] = run_reference(xml1, threads, probe=True)[1:]   # ray counts: untimed probe build, one frame
        nc, ns = _ray_count_cache[cfg]
        if cfg == 2 or frames == 1:
            secs = [run_reference(xml1, threads)[0] for _ in range(frames)]
            sec = sum(secs) / len(secs)
        else:
            # the reference renders every <Camera> of a scene in one process (main.cpp:142): `frames` identical cameras give
            # `frames` steps with ONE scene load (12 s for the 10 M-triangle PLY); its timer spans all of them
            xmlk = make_workload(cfg, os.path.join(tmp, "timed"), sw, sh, sspp, n_cameras=frames)
            sec = run_reference(xmlk, threads)[0] / frames
        kind = "reference"
    else:
        xml1 = make_workload(cfg, os.path.join(tmp, "probe"), sw, sh, sspp)
        secs = []
        for _ in range(frames):
            s_, nc, ns = oracle_port_frame(xml1, threads)
            secs.append(s_)
        sec = sum(secs) / len(secs)
        kind = "port"
    desc = "%s; %d render threads (largest divisor of the height <= %d host cores); %d rays per step; time = the binary's own 'Rendering took' (scene load excluded)" % (what, threads, cores, nc + ns)
    return sec, nc + ns, threads, kind, desc


def reference_arm(args, rank, world):
    if rank != 0:
        return 0
    cfg = args.config
    tmp = tempfile.mkdtemp(prefix="dt_bench_ref_")
    if args.warmup > 0:
        cpu_sample(cfg, os.path.join(tmp, "w"), args.warmup)
    sec, rays, threads, kind, desc = cpu_sample(cfg, os.path.join(tmp, "t"), args.steps)
    value = rays / (sec * 1e6)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": WORKLOADS[cfg]["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args.spp),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[2, 5])
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel of config 5 (a perfect square; default 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the config-2 ride-along numbers")
    args = ap.parse_args()
    if args.spp <= 0:
        args.spp = WORKLOADS[args.config]["spp"]
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)

    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    from dtb200 import capi
    from dtb200.scene import HostScene, GpuScene

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = args.config
    wl = WORKLOADS[cfg]
    W, H = scaled_resolution(cfg, world)
    tmp = tempfile.mkdtemp(prefix="dt_bench_r%d_" % rank)
    xml = make_workload(cfg, tmp, W, H, args.spp)
    lib = capi.load_dorktracer()
    if lib.dt_gpu_init(local_rank) < 0:
        raise RuntimeError(lib.dt_last_error().decode())

    def load_scene(path):
        t0 = time.perf_counter()
        hs_ = HostScene(path, gpu_build=True)          # Mesh::ConstructBVH on the GPU (dt_bvh2_build), bit-identical to the host build
        t1 = time.perf_counter()
        gs_ = GpuScene(hs_, device=local_rank)
        return hs_, gs_, t1 - t0, time.perf_counter() - t1

    hs, gs, t_load, t_upload = load_scene(xml)
    cam = hs.camera(0)
    cam.width, cam.height = W, H
    n_pix = W * H

    lib_stream = torch.cuda.ExternalStream(gs.stream_ptr, device=torch.device("cuda", local_rank))
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2
    ldr_host = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
    ldr_np = ldr_host.numpy()

    def hdr_tensor(ptr):
        class _Wrap:
            __cuda_array_interface__ = {"shape": (n_pix * 3,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Wrap(), device="cuda")

    # Multi-GPU gather: rank 0 exports its frame buffers (CUDA IPC), the others import them and their resolve kernel
    # stores the owned strips straight into rank 0's memory over NVLink (DT_FLAG_PEER_FRAME); one barrier orders
    # "all strips written" before "rank 0 reads".  If the IPC mapping is unavailable the ranks fall back, together, to
    # the NCCL reduce of the per-rank radiance frames (SUM == gather: every pixel is non-zero on exactly one rank).
    peer = False
    if world > 1 and not os.environ.get("DT_BENCH_NO_PEER"):
        ok = 1
        obj = [None]
        try:
            if rank == 0:
                obj[0] = gs.frame_export(W, H)
        except Exception as e:
            sys.stderr.write("frame_export failed: %s\n" % e); ok = 0
        dist.broadcast_object_list(obj, src=0)
        try:
            if rank != 0 and obj[0] is not None:
                gs.frame_import(obj[0])
            elif rank != 0:
                ok = 0
        except Exception as e:
            sys.stderr.write("rank %d: frame_import failed: %s\n" % (rank, e)); ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        peer = bool(flag.item())
    peer_flags = capi.DT_FLAG_PEER_FRAME if peer else 0

    split_ms = [0.0, 0.0]                                                      # this rank's render / wait-at-the-barrier parts of the timed steps

    def device_step(timed, scene=None, camera=None):
        """value path: inputs resident, no host copies; the frame ends tonemapped / clamped in device memory."""
        g, c = scene or gs, camera or cam
        flush_buf.fill_(rank + 1)                                              # L2 flush between iterations
        if world > 1:
            dist.barrier()                                                     # untimed: the ranks start the step together (the flush skews them)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(lib_stream)
        ptr, st = g.render_device(c, tile_rank=rank, tile_world=world, flags=peer_flags)
        e1.record(lib_stream)
        if world > 1:
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            if peer:
                dist.barrier()
            else:
                dist.reduce(hdr_tensor(ptr), dst=0, op=dist.ReduceOp.SUM)
            r1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) + r0.elapsed_time(r1)
            if rank == 0 and c.has_tonemapper:                                 # global tonemap of the gathered frame (device only)
                e1.record(lib_stream)
                g.frame_finish(c, on_device=True)
                e2.record(lib_stream)
                torch.cuda.synchronize()
                ms += e1.elapsed_time(e2)
            if timed:
                split_ms[0] += e0.elapsed_time(e1); split_ms[1] += r0.elapsed_time(r1)
        else:
            if c.has_tonemapper:
                g.frame_finish(c, on_device=True)
            e2.record(lib_stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e2)
        return ms, st

    def e2e_step():
        """public API with host buffers: D2H of the finished frame inside the timed region."""
        if world == 1:
            _, _, st = gs.render(cam, ldr=ldr_np, want_hdr=False)
        elif peer:
            ptr, st = gs.render_device(cam, tile_rank=rank, tile_world=world, flags=peer_flags)
            dist.barrier()
            torch.cuda.synchronize()
            if rank == 0:
                gs.frame_finish(cam, ldr=ldr_np)
            dist.barrier()                      # nobody overwrites rank 0's frame before it has been read out
        else:
            ptr, st = gs.render_device(cam, tile_rank=rank, tile_world=world)
            t = hdr_tensor(ptr)
            dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            if rank == 0:
                gs.finish_device(cam, ptr, ldr=ldr_np)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_warm = max(3, args.warmup)
    for _ in range(n_warm):
        device_step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wall0 = time.perf_counter()
    ms_sum, launches = 0.0, 0
    rays_c = rays_s = 0
    waves = 0
    for _ in range(args.steps):
        ms, st = device_step(True)
        ms_sum += ms
        launches += int(st.kernel_launches)
        rays_c += int(st.rays_closest); rays_s += int(st.rays_shadow)
        waves += int(st.waves)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    if world > 1 and os.environ.get("DT_BENCH_VERBOSE"):
        print("[bench] rank %d: render %.3f ms + barrier wait %.3f ms per step" % (rank, split_ms[0] / args.steps, split_ms[1] / args.steps), file=sys.stderr, flush=True)

    # e2e (host buffers), the same number of steps
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0

    # roofline pass (untimed for `value`): the same frame with DT_FLAG_SERIAL_WAVES, so that the CUDA-event time of each
    # traversal launch is that of the kernel running alone
    n_roof = 2 if cfg == 5 else max(3, min(10, args.steps))
    ms_closest = ms_shadow = ms_shade = ms_gen = ms_sort = 0.0
    rays_c_roof = rays_s_roof = 0
    n_closest_launches = n_waves_roof = 0
    bulk = {"waves": 0, "ms_closest": 0.0, "ms_shadow": 0.0, "rays_closest": 0, "rays_shadow": 0}
    for _ in range(n_roof):
        flush_buf.fill_(rank + 1)
        torch.cuda.synchronize()
        _, st = gs.render_device(cam, tile_rank=rank, tile_world=world, flags=capi.DT_FLAG_SERIAL_WAVES | peer_flags)
        ms_closest += st.ms_traverse_closest; ms_shadow += st.ms_traverse_shadow; ms_shade += st.ms_shade; ms_gen += st.ms_generate; ms_sort += st.ms_sort
        rays_c_roof += int(st.rays_closest); rays_s_roof += int(st.rays_shadow)
        n_closest_launches += int(st.launches_traverse_closest); n_waves_roof += int(st.waves)
        bulk["waves"] += int(st.bulk_waves); bulk["ms_closest"] += st.ms_bulk_closest; bulk["ms_shadow"] += st.ms_bulk_shadow
        bulk["rays_closest"] += int(st.rays_bulk_closest); bulk["rays_shadow"] += int(st.rays_bulk_shadow)
    # ray accounting (untimed): `value` counts the rays this library TRACES.  Under Russian roulette it does not follow paths whose
    # weight is exactly zero (dorktracer.h, DT_FLAG_KEEP_WEIGHTLESS_PATHS); the reference does, so the size of the reference's
    # ray tree for the same frame is measured once with the flag and reported beside it.
    rays_ref_tree = 0
    if cfg == 5:
        _, st = gs.render_device(cam, tile_rank=rank, tile_world=world, flags=capi.DT_FLAG_KEEP_WEIGHTLESS_PATHS | peer_flags)
        rays_ref_tree = int(st.rays_closest) + int(st.rays_shadow)
    barrier()

    # max over ranks of the times, sum over ranks of the rays
    vals = torch.tensor([ms_sum, e2e_s], dtype=torch.float64, device="cuda")
    cnts = torch.tensor([rays_c, rays_s, launches, rays_ref_tree], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnts, op=dist.ReduceOp.SUM)
    ms_sum_max, e2e_s_max = vals.tolist()
    rays_c_all, rays_s_all, launches_all, rays_ref_tree_all = cnts.tolist()

    if rank == 0:
        rays_per_step = (rays_c_all + rays_s_all) / args.steps
        ms_per_step = ms_sum_max / args.steps
        value = rays_per_step / (ms_per_step * 1e3)
        e2e_value = rays_per_step / (e2e_s_max / args.steps * 1e6)
        peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
        peak, peak_src, sm_mhz_max = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)", 1965.0
        if os.path.exists(peaks_path):
            try:
                pk = json.load(open(peaks_path))
                peak = float(pk["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
                sm_mhz_max = float(pk.get("sm_max_mhz", sm_mhz_max))
            except Exception:
                pass
        ncu = {}
        try:
            ncu = json.load(open(os.path.join(REPO, "profiles", "ncu_config%d_summary.json" % cfg)))
        except Exception:
            pass
        # dominant kernel = the traversal kernel with the larger share of the frame (kernels alone: SERIAL_WAVES pass, rank 0)
        # Path-traced frames: only the "bulk" waves (at least half of max_wave_rays alive) count -- the thousands of tail waves
        # of a Russian-roulette frame hold a handful of rays each and are all launch latency (dt_stats.bulk_*).
        use_bulk = cfg == 5 and bulk["waves"] > 0
        if use_bulk:
            kinds = {"closest": ("k_traverse_dyn<false> (closest-hit, persistent warps)", bulk["ms_closest"], bulk["rays_closest"], wl["b_closest"], bulk["waves"]),
                     "shadow": ("k_traverse_dyn<true> (any-hit / shadow rays, persistent warps)", bulk["ms_shadow"], bulk["rays_shadow"], wl["b_shadow"], bulk["waves"])}
        else:
            kinds = {"closest": ("k_traverse_dyn<false> (closest-hit, persistent warps)", ms_closest, rays_c_roof, wl["b_closest"], n_closest_launches),
                     "shadow": ("k_traverse_dyn<true> (any-hit / shadow rays, persistent warps)", ms_shadow, rays_s_roof, wl["b_shadow"], n_waves_roof)}
        dom = "shadow" if kinds["shadow"][1] > kinds["closest"][1] else "closest"
        other = "closest" if dom == "shadow" else "shadow"

        def roof(which):
            name, ms_k, rays_k, b_ray, launches_k = kinds[which]
            launches_k = max(1, launches_k)
            gbs = b_ray * rays_k / 1e9 / (ms_k / 1e3) if ms_k > 0 else 0.0
            nk = (ncu.get("kernels") or {}).get(which) or {}
            inst_per_ray = nk.get("warp_inst_per_ray")
            n_sm = int(torch.cuda.get_device_properties(local_rank).multi_processor_count)
            issue_peak = 4.0 * n_sm * sm_mhz_max * 1e6 / 1e9                                   # G warp-instructions / s
            issue = None
            if inst_per_ray and ms_k > 0:
                ach = inst_per_ray * rays_k / (ms_k / 1e3) / 1e9
                issue = {"achieved": ach, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": ach / issue_peak, "warp_inst_per_ray": inst_per_ray,
                         "ncu_issue_active_pct": nk.get("issue_active_pct"), "ncu_threads_per_inst": nk.get("threads_per_inst"), "ncu_branch_uniform_pct": nk.get("branch_uniform_pct")}
            hbm = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "algorithmic_bytes_per_ray": b_ray,
                   "algorithmic_bytes_per_launch": b_ray * rays_k / launches_k, "peak_source": peak_src}
            return name, ms_k, rays_k, launches_k, issue, hbm, nk

        name, ms_k, rays_k, launches_k, issue, hbm, nk = roof(dom)
        _, _, _, _, issue2, hbm2, _ = roof(other)
        how = ("CUDA events around each launch of %d frames rendered with DT_FLAG_SERIAL_WAVES (kernel alone on the GPU), L2 flushed before each frame%s; "
               "warp instructions per ray from the committed ncu capture of the same workload" % (n_roof, "; launches of the %d bulk waves (>= half of max_wave_rays alive) only" % bulk["waves"] if use_bulk else ""))
        if issue:
            roofline = {"bound": "issue", "kernel": name, "achieved": issue["achieved"], "peak": issue["peak"], "unit": issue["unit"], "frac": issue["frac"],
                        "traffic": nk.get("traffic_bytes_per_launch"), "issue": issue, "hbm": hbm}
        else:   # no ncu summary for this workload committed yet: the HBM figure alone
            roofline = dict(hbm, kernel=name, traffic=None, issue=None, hbm=hbm)
        roofline.update({"rays_per_launch": rays_k / launches_k, "avg_launch_ms": ms_k / launches_k, "how": how, "ncu_source": ncu.get("source"),
                         "other_kernel": {"kernel": kinds[other][0], "issue": issue2, "hbm": hbm2}})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, args.spp),
            "details": {
                "resolution_rendered": [W, H], "triangles": hs.n_triangles(),
                "rays_per_step": rays_per_step, "closest_rays_per_step": rays_c_all / args.steps, "shadow_rays_per_step": rays_s_all / args.steps,
                "waves_per_step_rank0": waves / args.steps,
                "ray_accounting": ({"rays_per_step": "rays traced by this library (what `value` and `e2e` count)",
                                    "rays_per_step_of_the_reference_ray_tree": rays_ref_tree_all,
                                    "mrays_per_s_counting_the_reference_ray_tree": rays_ref_tree_all / (ms_per_step * 1e3),
                                    "note": "under Russian roulette hits whose path weight is exactly (0,0,0) are not shaded (everything below them is an exact zero) and "
                                            "shadow rays whose contribution is exactly zero are not traced; the image is the same "
                                            "(tests/test_gpu_parity.py::test_config5_shape_path_tracing_robust_statistics); the reference follows / traces them, "
                                            "so its ray tree for this frame is larger (one untimed frame with DT_FLAG_KEEP_WEIGHTLESS_PATHS)"}
                                   if rays_ref_tree_all > 0 else None),
                "timing": "CUDA events on the library stream (render + resolve + tonemap; + barrier / reduce for N > 1), max over ranks",
                "gather": ("peer-memory stores (CUDA IPC, fused into the resolve kernel) + one barrier" if peer else "nccl-reduce") if world > 1 else "none",
                "scene_load_s": t_load, "scene_upload_s": t_upload,
                "stage_ms_per_step_rank0_kernels_alone": {"generate": ms_gen / n_roof, "traverse_closest": ms_closest / n_roof, "sort": ms_sort / n_roof, "shade": ms_shade / n_roof,
                                                          "traverse_shadow": ms_shadow / n_roof, "note": "DT_FLAG_SERIAL_WAVES pass of %d frames after the timed region" % n_roof},
                "wall_s_timed_region": t_wall,
            },
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(capi.dt_camera_desc) + C.sizeof(capi.dt_render_params),
                    "d2h_bytes_per_step": n_pix * 3, "ms_per_step": 1e3 * e2e_s_max / args.steps},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
        }
        if world == 1 and cfg == 5 and not args.no_other:
            # the round-1 headline workload rides along (N = 1 only): config 2, 20 frames, device-timed like `value`
            try:
                gs.close()
                xml2 = make_workload(2, tmp, 1920, 1080, 1)
                hs2, gs2, _, _ = load_scene(xml2)
                cam2 = hs2.camera(0)
                lib_stream = torch.cuda.ExternalStream(gs2.stream_ptr, device=torch.device("cuda", local_rank))
                for _ in range(5):
                    device_step(False, gs2, cam2)
                tot, rr = 0.0, 0
                for _ in range(20):
                    ms, st = device_step(False, gs2, cam2)
                    tot += ms; rr = int(st.rays_closest) + int(st.rays_shadow)
                line["other_workloads"] = {"config2": {"workload": workload_config(2, 1)["workload"], "value": rr / (tot / 20 * 1e3), "unit": UNIT, "ms_per_step": tot / 20,
                                                       "rays_per_step": rr, "steps": 20, "warmup": 5, "timing": "device, as `value`"}}
                gs2.close()
            except Exception as e:
                line["other_workloads"] = {"config2": {"error": str(e)[:200]}}
        if world == 1 and not args.no_cpu_baseline:
            try:
                sec, rays, threads, kind, desc = cpu_sample(cfg, os.path.join(tmp, "cpu"), 1)
                line["cpu_baseline"] = {"value": rays / (sec * 1e6), "unit": UNIT, "cores": threads, "kind": kind, "seconds": sec, "sample": desc}
            except Exception as e:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
