#!/usr/bin/env python3
"""bench.py — headline benchmark of the render hot path (contract: see the task's bench section).

Metric (BASELINE.json): Mrays/s (primary + secondary + shadow) for the whole job, frame ms.
A *step* is one frame of the workload through the hot path:
  N = 1 : BASELINE.json configs[1] — ~1M-triangle mesh (procedural "blob": the named bunny/dragon assets are
          stripped from the reference mount), mirror + dielectric recursion depth 6, 2 point lights, 1920x1080.
  N > 1 : the same scene and view with sqrt(N) x the resolution per axis (per-GPU pixel count fixed -> weak
          scaling); the image is tile-sharded (strips of eight 8x4-pixel tiles, round-robin) across the ranks, each rank renders
          its tiles with the whole scene replicated, and the per-rank radiance frames are combined on rank 0
          without a full-frame exchange: rank 0 exports its frame buffers (CUDA IPC) and every rank's resolve kernel
          stores its tiles straight into them over NVLink (P2P stores), then one barrier.  Fallback when the IPC
          mapping is unavailable: ONE NCCL reduce (every pixel is non-zero on exactly one rank, so SUM == gather).

  value      device time only: scene + camera resident, CUDA events on the library's stream (+ the reduce)
  e2e        the public C-ABI call dt_render() with a pinned HOST LDR buffer: D2H of the frame inside the
             timed region, wall clock between device synchronisations
  roofline   closest-hit traversal kernel: algorithmic bytes/ray (736 B for config 2) x rays
             / its CUDA-event time (measured live in a separate DT_FLAG_SERIAL_WAVES pass, each kernel alone on
             the GPU), against the measured HBM peak (MEASURED_PEAKS.json); `traffic` and the issue-slot / SIMT
             figures come from the committed ncu capture (profiles/ncu_traverse_summary.json)
  cpu_baseline  the compiled reference (oracle/_ref/raytracer) on this box's host cores, same frame

`--impl reference` times the reference's own CPU renderer on the same workload (rank 0 only).
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
B_RAY_CLOSEST = 736.0          # 32 in + 32 out + 6 x 80 B nodes + 4 x 48 B triangles
B_RAY_SHADOW = 708.0           # same minus the 32 B hit record plus a 4 B flag


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def make_workload(tmp, nlon=1000, nlat=499):
    from dtb200 import scenegen
    return scenegen.gen_config2(os.path.join(tmp, "c2"), nlon=nlon, nlat=nlat, width=1920, height=1080, depth=6)


def scaled_resolution(n):
    s = math.sqrt(n)
    w = int(round(1920 * s / 8.0)) * 8
    h = int(round(1080 * s / 8.0)) * 8
    return w, h


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


def run_reference_frame(xml_path, threads, probe=False):
    """One frame of the compiled reference (oracle/_ref).  Returns (seconds, closest, shadow)."""
    ref_dir = os.path.join(REPO, "oracle", "_ref")
    exe = os.path.join(ref_dir, "raytracer_probe" if probe else "raytracer")
    env = dict(os.environ, DT_THREADS=str(threads))
    cwd = os.path.dirname(os.path.abspath(xml_path))
    p = subprocess.run([exe, os.path.basename(xml_path)], cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=3600)
    out = p.stdout.decode(errors="replace")
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


def reference_arm(args, rank, world):
    if rank != 0:
        return 0
    tmp = tempfile.mkdtemp(prefix="dt_bench_ref_")
    xml = make_workload(tmp)
    cores = os.cpu_count() or 1
    threads = best_threads(1080, cores)
    kind = "reference" if have_ref() else "port"
    if kind == "reference":
        _, nc, ns = run_reference_frame(xml, threads, probe=True)         # ray counts (untimed probe build)
        frame = lambda: run_reference_frame(xml, threads)[0]
    else:
        _, nc, ns = oracle_port_frame(xml, threads)
        frame = lambda: oracle_port_frame(xml, threads)[0]
    for _ in range(args.warmup):
        frame()
    secs = [frame() for _ in range(args.steps)]
    ms = 1e3 * sum(secs) / len(secs)
    value = (nc + ns) / (ms * 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config2: 996002-triangle procedural mesh + ground + dielectric sphere, mirror/dielectric depth 6, 2 point lights, 1920x1080, 1 spp",
                   "rays_per_step": nc + ns, "note": "CPU arm renders the N=1 frame (bounded sample) on the host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "full 1920x1080 frame of the workload, %d render threads (largest divisor of 1080 <= %d host cores), time = the binary's own 'Rendering took'" % (threads, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tris", default="1000x499", help="blob tessellation nlon x nlat (default = 996000 triangles)")
    args = ap.parse_args()
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

    nlon, nlat = (int(x) for x in args.tris.split("x"))
    tmp = tempfile.mkdtemp(prefix="dt_bench_r%d_" % rank)
    xml = make_workload(tmp, nlon, nlat)
    if capi.load_dorktracer().dt_gpu_init(local_rank) < 0:
        raise RuntimeError(capi.load_dorktracer().dt_last_error().decode())
    t0 = time.perf_counter()
    hs = HostScene(xml, gpu_build=True)          # Mesh::ConstructBVH on the GPU (dt_bvh2_build), bit-identical to the host build
    t_load = time.perf_counter() - t0
    t0 = time.perf_counter()
    gs = GpuScene(hs, device=local_rank)
    t_upload = time.perf_counter() - t0
    cam = hs.camera(0)
    W, H = scaled_resolution(world)
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
    # stores the owned tiles straight into rank 0's memory over NVLink (DT_FLAG_PEER_FRAME); one barrier orders
    # "all tiles written" before "rank 0 reads".  If the IPC mapping is unavailable the ranks fall back, together, to
    # the NCCL reduce of the per-rank radiance frames.
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

    def device_step(timed):
        """value path: inputs resident, no host copies.  Returns (ms, stats)."""
        flush_buf.fill_(rank + 1)                                              # L2 flush between iterations
        if world > 1:
            dist.barrier()                                                     # untimed: the ranks start the step together (the flush skews them)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        ptr, st = gs.render_device(cam, tile_rank=rank, tile_world=world, flags=peer_flags)
        e1.record(lib_stream)
        ms = None
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
            if timed:
                split_ms[0] += e0.elapsed_time(e1); split_ms[1] += r0.elapsed_time(r1)
        else:
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
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

    for _ in range(max(3, args.warmup)):
        device_step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wall0 = time.perf_counter()
    ms_sum, launches = 0.0, 0
    rays_c = rays_s = 0
    ov_closest = ov_shadow = ov_shade = 0.0
    for _ in range(args.steps):
        ms, st = device_step(True)
        ms_sum += ms
        launches += int(st.kernel_launches)
        rays_c += int(st.rays_closest); rays_s += int(st.rays_shadow)
        ov_closest += st.ms_traverse_closest; ov_shadow += st.ms_traverse_shadow; ov_shade += st.ms_shade
    barrier()
    if world > 1 and os.environ.get("DT_BENCH_VERBOSE"):
        print("[bench] rank %d: render %.3f ms + barrier wait %.3f ms per step" % (rank, split_ms[0] / args.steps, split_ms[1] / args.steps), file=sys.stderr, flush=True)
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()

    # e2e (host buffers)
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0

    # roofline pass (untimed for `value`): the same frame with DT_FLAG_SERIAL_WAVES, so that the CUDA-event time of each
    # traversal launch is that of the kernel running alone (in the default mode shadow(k) overlaps closest(k+1))
    n_roof = max(3, min(10, args.steps))
    ms_closest = ms_shadow = ms_shade = ms_gen = 0.0
    rays_c_roof = rays_s_roof = 0
    n_closest_launches = 0
    for _ in range(n_roof):
        flush_buf.fill_(rank + 1)
        torch.cuda.synchronize()
        _, st = gs.render_device(cam, tile_rank=rank, tile_world=world, flags=capi.DT_FLAG_SERIAL_WAVES | peer_flags)
        ms_closest += st.ms_traverse_closest; ms_shadow += st.ms_traverse_shadow; ms_shade += st.ms_shade; ms_gen += st.ms_generate
        rays_c_roof += int(st.rays_closest); rays_s_roof += int(st.rays_shadow)
        n_closest_launches += int(st.launches_traverse_closest)
    barrier()

    # max over ranks of the times, sum over ranks of the rays
    vals = torch.tensor([ms_sum, e2e_s, ms_closest, ms_shadow], dtype=torch.float64, device="cuda")
    cnts = torch.tensor([rays_c, rays_s, launches, n_closest_launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnts, op=dist.ReduceOp.SUM)
    ms_sum_max, e2e_s_max, ms_closest_max, ms_shadow_max = vals.tolist()
    rays_c_all, rays_s_all, launches_all, closest_launches_all = cnts.tolist()

    if rank == 0:
        rays_per_step = (rays_c_all + rays_s_all) / args.steps
        ms_per_step = ms_sum_max / args.steps
        value = rays_per_step / (ms_per_step * 1e3)
        e2e_value = rays_per_step / (e2e_s_max / args.steps * 1e6)
        peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        if os.path.exists(peaks_path):
            try:
                peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
            except Exception:
                pass
        # roofline of the dominant kernel (closest-hit traversal): rank-0-local figures
        gb_closest = B_RAY_CLOSEST * rays_c_roof / 1e9
        achieved = gb_closest / (ms_closest / 1e3) if ms_closest > 0 else 0.0
        gb_shadow = B_RAY_SHADOW * rays_s_roof / 1e9
        achieved_shadow = gb_shadow / (ms_shadow / 1e3) if ms_shadow > 0 else 0.0
        ncu = {}
        try:
            ncu = json.load(open(os.path.join(REPO, "profiles", "ncu_traverse_summary.json")))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "config2: %d-triangle procedural mesh + ground + dielectric sphere, mirror/dielectric depth 6, 2 point lights, %dx%d, 1 spp%s"
                            % (hs.n_triangles(), W, H, "" if world == 1 else " (1920x1080 x %d pixels, tile-sharded over %d GPUs, %s)" % (world, world, "tiles stored into rank 0's frame over NVLink by the resolve kernel (CUDA IPC peer memory) + one barrier" if peer else "NCCL reduce to rank 0")),
                "rays_per_step": rays_per_step, "closest_rays_per_step": rays_c_all / args.steps, "shadow_rays_per_step": rays_s_all / args.steps,
                "l2": "flushed between timed iterations (256 MiB write)", "timing": "CUDA events on the library stream (+ barrier / reduce), max over ranks", "gather": ("peer-memory stores" if peer else "nccl-reduce") if world > 1 else "none",
                "scene_load_s": t_load, "scene_upload_s": t_upload,
                "stage_ms_per_step_rank0_kernels_alone": {"generate": ms_gen / n_roof, "traverse_closest": ms_closest / n_roof, "shade": ms_shade / n_roof, "traverse_shadow": ms_shadow / n_roof,
                                                          "note": "DT_FLAG_SERIAL_WAVES pass of %d frames after the timed region" % n_roof},
                "stage_ms_per_step_rank0_overlapped": {"traverse_closest": ov_closest / args.steps, "shade": ov_shade / args.steps, "traverse_shadow": ov_shadow / args.steps,
                                                       "note": "timed region; shadow(k) runs concurrently with closest(k+1)/shade(k+1), so these sum to more than ms_per_step"},
                "wall_s_timed_region": t_wall,
            },
            "roofline": {"bound": "hbm", "kernel": "k_traverse<false> (closest-hit, persistent warps)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu.get("traffic_bytes_per_launch"), "peak_source": peak_src,
                         "algorithmic_bytes_per_ray": B_RAY_CLOSEST, "rays_per_launch": rays_c_roof / max(1, n_closest_launches),
                         "algorithmic_bytes_per_launch": B_RAY_CLOSEST * rays_c_roof / max(1, n_closest_launches),
                         "avg_launch_ms": ms_closest / max(1, n_closest_launches),
                         "how": "CUDA events around each of the %d closest-hit launches of %d frames rendered with DT_FLAG_SERIAL_WAVES (kernel alone on the GPU), L2 flushed before each frame" % (n_closest_launches, n_roof),
                         "ncu": ncu.get("ncu"), "ncu_source": ncu.get("source"),
                         "shadow_kernel": {"achieved": achieved_shadow, "frac": achieved_shadow / peak, "algorithmic_bytes_per_ray": B_RAY_SHADOW}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": C.sizeof(capi.dt_camera_desc) + C.sizeof(capi.dt_render_params),
                    "d2h_bytes_per_step": n_pix * 3, "ms_per_step": 1e3 * e2e_s_max / args.steps},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cores = os.cpu_count() or 1
                threads = best_threads(1080, cores)
                if have_ref():
                    sec, _, _ = run_reference_frame(xml, threads)
                    kind = "reference"
                else:
                    sec, _, _ = oracle_port_frame(xml, threads)
                    kind = "port"
                line["cpu_baseline"] = {"value": rays_per_step / (sec * 1e6), "unit": UNIT, "cores": threads, "kind": kind, "seconds": sec,
                                        "sample": "one full 1920x1080 frame of the same workload (%d rays), %d render threads on %d host cores" % (int(rays_per_step), threads, cores)}
            except Exception as e:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
