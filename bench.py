#!/usr/bin/env python
"""bench.py -- headline benchmark of the ray-march hot path (BASELINE.json: rays/sec and ms/frame at
1080p, 1M particles; frames/sec over a sequence at 1-8 GPUs).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation

A "step" is one pass of the whole hot path over one synthetic 1M-particle frame at 1920x1080
(BASELINE.json configs[1]): grid build (neighbour search + AABB + occupancy grid) -> depth pre-pass ->
ray march + normals + shading.  `value` starts with the particle arrays resident in HBM; `e2e` runs the
same step through the C ABI with HOST buffers (pinned): host->device copy of the particles and
device->host copy of the RGBA image inside the timed region (`e2e.reference_protocol`: positions and
normals too).  Both go through the sequence API (fr_seq_*, --lanes frames in flight per GPU; K steps are
one timed region); `config.latency_ms_per_frame`, the per-stage times and the roofline kernel time come
from one frame at a time on one context with L2 flushed between steps.

N > 1 (torchrun, one rank per GPU): frame-parallel over an animation sequence -- every rank renders its
own frames, no data-path collective (pure partitioning), `scaling: weak`.  `--mode tiles` instead splits
ONE frame into interleaved 64x64 screen tiles per rank and gathers the RGBA tiles on rank 0 with NCCL
(bachelor-thesis_b200/multigpu.py; `scaling: strong`).

Timing: CUDA events on the stream the kernels are launched on (the context stream), L2 flushed between
steps by zeroing a 512 MiB buffer outside the timed events, max over ranks.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rays_per_sec_1080p_1M_particles"
UNIT = "rays/s"

CONFIGS = {
    # name: (particles, W, H, h, dx)
    "C1": (64_000, 1280, 720, 0.1, None),
    "C2": (1_000_000, 1920, 1080, 0.1, None),
    "C3": (4_000_000, 3840, 2160, 0.063, 0.0315),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_ncu_stats(cfg_name, kernel):
    """per-launch ncu figures of the dominant kernel for this workload (profiles/ncu_stats.json, written from the
    committed ncu captures by tools/ncu_summary.py); {} when there is no capture for it"""
    path = os.path.join(ROOT, "profiles", "ncu_stats.json")
    try:
        with open(path) as f:
            return json.load(f).get(cfg_name, {}).get(kernel, {})
    except Exception:
        return {}


def workload_string(cfg_name, n_actual, W, H):
    return (f"{cfg_name}: dam-break {n_actual} particles, {W}x{H}, default camera (R=10, fov 60), isotropic, "
            "MaxSteps 128, StepSize 0.009, Iso 1.0")


def camera():
    """reference default: orbit R = 10, fov 60 deg -- the exact bits the reference's camera TUs produce"""
    return importlib.import_module("bachelor-thesis_b200.camera").reference_default_camera()


def checker():
    """the CPU checkers (oracle/ + tests/oracle_lib.py): cpu_baseline leg and --impl reference only"""
    tests = os.path.join(ROOT, "tests")
    if tests not in sys.path:
        sys.path.insert(0, tests)
    import oracle_lib
    return oracle_lib


FRAME_T = 0.6      # the ONE synthetic frame every arm renders (scenes.dam_break default; tests/test_gpu_fullsize.py pins it)


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region.  NVML is polled from a thread every ~1 ms (the timed
    region of 20 steps is only ~15 ms, shorter than one `nvidia-smi -lms` period); nvidia-smi is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        # NVML queries go through the driver and disturb concurrent kernel launches: at 1 kHz the pipelined arm
        # (6 lanes launching ~120 kernels per frame time) became erratic (0.30 .. 0.66 ms per frame, r01u)
        self.period = float(os.environ.get("BENCH_SAMPLER_MS", "5")) * 1e-3
        self.sm, self.mx, self.reasons = [], [], set()
        self.running = False
        self.thread = None
        self.nvml = None
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        nv = self.nvml
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while self.running:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.mx.append(float(self.max_sm))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, b in bits.items():
                    if r & b:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(self.period)

    def _poll_smi(self):
        while self.running:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout
            except Exception:
                break
            for ln in out.splitlines():
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 9:
                    continue
                try:
                    self.sm.append(float(p[1]))
                    self.mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)

    def start(self):
        self.running = True
        self.thread = threading.Thread(target=self._poll_nvml if self.nvml else self._poll_smi, daemon=True)
        self.thread.start()

    def stop(self):
        self.running = False
        if self.thread:
            self.thread.join(timeout=6)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "nvml" if self.nvml else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------
def reference_step_fn(cfg_name, threads=0):
    """the reference's CPU path for one frame: Frame::Frame (search + AABB + density grid) + the per-pixel march
    over all pixels on `threads` host threads.  The depth image is an INPUT of the reference marcher (its
    GPU raster pass makes it); it is produced once, outside the timed step, by the oracle's restatement.
    step(pool=False) -> (seconds, build seconds, march seconds, positions, normals); pool=True runs the march on the
    reference's own ThreadPool (hardware_concurrency() - 1 workers, one atomic per pixel, ThreadPool.cpp:38-55)."""
    oracle_lib = checker()
    n, W, H, h, dx = CONFIGS[cfg_name]
    fm_scenes = importlib.import_module("bachelor-thesis_b200.scenes")
    xyz = fm_scenes.dam_break(n, h=h, dx=dx, t=FRAME_T)
    cam = camera()
    orc = oracle_lib.Oracle()
    depth = orc.frame(xyz, h, 2.0).depth_prepass(W, H, cam["view"], cam["proj"])
    s = oracle_lib.Settings()
    if oracle_lib.ref_available():
        ref = oracle_lib.Ref()
        kind = "reference"
        nthreads = threads or ref.lib.ref_hardware_threads()

        def step(pool=False):
            t0 = time.perf_counter()
            ds = ref.dataset(xyz, h, 2.0)
            t1 = time.perf_counter()
            pos, nrm, march_s = ds.march(W, H, s, cam["inv_proj_view"], cam["position"], depth, threads=nthreads,
                                         use_ref_pool=1 if pool else 0)
            ds.close()
            return time.perf_counter() - t0, t1 - t0, march_s, pos, nrm
    else:
        kind = "port"
        nthreads = threads or orc.lib.fo_get_threads()

        def step(pool=False):
            t0 = time.perf_counter()
            f = orc.frame(xyz, h, 2.0)
            t1 = time.perf_counter()
            pos, nrm, *_ = f.march(W, H, s, cam["inv_proj_view"], cam["position"], depth, threads=nthreads, want_band=False)
            t2 = time.perf_counter()
            return t2 - t0, t1 - t0, t2 - t1, pos, nrm
    return step, kind, nthreads, (n, W, H, len(xyz)), depth


def image_digest(pos, nrm):
    """sha256 over the bits of the hit positions and normals: equal in both arms <=> bit-identical images"""
    import hashlib
    m = hashlib.sha256()
    m.update(np.ascontiguousarray(pos, np.float32).tobytes())
    m.update(np.ascontiguousarray(nrm, np.float32).tobytes())
    return m.hexdigest()[:16]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, kind, nthreads, (n, W, H, n_actual), _ = reference_step_fn(args.config)
    budget = time.perf_counter() + 240.0     # full C2 frames take ~0.45 s each on the host cores
    warm = 0
    for _ in range(args.warmup):
        step()
        warm += 1
        if time.perf_counter() > budget - 120.0:
            break
    times, builds, marches = [], [], []
    pos = nrm = None
    for _ in range(args.steps):
        t, b, m, pos, nrm = step()
        times.append(t); builds.append(b); marches.append(m)
        if time.perf_counter() > budget:
            break
    ms = 1e3 * sum(times) / len(times)
    value = W * H / (ms * 1e-3)
    pool = None
    if kind == "reference":
        tp = step(pool=True)
        pool = {"ms_per_step": 1e3 * tp[0], "march_ms": 1e3 * tp[2], "value": W * H / tp[0],
                "note": "the march on the reference's own ThreadPool (hardware_concurrency() - 1 workers, one atomic per pixel, "
                        "ThreadPool.cpp:38-55; pixel W*H-1 is skipped there) instead of the harness's chunked parallel-for"}
    sample = (f"{len(times)} full frames of {args.config} ({n_actual} particles, {W}x{H}): Frame::Frame build "
              f"{1e3 * sum(builds) / len(builds):.0f} ms + march {1e3 * sum(marches) / len(marches):.0f} ms per frame; "
              "depth image precomputed (the reference rasterises it on its GPU); neighbour search = API-compatible "
              "stand-in for the un-vendored CompactNSearch fork")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.config, n_actual, W, H),
                   "step": "Frame::Frame (search + AABB + density grid) + PerPixel_Isotropic over all pixels, host cores",
                   "image_sha256_16": image_digest(pos, nrm)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": kind, "sample": sample, "reference_thread_pool": pool},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)



class _DevPtr:
    """a raw device pointer as a CUDA array (for torch.as_tensor): lets the bench poison a library-owned image"""
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def host_ceiling(torch, dist, dev, world, h2d_bytes, d2h_bytes, reps=40):
    """what the host side of this box sustains with all ranks copying at once: pinned host -> device of one frame's
    particles and device -> pinned host of one image, on two streams, nothing else running.  GB/s per rank."""
    h_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def pair():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    for _ in range(3):
        pair()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = float("inf")
    for _ in range(5):                   # a capability: the best of five rounds (other host threads disturb single rounds)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            pair()
        torch.cuda.synchronize()
        ms = min(ms, (time.perf_counter() - t0) / reps * 1e3)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"ms_per_frame_pair": ms, "gbs_per_rank": (h2d_bytes + d2h_bytes) / ms / 1e6,
            "gbs_all_ranks": world * (h2d_bytes + d2h_bytes) / ms / 1e6,
            "note": "all ranks at once: one frame's particles host -> device and one RGBA image device -> host per pair, pinned "
                    "memory, two streams, max over ranks; the e2e arm cannot be faster than this per frame"}


def balanced_strips(weights, W, world):
    """vertical strips [x0, x1) with bounds at multiples of 64 pixels and about equal cost: `weights` = covered pixels per
    64-pixel column block of the frame (from the previous render of the sequence -- here the one-GPU render of the same
    frame)"""
    nb = (W + 63) // 64
    w = np.asarray(weights, np.float64)[:nb]
    cum = np.concatenate([[0.0], np.cumsum(w)])
    bounds = [0]
    for k in range(1, world):
        target = cum[-1] * k / world
        b = int(np.argmin(np.abs(cum - target)))
        b = max(b, bounds[-1] + 1)
        b = min(b, nb - (world - k))
        bounds.append(b)
    bounds.append(nb)
    return [(64 * bounds[k], min(64 * bounds[k + 1], W)) for k in range(world)]


def tiles_subrecord(fm, torch, dist, dev, rank, world, local, cam_args, cam, tile, steps=12):
    """BASELINE config C3 (4M particles, 3840x2160): ONE frame split over the ranks, every rank's shading epilogue storing
    its pixels straight into the presenting GPU's image over NVLink peer memory (fr_ipc_*); strong scaling against the
    same frame rendered by one GPU alone, and bit-identity of the assembled image.  Two partitions: `regions` -- each
    rank owns a vertical strip and builds its frame from the particles that can influence it (fr_set_region_partition:
    build, pre-pass and march all shrink) -- and, for comparison, interleaved tiles over replicated frames."""
    n, W, H, h, dx = CONFIGS["C3"]
    xyz = fm.scenes.dam_break(n, h=h, dx=dx, t=FRAME_T)
    npart = len(xyz)
    d_xyz = torch.from_numpy(xyz).to(dev)
    ctx = fm.Context(W, H, device=local)
    ctx.set_camera(*cam_args)
    ctx.set_settings(fm.VisualizationSettings())
    ctx.set_stage_timing(False)       # one frame at a time, as the RayMarcher shim runs it: no per-stage events (latency mode)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def timed(step_fn, k):
        for _ in range(3):
            step_fn()
        ctx.wait()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        total, wall = 0.0, 0.0
        for _ in range(k):
            with torch.cuda.stream(stream):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(stream):
                e0.record(stream)
            step_fn()
            with torch.cuda.stream(stream):
                e1.record(stream)
            ctx.wait()
            if world > 1:
                dist.barrier()                  # every rank's pixels have landed in the presenter's image
            wall += time.perf_counter() - t0
            total += e0.elapsed_time(e1)
        ms = total / k
        mine = ms
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, 1e3 * wall / k, mine

    def step():
        ctx.build_frame_device(0, d_xyz.data_ptr(), npart, h, 2.0)
        ctx.render_async(fm.FR_PASS_ALL)

    one_ms, _, _ = timed(step, steps)                     # this GPU alone, the whole frame
    rec = {"workload": workload_string("C3", npart, W, H), "n_gpus": world, "ms_per_frame_one_gpu": one_ms}
    if world == 1:
        ctx.close()
        return rec
    full_depth, _, _, full_rgba = ctx.download(True, False, False, True)
    want = full_rgba.copy() if rank == 0 else None
    handle = [ctx.ipc_export_color() if rank == 0 else None]
    dist.broadcast_object_list(handle, src=0)
    if rank != 0:
        ctx.ipc_open_color_target(handle[0])

    def poison():
        if rank == 0:
            torch.as_tensor(_DevPtr(ctx.device_images()["rgba"], W * H * 4), device=dev).fill_(0x5a)      # stale pixels must not pass
            torch.cuda.synchronize()
        dist.barrier()

    def check():
        same, differing = None, None
        if rank == 0:
            got = ctx.download(False, False, False, True)[3]
            differing = int((got != want).any(-1).sum())
            same = differing == 0
        dist.barrier()
        return same, differing

    # (1) regions: vertical strips balanced by particle count, filtered frame builds
    covered_cols = (full_depth != 1.0).sum(axis=0).astype(np.float64)
    pad = (-len(covered_cols)) % 64
    weights = np.pad(covered_cols, (0, pad)).reshape(-1, 64).sum(axis=1) + 1.0
    # feedback balancing, as a sequence renderer does from frame to frame: start from the covered pixels per 64-pixel
    # column block, then rescale every strip's blocks by its rank's measured time and split again
    balance_log = []
    for it in range(4):
        strips = balanced_strips(weights, W, world)
        x0, x1 = strips[rank]
        ctx.set_region_partition(x0, 0, x1, H)
        dist.barrier()
        _, _, my_try = timed(step, 3)
        tries = [None] * world
        dist.all_gather_object(tries, float(my_try))
        balance_log.append([round(t, 4) for t in tries])
        if it == 3:
            break
        mean_t = sum(tries) / world
        for k, (a, b) in enumerate(strips):
            weights[a // 64:(b + 63) // 64] *= (tries[k] / mean_t) ** 1.5
    poison()
    n_ms, n_wall, my_ms = timed(step, steps)
    same, differing = check()
    per_rank = [None] * world
    dist.all_gather_object(per_rank, round(my_ms, 4))
    ctx.set_region_partition()
    # (2) interleaved tiles over replicated frames (round 1's partition)
    ctx.set_tile_partition(rank, world, tile, tile)
    poison()
    t_ms, t_wall, _ = timed(step, max(4, steps // 2))
    t_same, t_diff = check()
    if rank != 0:
        ctx.ipc_close_color_target()
    ctx.close()
    rec.update({"ms_per_frame": n_ms, "ms_per_frame_wall_with_barrier": n_wall, "scaling": "strong",
                "efficiency_vs_one_gpu": one_ms / (world * n_ms), "speedup_vs_one_gpu": one_ms / n_ms,
                "bit_identical": same, "pixels_differing": differing,
                "partition": "regions", "strips_px": strips, "ms_per_rank": per_rank, "balancing_trials_ms_per_rank": balance_log,
                "parallelism": "region-parallel: vertical strips (bounds at multiples of 64 px, balanced by feedback: covered pixels of the previous render, rescaled by the ranks' measured times over three trial frames); each rank "
                               "builds its frame from the particles within 3.6 h of its strip's frustum (fr_set_region_partition) and its march "
                               "epilogue stores the pixels into the presenting GPU's image over NVLink peer memory (fr_ipc_*), no gather",
                "interleaved_tiles": {"tile": tile, "ms_per_frame": t_ms, "ms_per_frame_wall_with_barrier": t_wall,
                                      "efficiency_vs_one_gpu": one_ms / (world * t_ms), "bit_identical": t_same, "pixels_differing": t_diff,
                                      "parallelism": f"tile-parallel {tile}x{tile} interleaved over replicated frames, same peer-memory colour target"},
                "timing": "CUDA events on each rank's context stream around build + pre-pass + march of its share, max over ranks; "
                          "L2 flushed between frames; ms_per_frame_wall_with_barrier adds the NCCL barrier that tells the presenter "
                          "the image is complete"})
    return rec


def aniso_subrecord(fm, ctx, torch, stream, flush, xyz, d_ptr, npart, h, W, H, args, cam, settings, sm_mhz, steps=6):
    """the reference's DEFAULT path (EnableAnisotropy = true, AdvancedRenderer.cpp:23): PerPixel_Anisotropic
    (RayMarcher.cpp:346-423) on the same frame -- one frame at a time, per-stage device times, parity against the
    reference's own TUs, its CPU time beside it"""
    ctx.set_settings(fm.VisualizationSettings(EnableAnisotropy=True))
    log = []
    for k in range(steps + 2):
        with torch.cuda.stream(stream):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        ctx.build_frame_device(0, d_ptr, npart, h, 2.0)
        ctx.render_async(fm.FR_PASS_ALL)
        with torch.cuda.stream(stream):
            e1.record(stream)
        ctx.wait()
        torch.cuda.synchronize()
        if k >= 2:
            t = ctx.timings()
            t["frame_ms"] = e0.elapsed_time(e1)
            log.append(t)
    tim = {k: float(np.mean([t[k] for t in log])) for k in log[0]}
    cnt = ctx.counters()
    rec = {"workload": workload_string(args.config, npart, W, H).replace("isotropic", "anisotropic (k_n 0.5, k_r 2, k_s 2000, N_eps 1)"),
           "ms_per_frame": tim["frame_ms"], "value": W * H / (tim["frame_ms"] * 1e-3), "unit": UNIT,
           "stage_ms": {k: tim[k] for k in ("grid_ms", "depth_ms", "classify_ms", "march_first_ms", "march_long_ms")},
           "counters": {k: cnt[k] for k in ("covered_rays", "hit_rays", "ray_steps", "candidates", "neighbours", "queued_rays")}}
    ncu = load_ncu_stats(args.config + "_aniso", "k_march_long")
    if ncu.get("warp_instructions"):
        ipeak = 148 * 4 * sm_mhz * 1e6 / 1e9
        iach = ncu["warp_instructions"] / (tim["march_long_ms"] * 1e-3) / 1e9
        rec["roofline"] = {"bound": "issue", "kernel": "k_march_long<ANISO>", "achieved": iach, "peak": ipeak, "unit": "Ginst/s", "frac": iach / ipeak,
                           "traffic": ncu.get("dram_bytes"), "warp_instructions_per_launch": ncu["warp_instructions"],
                           "warp_instructions_source": f"profiles/{ncu.get('source', 'ncu_stats.json')}"}
    if not args.no_cpu_baseline:
        oracle_lib = checker()
        g_depth, g_pos, g_nrm, _ = ctx.download()
        s = oracle_lib.Settings(anisotropic=1)
        t0 = time.perf_counter()
        if oracle_lib.ref_available():
            ref = oracle_lib.Ref()
            ds = ref.dataset(xyz, h, 2.0)
            t1 = time.perf_counter()
            pos, nrm, march_s = ds.march(W, H, s, cam["inv_proj_view"], cam["position"], g_depth, threads=ref.lib.ref_hardware_threads())
            ds.close()
            kind, cores = "reference", ref.lib.ref_hardware_threads()
        else:
            orc = oracle_lib.Oracle()
            f = orc.frame(xyz, h, 2.0)
            t1 = time.perf_counter()
            pos, nrm, *_ = f.march(W, H, s, cam["inv_proj_view"], cam["position"], g_depth, want_band=False)
            march_s = time.perf_counter() - t1
            kind, cores = "port", orc.lib.fo_get_threads()
        total = time.perf_counter() - t0
        u = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
        diff = (u(g_pos) != u(pos)).any(-1) | (u(g_nrm) != u(nrm)).any(-1)
        rec["cpu_baseline"] = {"value": W * H / total, "unit": UNIT, "cores": cores, "kind": kind,
                               "sample": f"one full frame: Frame::Frame build {1e3 * (t1 - t0):.0f} ms + PerPixel_Anisotropic march {1e3 * march_s:.0f} ms"}
        rec["parity_checked"] = True
        rec["pixels_differing"] = int(diff.sum())
    ctx.set_settings(settings)
    return rec

# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    fm = importlib.import_module("bachelor-thesis_b200")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    n, W, H, h, dx = CONFIGS[args.config]
    cam = camera()
    cam_args = (cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
    tiles_mode = args.mode == "tiles" and world > 1
    # frames in flight per GPU; every lane has a host worker thread (see --wait)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    cores = os.cpu_count() or 8
    # frames in flight per GPU: 8 on one GPU (r03d/e: host -> host 0.257-0.268 ms per frame against 0.28-0.29 with 5 or 6 and
    # 0.265-0.279 with 9 or 10; the device-resident rate does not care: 0.225-0.229 from 4 to 12); 6 with several ranks
    # on one host, where the lane threads outnumber the cores
    lanes = 1 if tiles_mode else (args.lanes if args.lanes > 0 else (8 if world == 1 else 6))
    # Every step renders THE SAME synthetic frame (FRAME_T) -- the one the reference arm, the cpu_baseline leg and
    # tests/test_gpu_fullsize.py march -- out of n_frames separate device / pinned-host buffers, enough of them that
    # the particle INPUTS alone exceed the 126 MB L2 (the pipelined arm does not flush).  Frame-parallel: rank r
    # renders steps r, r + world, ... of that sequence
    n_frames = 1 if tiles_mode else max(4, min(24, -(-150_000_000 // (12 * n))))
    frame0 = fm.scenes.dam_break(n, h=h, dx=dx, t=FRAME_T)
    frames = [frame0.copy() for _ in range(n_frames)]
    n_actual = [len(f) for f in frames]
    settings = fm.VisualizationSettings(FastNormals=args.fast_normals)

    ctx = fm.Context(W, H, device=local)
    ctx.set_camera(*cam_args)
    if tiles_mode:
        ctx.set_tile_partition(rank, world, args.tile, args.tile)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    d_frames = [torch.from_numpy(f).to(dev) for f in frames]                    # inputs resident in HBM
    h_frames = [torch.from_numpy(f).pin_memory() for f in frames]               # e2e: pinned host inputs
    n_out = max(2, lanes)
    h_pos = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(n_out)]
    h_nrm = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(n_out)]
    h_rgba = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(n_out)]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)               # > 126 MB L2
    mg = importlib.import_module("bachelor-thesis_b200.multigpu")
    peer_mode = tiles_mode and args.gather == "peer"
    if peer_mode:
        # the presenting GPU's colour image is every rank's colour target (CUDA IPC, NVLink peer stores): no gather
        handle = [ctx.ipc_export_color() if rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            ctx.ipc_open_color_target(handle[0])
    elif tiles_mode:
        rgba_dev = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
        ctx.set_color_target(rgba_dev.data_ptr())
        owner = torch.from_numpy(mg.tile_owner_map(W, H, world, args.tile, args.tile)).to(dev)
    ctx.set_settings(settings)
    torch.cuda.synchronize()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- one frame at a time on one context: per-frame latency, per-stage device times, the roofline kernel -----
    def device_step(k):
        f = k % n_frames
        ctx.build_frame_device(0, d_frames[f].data_ptr(), n_actual[f], h, 2.0)
        ctx.render_async(fm.FR_PASS_ALL)
        if peer_mode:
            ctx.wait()
            dist.barrier()                   # every rank's tiles have landed in the presenter's image
        elif tiles_mode:
            ctx.wait()
            # exchange step of the tile-parallel path as a collective: RGBA tiles -> presenting GPU, merged there
            mg.gather_tiles(rgba_dev, rank, world, owner, dst=0)

    def timed_serial(step_fn, steps, warmup, sampler=None, stage_log=None):
        for k in range(warmup):
            step_fn(k)
        ctx.wait()
        launches_before = ctx.counters()["kernel_launches"]
        sync_all()
        if sampler:
            sampler.start()
        evs = []
        wall0 = time.perf_counter()
        for k in range(steps):
            if stage_log is not None and k > 0:
                stage_log.append(ctx.timings())                 # waits for step k-1 (it has to finish anyway)
            with torch.cuda.stream(stream):
                flush.zero_()                                   # L2 flush, outside the timed events
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            step_fn(warmup + k)
            with torch.cuda.stream(stream):
                e1.record(stream)
            evs.append((e0, e1))
        ctx.wait()
        if stage_log is not None:
            stage_log.append(ctx.timings())
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
        sync_all()
        clocks = sampler.stop() if sampler else None
        total_ms = max_over_ranks(float(sum(a.elapsed_time(b) for a, b in evs)))
        timed_serial.launches = ctx.counters()["kernel_launches"] - launches_before
        return total_ms / steps, wall, clocks

    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER") else None
    if sampler and "BENCH_SAMPLER_MS" not in os.environ:
        # a handful of samples inside the timed region (~0.3 ms per step), never faster than 1 kHz
        sampler.period = min(5.0, max(1.0, args.steps * 0.3 / 8.0)) * 1e-3
    stage_log = []
    ctx.wait()
    serial_ms, serial_wall, serial_clocks = timed_serial(device_step, args.steps, args.warmup,
                                                         sampler if tiles_mode else None, stage_log)
    cnt = ctx.counters()                                        # counters of the last step
    serial_launches = timed_serial.launches
    tim = {k: float(np.mean([t[k] for t in stage_log])) for k in stage_log[0]}
    # the same frames without the per-stage events (fr_set_stage_timing(0), what the RayMarcher shim runs): the
    # kernels of a frame then form chains of programmatic dependent launches, and the depth pre-pass runs on a second
    # stream beside the frame build (so latency_ms_per_frame is below the sum of stage_ms, which are timed one stage
    # after the other)
    serial_with_events_ms = serial_ms
    if not tiles_mode:
        ctx.set_stage_timing(False)
        serial_ms, _, _ = timed_serial(device_step, args.steps, args.warmup)
        ctx.set_stage_timing(True)

    # ---- the sequence: `lanes` frames in flight (fr_seq_*), K steps timed as one region --------------------------
    def timed_sequence(seq, submit, steps, warmup, sampler=None):
        lane_ctx = [seq.context(l) for l in range(lanes)]
        # every lane must have rendered before the clock starts (allocations, graph instantiation): at least two
        # frames per lane, whatever W says
        for k in range(max(warmup, 2 * lanes)):
            submit(k)
        seq.drain()
        before = sum(c.counters()["kernel_launches"] for c in lane_ctx)
        sync_all()
        if sampler:
            sampler.start()
        wall0 = time.perf_counter()
        seq.timer_begin()
        for k in range(steps):
            submit(warmup + k)
        ms = seq.timer_end()                                    # drains; CUDA events, latest lane
        wall = time.perf_counter() - wall0
        sync_all()
        clocks = sampler.stop() if sampler else None
        timed_sequence.launches = sum(c.counters()["kernel_launches"] for c in lane_ctx) - before
        return max_over_ranks(ms) / steps, wall, clocks

    if tiles_mode:
        ms_step, wall, clocks, kernel_launches_timed = serial_ms, serial_wall, serial_clocks, serial_launches
        if peer_mode:
            dist.barrier()
            if rank != 0:
                ctx.ipc_close_color_target()
        else:
            ctx.set_color_target(None)

        def e2e_tiles(k):
            ctx.upload_frame_ptr(0, h_frames[0].data_ptr(), n_actual[0], h, 2.0)
            ctx.render_async(fm.FR_PASS_ALL)
            ctx.download_ptrs(rgba=h_rgba[0].data_ptr())
        e2e_ms, _, _ = timed_serial(e2e_tiles, args.steps, 3)
        e2e_ref_ms = None
    else:
        seq = fm.Sequence(W, H, lanes=lanes, device=local)
        # lane workers wait inside the driver unless the host is oversubscribed ((lanes + 1) threads per rank)
        yielding = args.wait == "yield" or (args.wait == "auto" and (lanes + 1) * local_world > cores)
        seq.set_yielding(yielding)
        seq.set_camera(*cam_args)
        seq.set_settings(settings)

        def submit_device(k):
            f = k % n_frames
            seq.submit_ptrs(d_frames[f].data_ptr(), n_actual[f], h, 2.0, on_device=True)

        def submit_e2e(k):
            # what a host application calls per frame: particles (pinned host) -> finished colour image (pinned host)
            f = k % n_frames
            seq.submit_ptrs(h_frames[f].data_ptr(), n_actual[f], h, 2.0, rgba=h_rgba[k % n_out].data_ptr())

        def submit_e2e_refproto(k):
            # the reference's RayMarcher protocol: positions and normals also come back to the host (Prepare's buffers)
            f = k % n_frames
            o = k % n_out
            seq.submit_ptrs(h_frames[f].data_ptr(), n_actual[f], h, 2.0, positions=h_pos[o].data_ptr(),
                            normals=h_nrm[o].data_ptr(), rgba=h_rgba[o].data_ptr())

        ms_step, wall, clocks = timed_sequence(seq, submit_device, args.steps, args.warmup, sampler)
        kernel_launches_timed = timed_sequence.launches
        e2e_ms, _, _ = timed_sequence(seq, submit_e2e, args.steps, max(3, min(args.warmup, 6)))
        e2e_ref_ms, _, _ = timed_sequence(seq, submit_e2e_refproto, max(4, args.steps // 2), 3)

        # presentation path: the finished image of every frame goes into a frame ring on the presenting GPU
        # (rank 0) -- the shading epilogue of each rank's march stores its pixels there over NVLink peer memory, a flag
        # behind it says the frame is complete -- so the host path is the particle upload alone
        ring_slots = 2 * lanes
        present = {}
        if rank == 0:
            ring = fm.DeviceBuffer(world * ring_slots * W * H * 4, device=local)
            rflags = fm.DeviceBuffer(4 * world * ring_slots + 256, device=local)
            handles = [ring.export(), rflags.export()]
        else:
            handles = [None, None]
        if world > 1:
            dist.broadcast_object_list(handles, src=0)
            if rank != 0:
                ring = fm.DeviceBuffer(device=local, handle=handles[0])
                rflags = fm.DeviceBuffer(device=local, handle=handles[1])

        def submit_present(k):
            f = k % n_frames
            slot = rank * ring_slots + k % ring_slots
            seq.submit_ptrs(h_frames[f].data_ptr(), n_actual[f], h, 2.0, rgba_device=ring.ptr + slot * W * H * 4,
                            done_flag_device=rflags.ptr + 4 * slot, done_value=k + 1)

        present_ms, _, _ = timed_sequence(seq, submit_present, args.steps, max(3, min(args.warmup, 6)))
        sync_all()
        if rank == 0:
            got = torch.as_tensor(_DevPtr(ring.ptr, world * ring_slots * W * H * 4), device=dev).view(world * ring_slots, H, W, 4)
            fl = torch.as_tensor(_DevPtr(rflags.ptr, 4 * world * ring_slots), device=dev).view(torch.int32)
            same = all(bool(torch.equal(got[s], got[0])) for s in range(world * ring_slots))       # every buffer holds the same frame
            present = {"ms_per_step": present_ms, "ring_slots_per_rank": ring_slots, "all_slots_hold_the_frame": same,
                       "flags_set": int((fl > 0).sum().item()), "image_matches_host_arm": bool(np.array_equal(got[0].cpu().numpy(), h_rgba[0].numpy()))}
        sync_all()
        ring.close(); rflags.close()
        seq.close()
    ceiling, tiles_rec = None, None
    if not tiles_mode:
        ceiling = host_ceiling(torch, dist, dev, world, 12 * n_actual[0], W * H * 4)
        if not args.no_tiles:
            tiles_rec = tiles_subrecord(fm, torch, dist, dev, rank, world, local, cam_args, cam, args.tile)

    units = W * H * (1 if tiles_mode else world)          # rays per step over all ranks
    value = units / (ms_step * 1e-3)
    e2e_value = units / (e2e_ms * 1e-3)

    if rank == 0:
        peak, peak_src = load_peaks()
        npart = n_actual[(args.warmup + args.steps - 1) % n_frames]
        # per-kernel device times of the stages (CUDA events on the context stream, one frame at a time, L2 flushed)
        stages = {"grid build (5 kernels)": tim["grid_ms"], "depth pre-pass (7 kernels)": tim["depth_ms"],
                  "k_classify": tim["classify_ms"], "k_march_first": tim["march_first_ms"], "k_march_long": tim["march_long_ms"]}
        # dominant KERNEL (not stage): the depth stage is five kernels, the largest of which (k_depth_splat) is ~43% of the
        # stage in every committed ncu launch list (profiles/*_launches_C2.csv)
        dom = "k_march_first" if tim["march_first_ms"] >= 0.45 * tim["depth_ms"] else "depth pre-pass (7 kernels)"
        if dom == "k_march_first":
            # C_step x 16 B per density evaluation (the candidates of the 27-cell query, which any 27-cell
            # method incl. the CPU reference must examine) + per covered pixel 4 (depth) + 32 (pos, nrm) + 4 (rgba)
            alg = 16.0 * cnt["first_candidates"] + 40.0 * cnt["covered_rays"]
            dur_ms = tim["march_first_ms"]
            note = "16 B x candidates examined by the first sample of every covered ray + 40 B x covered pixels"
        else:
            alg = 16.0 * npart + 4.0 * W * H + 4.0 * cnt["covered_rays"] + 32.0 * npart
            dur_ms = tim["depth_ms"]
            note = "16 B x particles + 32 B x particles (splat records) + 4 B x pixels (clear) + 4 B x covered pixels"
        achieved = alg / (dur_ms * 1e-3) / 1e9
        ncu = load_ncu_stats(args.config, dom)
        sm_mhz_now = (clocks or {}).get("sm_mhz") or 1965.0
        hbm_alg = {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                   "algorithmic_bytes": alg, "algorithmic_bytes_def": note,
                   "note": "SURVEY 8(d)'s figure: the bytes a 27-cell method examines, as if every ray fetched its own candidates from "
                           "HBM.  Not a physical HBM fraction: the kernel's DRAM traffic is `traffic` (candidates are L1 / L2 hits)"}
        if ncu.get("warp_instructions"):
            # the binding resource (ncu: DRAM < 1 %, L2 < 5 % of peak, issue slots ~60-70 % busy in active cycles): warp
            # instructions issued per second against 4 schedulers x 148 SMs x SM clock
            ipeak = 148 * 4 * sm_mhz_now * 1e6 / 1e9
            iach = ncu["warp_instructions"] / (dur_ms * 1e-3) / 1e9
            roof = {"bound": "issue", "kernel": dom, "achieved": iach, "peak": ipeak, "unit": "Ginst/s", "frac": iach / ipeak,
                    "traffic": ncu.get("dram_bytes"),
                    "peak_source": f"4 warp schedulers x 148 SMs x {sm_mhz_now:.0f} MHz (SM clock sampled in this run)",
                    "warp_instructions_per_launch": ncu["warp_instructions"],
                    "warp_instructions_source": f"profiles/{ncu.get('source', 'ncu_stats.json')} (ncu smsp__inst_executed.sum of the committed capture of this kernel)"}
        else:
            roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu.get("dram_bytes"), "peak_source": peak_src}
        roof.update({"hbm_algorithmic": hbm_alg, "kernel_ms": dur_ms, "kernel_share_of_step": dur_ms / serial_with_events_ms, "stage_ms": stages,
                     "note": "kernel_ms is the kernel's CUDA-event time with one frame in flight and L2 flushed.  The path is bound by "
                             "instruction issue (FP32 without FMA contraction, correctly rounded division / sqrt sequences), not by HBM: "
                             "compulsory HBM traffic of a frame is ~0.1 GB (DESIGN.md 3.8)",
                     "ncu": ncu})
        # the march is a gather served by L1/L2, so the same algorithmic bytes also go against the L2 -> SM
        # bandwidth, measured live (fr_measure_l2_bandwidth: read-only streaming over an L2-resident 32 MB buffer)
        try:
            l2_peak = max(ctx.measure_l2_bandwidth(32, 40) for _ in range(3))
        except Exception:
            l2_peak = None
        if l2_peak:
            l2_actual = 32.0 * ncu["l2_sectors"] if ncu.get("l2_sectors") else None
            roof["l2"] = {"peak": l2_peak, "unit": "GB/s", "peak_source": "measured in this run (fr_measure_l2_bandwidth, 32 MB, ld.cg)",
                          "achieved_unshared": achieved, "frac_unshared": achieved / l2_peak,
                          "traffic": l2_actual,
                          "achieved_actual": (l2_actual / (dur_ms * 1e-3) / 1e9) if l2_actual else None,
                          "note": "unshared = as if every ray fetched its own candidates (the CPU's access pattern); actual = ncu "
                                  "lts__t_sectors x 32 B per launch: the rays of a warp share their candidates through L1"}
        if dom == "k_march_first":
            # SURVEY 8(d) algorithmic flops: 9 per candidate (distance test) + 10 per in-range kernel evaluation + 25 per
            # in-range gradient term (the first sample carries the normal's gradient sum)
            sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
            flops = 9.0 * cnt["first_candidates"] + 35.0 * cnt["neighbours"] * (cnt["first_candidates"] / max(cnt["candidates"], 1))
            peak_tf = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
            roof["fp32"] = {"achieved": flops / (dur_ms * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                            "frac": flops / (dur_ms * 1e-3) / 1e12 / peak_tf,
                            "note": "algorithmic flops against the FMA peak at the sampled SM clock; the kernel cannot use FMA (one IEEE "
                                    "rounding per operation, as the reference) and its divisions / square roots are correctly rounded "
                                    "multi-instruction sequences, so issue slots (fp32_issue_frac) are the meaningful measure"}
        if ncu.get("warp_instructions"):
            roof["fp32_issue_frac"] = ncu["warp_instructions"] / (148 * 4 * sm_mhz_now * 1e6 * dur_ms * 1e-3)
        cpu = None
        parity = {"parity_checked": False, "pixels_differing": None}
        if world == 1 and not args.no_cpu_baseline:
            step, kind, nthreads, _, ref_depth = reference_step_fn(args.config)
            ts = [step() for _ in range(3)]
            best = min(t[0] for t in ts)
            cpu = {"value": W * H / best, "unit": UNIT, "cores": nthreads, "kind": kind,
                   "sample": f"best of 3 full frames of {args.config} on the host cores: Frame::Frame build "
                             f"{1e3 * min(t[1] for t in ts):.0f} ms + march {1e3 * min(t[2] for t in ts):.0f} ms "
                             "(depth image precomputed; neighbour search = stand-in for the un-vendored CompactNSearch fork)"}
            if kind == "reference":
                tp = step(pool=True)
                cpu["reference_thread_pool"] = {"ms_per_step": 1e3 * tp[0], "march_ms": 1e3 * tp[2], "value": W * H / tp[0],
                                                "note": "the march on the reference's own ThreadPool (hardware_concurrency() - 1 workers, "
                                                        "one atomic per pixel, ThreadPool.cpp:38-55)"}
            # parity of the frame just timed: this context's images against the CPU arm's, bit for bit
            ctx.build_frame_device(0, d_frames[0].data_ptr(), n_actual[0], h, 2.0)
            ctx.render(fm.FR_PASS_ALL)
            g_depth, g_pos, g_nrm, _ = ctx.download()
            u = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
            ref_pos, ref_nrm = ts[0][3], ts[0][4]
            diff = (u(g_pos) != u(ref_pos)).any(-1) | (u(g_nrm) != u(ref_nrm)).any(-1)
            parity = {"parity_checked": True, "pixels_differing": int(diff.sum()),
                      "depth_pixels_differing": int((u(g_depth) != u(ref_depth)).sum()),
                      "parity_against": f"{kind}: hit mask, positions and normals bit for bit ({'oracle/_ref = the reference TUs' if kind == 'reference' else 'oracle C port'}); "
                                        "depth image against the oracle's restatement of depth.vert/frag",
                      "image_sha256_16": image_digest(g_pos, g_nrm)}
        aniso = None
        if world == 1 and not args.no_aniso and not tiles_mode:
            aniso = aniso_subrecord(fm, ctx, torch, stream, flush, frames[0], d_frames[0].data_ptr(), n_actual[0], h, W, H, args, cam, settings,
                                    (clocks or {}).get("sm_mhz") or 1965.0)
        if tiles_mode:
            par = (f"tile-parallel {args.tile}x{args.tile} interleaved, every rank's shading epilogue stores its tiles into the presenting GPU's image "
                   "over NVLink peer memory (CUDA IPC), barrier" if peer_mode else f"tile-parallel {args.tile}x{args.tile} interleaved + NCCL gather")
            l2 = "flushed between steps (512 MiB memset outside the timed events)"
        else:
            par = (f"frame-parallel x{world}, {lanes} frames in flight per GPU (fr_seq_*), lane workers "
                   f"{'poll pinned memory and yield' if yielding else 'wait in the driver'} ({cores} host cores)")
            l2 = (f"inputs larger than L2: {n_frames} separate buffers of {12 * npart / 1e6:.1f} MB cycled "
                  f"({n_frames * 12 * npart / 1e6:.0f} MB of particles, + {lanes} x {W * H * 40 / 1e6:.0f} MB of images; every buffer holds the same frame); "
                  "latency_ms_per_frame / stage_ms: L2 flushed between steps (512 MiB memset outside the timed events)")
        e2e = {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(12 * npart), "d2h_bytes_per_step": int(W * H * 4),
               "path": ("fr_upload_frame(host xyz) -> fr_render_async(ALL) -> fr_download(rgba), pinned host buffers" if tiles_mode else
                        f"fr_seq_submit(host xyz -> host rgba), pinned host buffers, {lanes} frames in flight")}
        if ceiling:
            e2e["host_ceiling"] = ceiling
            e2e["host_ceiling_gbs"] = ceiling["gbs_all_ranks"]
            e2e["fraction_of_host_ceiling"] = ceiling["ms_per_frame_pair"] / e2e_ms
        if not tiles_mode and present:
            e2e["present"] = {"value": units / (present["ms_per_step"] * 1e-3), **present,
                              "h2d_bytes_per_step": int(12 * npart), "d2h_bytes_per_step": 0,
                              "path": "fr_seq_submit(host xyz -> colour image in a frame ring on the presenting GPU over NVLink peer memory "
                                      "(fr_ipc_open_buffer), completion flag behind it); what the reference's swapchain presentation needs -- "
                                      "no image crosses PCIe"}
        if e2e_ref_ms is not None:
            e2e["reference_protocol"] = {"value": units / (e2e_ref_ms * 1e-3), "ms_per_step": e2e_ref_ms,
                                         "d2h_bytes_per_step": int(W * H * (16 + 16 + 4)),
                                         "path": "same, plus positions and normals copied back as RayMarcher::Prepare's host buffers expect"}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if tiles_mode else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.config, npart, W, H),
                       "step": "grid build + depth pre-pass + march/normals/shade of one frame, particles resident in HBM",
                       "parallelism": par, "l2": l2,
                       "normals": "fast (FMA + approximate reciprocal, ~1e-6)" if args.fast_normals else "bit-exact with the reference",
                       "latency_ms_per_frame": serial_ms, "latency_with_stage_events_ms": serial_with_events_ms, "host_cores": cores,
                       "stage_ms": tim, "counters": {k: cnt[k] for k in ("covered_rays", "hit_rays", "ray_steps", "candidates",
                                                                         "neighbours", "skip_iterations", "early_exits")},
                       "covered_rays_per_s": cnt["covered_rays"] * (1 if tiles_mode else world) / (ms_step * 1e-3),
                       "ray_steps_per_s": cnt["ray_steps"] * (1 if tiles_mode else world) / (ms_step * 1e-3),
                       "frames_per_s": (1 if tiles_mode else world) / (ms_step * 1e-3),
                       "wall_s_timed_region": wall},
            "clocks": clocks,
            "parity_checked": parity["parity_checked"], "pixels_differing": parity["pixels_differing"], "parity": parity,
            "e2e": e2e,
            "gpu_launches": int(kernel_launches_timed),
            "roofline": roof,
            "cpu_baseline": cpu,
            "tiles": tiles_rec,
            "aniso": aniso,
        }
        print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default="frames", choices=["frames", "tiles"])
    ap.add_argument("--lanes", type=int, default=0, help="frames in flight per GPU (fr_seq_create); 1 = one frame at a time; 0 = 8 on one GPU, 6 per rank otherwise")
    ap.add_argument("--wait", default="auto", choices=["auto", "spin", "yield"], help="how lane workers wait for the GPU (fr_seq_set_yielding)")
    ap.add_argument("--tile", type=int, default=128, help="--mode tiles: partition tile size in pixels (multiple of 64)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="--mode tiles: peer = ranks render into the presenter's image over NVLink peer memory; nccl = gather collective")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tiles", action="store_true", help="skip the C3 tile-parallel sub-record")
    ap.add_argument("--no-aniso", action="store_true", help="skip the anisotropic (reference default path) sub-record")
    ap.add_argument("--fast-normals", action="store_true", help="fr_settings.fast_normals (default: normals bit-exact)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
