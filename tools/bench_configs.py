"""The BASELINE.json configs that are not the headline bench line, measured through the public API on one GPU:

    python tools/bench_configs.py c5     smoothing-radius / step-size sweep at 1M particles, 1080p (neighbour count 20..120)
    python tools/bench_configs.py c4     240-frame dam-break sequence, 1M particles, 1080p, host particles -> host RGBA
    python tools/bench_configs.py c3     4M particles at 3840x2160 on one GPU
    python tools/bench_configs.py aniso  C1 / C2 with the reference's default (anisotropic) kernel

One JSON line per measurement.  Times are CUDA-event device times (fr_seq_timer_* / fr_get_timings).
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
fm = importlib.import_module("bachelor-thesis_b200")
from conftest import golden_camera  # noqa: E402

CAM = golden_camera("camera_default_16x9")
CAM_ARGS = (CAM["view"], CAM["proj"], CAM["inv_proj_view"], CAM["position"], CAM["system"].reshape(3, 3)[2])


def one_context(xyz, W, H, h, settings, reps=12):
    """median per-stage times of `reps` frames, one at a time"""
    ctx = fm.Context(W, H)
    ctx.set_camera(*CAM_ARGS)
    ctx.set_settings(settings)
    rows = []
    for r in range(reps + 2):
        ctx.upload_frame(0, xyz, h, 2.0)
        ctx.render(fm.FR_PASS_ALL)
        if r >= 2:
            rows.append(ctx.timings())
    cnt = ctx.counters()
    ctx.close()
    med = {k: float(np.median([t[k] for t in rows])) for k in rows[0]}
    med["frame_ms"] = med["grid_ms"] + med["depth_ms"] + med["march_ms"]
    return med, cnt


def pipelined(frames, W, H, h, settings, lanes=4, steps=60):
    import torch
    dev = [torch.from_numpy(f).cuda() for f in frames]
    seq = fm.Sequence(W, H, lanes=lanes)
    seq.set_camera(*CAM_ARGS)
    seq.set_settings(settings)
    for k in range(8):
        seq.submit_ptrs(dev[k % len(dev)].data_ptr(), len(frames[k % len(dev)]), h, 2.0, on_device=True)
    seq.timer_begin()
    for k in range(steps):
        seq.submit_ptrs(dev[k % len(dev)].data_ptr(), len(frames[k % len(dev)]), h, 2.0, on_device=True)
    ms = seq.timer_end() / steps
    seq.close()
    return ms


def c5():
    dx, W, H = 0.05, 1920, 1080
    for ratio in (1.68, 1.93, 2.29, 2.67, 3.06):
        h = float(np.float32(ratio * dx))
        xyz = fm.scenes.dam_break(1_000_000, h=h, dx=dx)
        for div in (20, 11, 5):
            step = float(np.float32(h / div))
            s = fm.VisualizationSettings(StepSize=step, MaxSteps=int(round(1.152 / step)))
            med, cnt = one_context(xyz, W, H, h, s, reps=8)
            ms = pipelined([xyz, xyz[::-1].copy()], W, H, h, s, steps=40) if div == 11 else None
            print(json.dumps({"config": "C5", "h_over_dx": ratio, "h": h, "step": f"h/{div}", "max_steps": s.MaxSteps, "particles": len(xyz),
                              "neighbours_per_sample": cnt["neighbours"] / max(cnt["ray_steps"], 1),
                              "candidates_per_sample": cnt["candidates"] / max(cnt["ray_steps"], 1),
                              "covered_rays": cnt["covered_rays"], "ray_steps": cnt["ray_steps"], "stage_ms": med,
                              "pipelined_ms_per_frame": ms}), flush=True)


def c4():
    import torch
    W, H, n_frames, lanes = 1920, 1080, 240, 4
    t0 = time.perf_counter()
    frames = [torch.from_numpy(fm.scenes.dam_break(1_000_000, t=k / (n_frames - 1))).pin_memory() for k in range(n_frames)]
    gen_s = time.perf_counter() - t0
    outs = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(lanes)]
    seq = fm.Sequence(W, H, lanes=lanes)
    seq.set_camera(*CAM_ARGS)
    seq.set_settings(fm.VisualizationSettings())
    for k in range(8):
        seq.submit_ptrs(frames[k].data_ptr(), frames[k].shape[0], 0.1, 2.0, rgba=outs[k % lanes].data_ptr())
    passes = []
    for rep in range(2):            # first pass: the tables grow with the spreading fluid (allocations); second: steady state
        seq.timer_begin()
        w0 = time.perf_counter()
        for k in range(n_frames):
            seq.submit_ptrs(frames[k].data_ptr(), frames[k].shape[0], 0.1, 2.0, rgba=outs[k % lanes].data_ptr())
        ms = seq.timer_end()
        wall = time.perf_counter() - w0
        passes.append(ms / n_frames)
    seq.close()
    print(json.dumps({"config": "C4 on one GPU", "frames": n_frames, "ms_per_frame_first_and_second_pass": passes, "particles_min_max": [min(f.shape[0] for f in frames), max(f.shape[0] for f in frames)],
                      "lanes": lanes, "device_ms_total": ms, "ms_per_frame": ms / n_frames, "frames_per_s": n_frames / (ms * 1e-3),
                      "wall_s": wall, "path": "pinned host particles -> fr_seq_submit -> pinned host RGBA", "generate_s": gen_s}), flush=True)


def c3():
    W, H, h, dx = 3840, 2160, 0.063, 0.0315
    xyz = fm.scenes.dam_break(4_000_000, h=h, dx=dx)
    med, cnt = one_context(xyz, W, H, h, fm.VisualizationSettings(), reps=6)
    ms = pipelined([xyz, xyz[::-1].copy(), xyz.copy()], W, H, h, fm.VisualizationSettings(), lanes=3, steps=20)
    print(json.dumps({"config": "C3 on one GPU", "particles": len(xyz), "stage_ms": med, "pipelined_ms_per_frame": ms,
                      "covered_rays": cnt["covered_rays"], "ray_steps": cnt["ray_steps"]}), flush=True)


def aniso():
    for name, (n, W, H) in {"C1": (64_000, 1280, 720), "C2": (1_000_000, 1920, 1080)}.items():
        xyz = fm.scenes.dam_break(n)
        med, cnt = one_context(xyz, W, H, 0.1, fm.VisualizationSettings(EnableAnisotropy=True), reps=5)
        print(json.dumps({"config": name + " anisotropic", "particles": len(xyz), "stage_ms": med, "covered_rays": cnt["covered_rays"],
                          "hit_rays": cnt["hit_rays"], "ray_steps": cnt["ray_steps"], "candidates": cnt["candidates"]}), flush=True)


if __name__ == "__main__":
    {"c5": c5, "c4": c4, "c3": c3, "aniso": aniso}[sys.argv[1]]()
