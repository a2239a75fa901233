"""One frame at a time on one context (BASELINE config C2 or C3): device time per frame between CUDA events on the
context's stream, L2 flushed between frames, with and without the per-stage events.

    [FLUIDMARCH_PDL=0] [FLUIDMARCH_LIB=...] python tools/latency_probe.py [C2|C3] [frames]
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fm = importlib.import_module("bachelor-thesis_b200")

CONFIGS = {"C1": (64_000, 1280, 720, 0.1, None), "C2": (1_000_000, 1920, 1080, 0.1, None), "C3": (4_000_000, 3840, 2160, 0.063, 0.0315)}


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    n, W, H, h, dx = CONFIGS[cfg]
    xyz = fm.scenes.dam_break(n, h=h, dx=dx, t=0.6)
    if os.environ.get("SHUFFLE") == "1":                      # particle order as a simulation leaves it: arbitrary
        xyz = xyz[np.random.default_rng(0).permutation(len(xyz))]
    cam = fm.camera.reference_default_camera()
    dev = torch.device("cuda:0")
    d_xyz = torch.from_numpy(np.ascontiguousarray(xyz)).to(dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    ctx = fm.Context(W, H)
    ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
    ctx.set_settings(fm.VisualizationSettings())
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    out = {}
    for timing in (True, False, True, False):
        ctx.set_stage_timing(timing)
        evs = []
        for k in range(frames + 5):
            with torch.cuda.stream(stream):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            ctx.build_frame_device(0, d_xyz.data_ptr(), len(xyz), h, 2.0)
            ctx.render_async(fm.FR_PASS_ALL)
            with torch.cuda.stream(stream):
                e1.record(stream)
            ctx.wait()
            if k >= 5:
                evs.append(e0.elapsed_time(e1))
        out.setdefault("timed" if timing else "untimed", []).append(round(float(np.mean(evs)), 4))
    print(cfg, "pdl", os.environ.get("FLUIDMARCH_PDL", "1"), out, flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
