"""Profiling driver: a few whole steps (grid build -> depth -> march/shade) of one config through the C ABI,
no torch, no CPU baseline -- the command ncu wraps (see profiles/README.md).

    python tools/prof_step.py [C1|C2|C3] [steps]
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
fm = importlib.import_module("bachelor-thesis_b200")
from conftest import golden_camera  # noqa: E402

CONFIGS = {"C1": (64_000, 1280, 720, 0.1, None), "C2": (1_000_000, 1920, 1080, 0.1, None),
           "C3": (4_000_000, 3840, 2160, 0.063, 0.0315)}


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n, W, H, h, dx = CONFIGS[cfg]
    xyz = fm.scenes.dam_break(n, h=h, dx=dx)
    cam = golden_camera("camera_default_16x9")
    ctx = fm.Context(W, H)
    ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
    ctx.set_settings(fm.VisualizationSettings())
    if os.environ.get("FR_ASYNC") == "0":
        ctx.set_async_build(False)          # every build waits once for the device; the march takes the view by value
    log = []
    for _ in range(steps):
        ctx.upload_frame(0, xyz, h, 2.0)
        ctx.render(fm.FR_PASS_ALL)
        log.append(ctx.timings())
    tail = log[min(2, len(log) - 1):]                    # the first steps allocate
    mean = {k: round(sum(t[k] for t in tail) / len(tail), 4) for k in tail[0]}
    print(cfg, mean, ctx.counters())
    ctx.close()


if __name__ == "__main__":
    main()
