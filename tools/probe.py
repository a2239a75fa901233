"""GPU probe: stage timings + counters for a config (diagnostics, not the bench)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
fm = importlib.import_module("bachelor-thesis_b200")
from conftest import golden_camera

def run(n, W, H, reps=5, h=0.1, dx=None, aniso=False):
    xyz = fm.scenes.dam_break(n, h=h, dx=dx)
    cam = golden_camera("camera_default_16x9")
    ctx = fm.Context(W, H)
    ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
    ctx.set_settings(fm.VisualizationSettings(EnableAnisotropy=aniso))
    for r in range(reps):
        t0 = time.time()
        ctx.upload_frame(0, xyz, h, 2.0)
        t1 = time.time()
        ctx.render(fm.FR_PASS_ALL)
        t2 = time.time()
        out = ctx.download()
        t3 = time.time()
        print(n, W, H, "aniso" if aniso else "iso", "rep", r, {k: round(v, 4) for k, v in ctx.timings().items()},
              "wall upload+build %.2fms render %.2fms download %.2fms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
    print(ctx.counters(), ctx.frame_info(0), flush=True)
    ctx.close()

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "aniso":
        run(64000, 1280, 720, aniso=True)
        run(1000000, 1920, 1080, aniso=True)
        sys.exit(0)
    run(64000, 1280, 720)
    run(1000000, 1920, 1080)
    run(4000000, 3840, 2160, reps=3, h=0.063, dx=0.0315)
