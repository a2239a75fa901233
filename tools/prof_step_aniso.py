"""like prof_step.py with EnableAnisotropy (the reference's default path): python tools/prof_step_aniso.py [C1|C2] [steps]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
fm = importlib.import_module("bachelor-thesis_b200")
from conftest import golden_camera  # noqa: E402
CONFIGS = {"C1": (64_000, 1280, 720, 0.1, None), "C2": (1_000_000, 1920, 1080, 0.1, None)}
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, W, H, h, dx = CONFIGS[cfg]
xyz = fm.scenes.dam_break(n, h=h, dx=dx)
cam = golden_camera("camera_default_16x9")
ctx = fm.Context(W, H)
ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
ctx.set_settings(fm.VisualizationSettings(EnableAnisotropy=True))
for _ in range(steps):
    ctx.upload_frame(0, xyz, h, 2.0)
    ctx.render(fm.FR_PASS_ALL)
print(cfg, "aniso", ctx.timings(), ctx.counters())
ctx.close()
