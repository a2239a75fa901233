"""background_fast against shade_pixel: FLUIDMARCH_BGFAST=0 (exact everywhere) and the default must give the same RGBA
for several cameras and sizes; prints the number of differing pixels (must be 0)."""
import importlib, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) > 1:
    fm = importlib.import_module("bachelor-thesis_b200")
    from conftest import golden_camera
    xyz = fm.scenes.dam_break(30000, t=0.6)
    out = {}
    for cam_name in ("camera_default_16x9", "camera_close_16x9"):
        for W, H in ((1920, 1080), (1280, 720), (333, 187)):
            cam = golden_camera(cam_name)
            c = fm.Context(W, H)
            c.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
            c.upload_frame(0, xyz, 0.1, 2.0)
            c.render(fm.FR_PASS_ALL)
            out[f"{cam_name}_{W}x{H}"] = c.download()[3]
            c.close()
    np.savez(sys.argv[1], **out)
else:
    for mode in ("0", "1"):
        subprocess.check_call([sys.executable, __file__, f"/tmp/bg_{mode}.npz"], env=dict(os.environ, FLUIDMARCH_BGFAST=mode))
    a, b = np.load("/tmp/bg_0.npz"), np.load("/tmp/bg_1.npz")
    for k in a.files:
        print(k, "pixels differing:", int((a[k] != b[k]).any(axis=-1).sum()), "of", a[k].shape[0] * a[k].shape[1])
