"""the colour image of two builds must be byte-identical (tools/ab check for exact-arithmetic substitutions)"""
import importlib, os, subprocess, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    fm = importlib.import_module("bachelor-thesis_b200")
    from conftest import golden_camera
    for camn in ("camera_default_16x9", "camera_close_16x9", "camera_orbit_a_16x9", "camera_orbit_b_16x9"):
        cam = golden_camera(camn)
        ctx = fm.Context(1280, 720)
        ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
        ctx.set_settings(fm.VisualizationSettings())
        ctx.upload_frame(0, fm.scenes.dam_break(64000), 0.1, 2.0)
        ctx.render(fm.FR_PASS_ALL)
        np.save(sys.argv[1] + "_" + camn + ".npy", ctx.download(False, False, False, True)[3])
        ctx.close()
else:
    for v in ("old", "new"):
        subprocess.check_call([sys.executable, __file__, "/tmp/rgba_" + v], env=dict(os.environ, FLUIDMARCH_AB="1", FLUIDMARCH_LIB=os.path.join(ROOT, "build_variants", v, "libfluidmarch.so")))
    for camn in ("camera_default_16x9", "camera_close_16x9", "camera_orbit_a_16x9", "camera_orbit_b_16x9"):
        a, b = np.load(f"/tmp/rgba_old_{camn}.npy"), np.load(f"/tmp/rgba_new_{camn}.npy")
        print(camn, "identical" if np.array_equal(a, b) else f"DIFFERENT in {(a != b).any(-1).sum()} pixels", len(np.unique(a.reshape(-1, 4), axis=0)), "colours")
