"""One GPU plays every rank of a region-parallel C3 frame in turn: per-stage device times of each strip
(python tools/region_probe.py [world])"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
fm = importlib.import_module("bachelor-thesis_b200")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n, W, H, h, dx = bench.CONFIGS["C3"]
xyz = fm.scenes.dam_break(n, h=h, dx=dx, t=bench.FRAME_T)
cam = bench.camera()
ctx = fm.Context(W, H)
ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
ctx.set_settings(fm.VisualizationSettings())


def run(label):
    log = []
    for _ in range(8):
        ctx.upload_frame(0, xyz, h, 2.0)
        ctx.render(fm.FR_PASS_ALL)
        log.append(ctx.timings())
    tail = log[2:]
    m = {k: round(sum(t[k] for t in tail) / len(tail), 4) for k in ("grid_ms", "depth_ms", "classify_ms", "march_first_ms", "march_long_ms")}
    c = ctx.counters()
    print(label, m, "sum", round(sum(m.values()), 4), "covered", c["covered_rays"], "queued", c["queued_rays"], flush=True)


run("whole frame")
import numpy as np  # noqa: E402
cov = (ctx.download(True, False, False, False)[0] != 1.0).sum(axis=0).astype(np.float64)
cov = np.pad(cov, (0, (-len(cov)) % 64)).reshape(-1, 64).sum(axis=1)
for k, (x0, x1) in enumerate(bench.balanced_strips(cov, W, world)):
    ctx.set_region_partition(x0, 0, x1, H)
    run(f"strip {k} [{x0},{x1})")
