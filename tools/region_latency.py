"""One GPU plays every rank of a region-parallel C3 frame in turn, without per-stage events (as bench.py's one-frame arm
runs it): device time of build + pre-pass + march of each strip, and whether the strip's pixels equal the whole frame's.
    [FLUIDMARCH_OVERLAP_REGION=0] python tools/region_latency.py [world]"""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
fm = importlib.import_module("bachelor-thesis_b200")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n, W, H, h, dx = bench.CONFIGS["C3"]
xyz = fm.scenes.dam_break(n, h=h, dx=dx, t=bench.FRAME_T)
cam = bench.camera()
dev = torch.device("cuda:0")
d_xyz = torch.from_numpy(np.ascontiguousarray(xyz)).to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ctx = fm.Context(W, H)
ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
ctx.set_settings(fm.VisualizationSettings())
ctx.set_stage_timing(False)
stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)


def run(frames=10):
    ms = []
    for k in range(frames + 3):
        with torch.cuda.stream(stream):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        ctx.build_frame_device(0, d_xyz.data_ptr(), len(xyz), h, 2.0)
        ctx.render_async(fm.FR_PASS_ALL)
        with torch.cuda.stream(stream):
            e1.record(stream)
        ctx.wait()
        if k >= 3:
            ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms))


whole = run()
full = ctx.download()
print("whole frame", round(whole, 4), flush=True)
cov = (full[0] != 1.0).sum(axis=0).astype(np.float64)
cov = np.pad(cov, (0, (-len(cov)) % 64)).reshape(-1, 64).sum(axis=1)
times = []
for k, (x0, x1) in enumerate(bench.balanced_strips(cov, W, world)):
    ctx.set_region_partition(x0, 0, x1, H)
    t = run()
    got = ctx.download()
    same = all(np.array_equal(np.ascontiguousarray(g[:, x0:x1]).view(np.uint8), np.ascontiguousarray(w[:, x0:x1]).view(np.uint8)) for g, w in zip(got, full))
    times.append(t)
    print(f"strip {k} [{x0},{x1}) {t:.4f} ms  identical {same}", flush=True)
print("max strip", round(max(times), 4), "efficiency", round(whole / world / max(times), 3))
