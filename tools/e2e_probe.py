"""Where does the host->host sequence lose time against the device-resident one?  ms per frame (CUDA events) at C2 for
every combination of {particles on host, on device} x {no download, RGBA download}.   python tools/e2e_probe.py [lanes]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
fm = importlib.import_module("bachelor-thesis_b200")
from conftest import golden_camera

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 6
W, H = 1920, 1080
cam = golden_camera("camera_default_16x9")
frames = [fm.scenes.dam_break(1_000_000, t=0.45 + 0.02 * k) for k in range(13)]
h = [torch.from_numpy(f).pin_memory() for f in frames]
d = [torch.from_numpy(f).cuda() for f in frames]
out = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(lanes)]
seq = fm.Sequence(W, H, lanes=lanes)
seq.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
seq.set_settings(fm.VisualizationSettings())
for host_in in (False, True):
    for download in (False, True):
        def submit(k):
            f = k % 13
            src = h[f] if host_in else d[f]
            seq.submit_ptrs(src.data_ptr(), len(frames[f]), 0.1, 2.0, on_device=not host_in, rgba=out[k % lanes].data_ptr() if download else 0)
        for k in range(12): submit(k)
        seq.drain()
        res = []
        for rep in range(3):
            seq.timer_begin()
            for k in range(200): submit(k)
            res.append(seq.timer_end() / 200)
        print(f"lanes {lanes}  particles {'host' if host_in else 'device'}  download {'rgba' if download else 'none'}: " + " ".join(f"{r:.4f}" for r in res), flush=True)
seq.close()
