"""does PCIe traffic by itself slow the kernels?  device-resident sequence (no copies of its own) with and without a
background thread that keeps both copy engines busy."""
import importlib, os, sys, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
fm = importlib.import_module("bachelor-thesis_b200")
from conftest import golden_camera
lanes = 6
W, H = 1920, 1080
cam = golden_camera("camera_default_16x9")
frames = [fm.scenes.dam_break(1_000_000, t=0.45 + 0.02 * k) for k in range(13)]
d = [torch.from_numpy(f).cuda() for f in frames]
seq = fm.Sequence(W, H, lanes=lanes)
seq.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
seq.set_settings(fm.VisualizationSettings())
def run(tag):
    for k in range(12): seq.submit_ptrs(d[k % 13].data_ptr(), len(frames[k % 13]), 0.1, 2.0, on_device=True)
    seq.drain()
    res = []
    for rep in range(3):
        seq.timer_begin()
        for k in range(300): seq.submit_ptrs(d[k % 13].data_ptr(), len(frames[k % 13]), 0.1, 2.0, on_device=True)
        res.append(seq.timer_end() / 300)
    print(tag, " ".join(f"{r:.4f}" for r in res), flush=True)
run("no background copies:        ")
stop = False
h_in = torch.empty(12_060_000, dtype=torch.uint8).pin_memory(); h_out = torch.empty(8_294_400, dtype=torch.uint8).pin_memory()
d_in = torch.empty_like(h_in, device="cuda"); d_out = torch.empty_like(h_out, device="cuda")
def bg(direction):
    s = torch.cuda.Stream()
    n = 0
    with torch.cuda.stream(s):
        while not stop:
            if direction == 0: d_in.copy_(h_in, non_blocking=True)
            else: h_out.copy_(d_out, non_blocking=True)
            n += 1
            if n % 8 == 0: s.synchronize()
for dirs, tag in (((0,), "background H2D stream:       "), ((1,), "background D2H stream:       "), ((0, 1), "background H2D + D2H streams:")):
    stop = False
    ts = [threading.Thread(target=bg, args=(x,)) for x in dirs]
    for t in ts: t.start()
    run(tag)
    stop = True
    for t in ts: t.join()
seq.close()
