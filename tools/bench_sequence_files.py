"""Sequence playback from .bgeo files on disk: frames/s of  file -> GPU decode -> grid build -> depth ->
march -> shade -> RGBA on the host,  `lanes` frames in flight, beside the reference's loader (Partio::read +
Dataset::ReadFile through oracle/_ref) on the same files.  Diagnostics, not the headline bench.

    python tools/bench_sequence_files.py [n_particles] [frames] [lanes] [extras]
"""
import importlib
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
fm = importlib.import_module("bachelor-thesis_b200")
import oracle_lib  # noqa: E402
from conftest import golden_camera  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    extras = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    W, H = 1920, 1080
    d = tempfile.mkdtemp(prefix="fm_seq_")
    ref = oracle_lib.Ref() if oracle_lib.ref_available() else None
    try:
        for i in range(1, frames + 1):
            xyz = fm.scenes.dam_break(n, t=0.45 + 0.3 * (i - 1) / max(frames - 1, 1))
            path = os.path.join(d, f"ParticleData_Fluid_{i}.bgeo")
            if extras and ref is not None and ref.has_partio:
                ref.partio_write(path, xyz, extras, False)      # velocity / id / density in the records, like a simulator export
            else:
                fm.bgeo_write(path, xyz)
        count = fm.dataset_count(os.path.join(d, "ParticleData_Fluid_"), ".bgeo")
        info = fm.bgeo_probe(os.path.join(d, "ParticleData_Fluid_1.bgeo"))
        cam = golden_camera("camera_default_16x9")
        seq = fm.Sequence(W, H, lanes=lanes)
        seq.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
        seq.set_settings(fm.VisualizationSettings())
        import torch
        outs = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(lanes)]
        paths = [os.path.join(d, f"ParticleData_Fluid_{i}.bgeo") for i in range(1, count + 1)]

        def play(reps):
            k = 0
            for _ in range(reps):
                for p in paths:
                    seq.submit_ptrs(0, 0, 0.1, 2.0, bgeo_path=p, rgba=outs[k % lanes].data_ptr())
                    k += 1
            seq.drain()
            return k
        play(1)                                             # warm-up: page cache, allocations
        t0 = time.perf_counter()
        k = play(3)
        dt = time.perf_counter() - t0
        out = {"what": "sequence playback from .bgeo files", "particles": int(info["num_particles"]), "record_words": int(info["record_words"]),
               "file_mb": info["file_bytes"] / 1e6, "frames": k, "lanes": lanes, "frames_per_s": k / dt, "ms_per_frame": 1e3 * dt / k,
               "file_gb_per_s": k * info["file_bytes"] / dt / 1e9}
        seq.close()
        if ref is not None and ref.has_partio:
            t0 = time.perf_counter()
            for p in paths[:4]:
                ref.partio_read(p)
            out["partio_read_ms_per_frame"] = 1e3 * (time.perf_counter() - t0) / len(paths[:4])
        t0 = time.perf_counter()
        for p in paths[:4]:
            fm.bgeo_read(p)
        out["host_decode_ms_per_frame"] = 1e3 * (time.perf_counter() - t0) / len(paths[:4])
        print(json.dumps(out))
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
