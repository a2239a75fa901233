"""How long the frame build takes on degenerate input (crowded cells are sorted by one CTA each, k_cell_order):
    python tools/crowded_probe.py
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fm = importlib.import_module("bachelor-thesis_b200")


def run(name, xyz, h):
    ctx = fm.Context(64, 64)
    for k in range(3):
        t0 = time.perf_counter()
        ctx.upload_frame(0, xyz, h, 2.0)
        info = ctx.frame_info(0)
        dt = (time.perf_counter() - t0) * 1e3
    g = ctx.download_frame(0)
    cs = g["cell_start"].astype(np.int64)
    idx = g["sorted_index"].astype(np.int64)
    ok = np.array_equal(np.sort(idx), np.arange(len(xyz)))
    inner = np.ones(len(idx), bool)
    inner[cs[1:-1][cs[1:-1] < len(idx)]] = False
    inner[0] = False
    asc = bool((np.diff(idx)[inner[1:]] > 0).all())
    print(f"{name}: n {len(xyz)} h {h} largest cell {int(np.diff(cs).max())} upload+build wall {dt:.2f} ms grid_ms {ctx.timings()['grid_ms']:.3f} permutation {ok} ascending {asc}", flush=True)
    ctx.close()


def main():
    rng = np.random.default_rng(1)
    run("one cell, 1M", rng.uniform(0.001, 0.049, (1_000_000, 3)).astype(np.float32), 0.1)
    run("one cell, 4M", rng.uniform(0.001, 0.049, (4_000_000, 3)).astype(np.float32), 0.1)
    xyz = fm.scenes.dam_break(1_000_000, t=0.6)
    run("C2 scene, h = 1.0", xyz, 1.0)
    run("C2 scene, h = 0.1", xyz, 0.1)


if __name__ == "__main__":
    main()
