"""Reads the per-warp records of a -DFM_FIRST_PROFILE build (tools/build_variant.sh zz_fprof -DFM_FIRST_PROFILE):
when each warp of k_march_first finished, its longest tile.

    FLUIDMARCH_LIB=build_variants/zz_fprof/libfluidmarch.so FLUIDMARCH_AB=1 python tools/first_profile.py [C2|C3]
"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fm = importlib.import_module("bachelor-thesis_b200")

CONFIGS = {"C2": (1_000_000, 1920, 1080, 0.1, None), "C3": (4_000_000, 3840, 2160, 0.063, 0.0315)}


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
    n, W, H, h, dx = CONFIGS[cfg]
    xyz = fm.scenes.dam_break(n, h=h, dx=dx, t=0.6)
    cam = fm.camera.reference_default_camera()
    ctx = fm.Context(W, H)
    ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
    ctx.set_settings(fm.VisualizationSettings())
    for _ in range(4):
        ctx.upload_frame(0, xyz, h, 2.0)
        ctx.render(fm.FR_PASS_ALL)
        t = ctx.timings()
    lib = fm.load()
    buf = np.zeros(4 * 16384, dtype=np.uint64)
    lib.fr_debug_first_profile.restype = C.c_int
    lib.fr_debug_first_profile.argtypes = [C.c_void_p, C.c_int]
    assert lib.fr_debug_first_profile(buf.ctypes.data, buf.size) == 0
    r = buf.reshape(-1, 4)
    r = r[r[:, 0] != 0]
    start, fin = r[:, 0].astype(np.int64), r[:, 1].astype(np.int64)
    t0 = start.min()
    end = (fin - t0) / 1000.0
    longest = (r[:, 2] >> np.uint64(16)).astype(np.int64) / 1000.0
    tiles = (r[:, 2] & np.uint64(0xffff)).astype(np.int64)
    sk = (r[:, 3] >> np.uint64(32)).astype(np.int64)
    print(cfg, "march_first_ms", t["march_first_ms"], "warps", len(r), "tiles", tiles.sum())
    print("warp start spread us:", (start.max() - t0) / 1000.0)
    print("finish us: min %.1f p10 %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f mean %.1f" % (
        end.min(), *np.percentile(end, [10, 50, 90, 99]), end.max(), end.mean()))
    print("tiles per warp: min %d mean %.2f max %d" % (tiles.min(), tiles.mean(), tiles.max()))
    print("longest tile per warp us: p50 %.1f p90 %.1f p99 %.1f max %.1f" % (*np.percentile(longest, [50, 90, 99]), longest.max()))
    order = np.argsort(-end)[:12]
    for i in order:
        print("  late warp: finish %.1f us, tiles %d, longest tile %.1f us (max lane skips %d, tile y %d x %d)" % (
            end[i], tiles[i], longest[i], sk[i], (int(r[i, 3]) >> 16) & 0xffff, int(r[i, 3]) & 0xffff))
    # skips of the longest tiles
    order = np.argsort(-longest)[:12]
    for i in order:
        print("  long tile: %.1f us, max lane skips %d, warp finished %.1f" % (longest[i], sk[i], end[i]))
    ctx.close()


if __name__ == "__main__":
    main()
