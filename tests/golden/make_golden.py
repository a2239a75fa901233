"""Generates the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference, i.e. oracle/_ref/libfluidref.so):
    python tests/golden/make_golden.py
Outputs (committed):
  camera_*.json          matrices of the reference's Camera3D / CameraController3D for a few orbits
  kernel_vectors.npz     CubicSplineKernel::W / gradW and intersectAABB known answers
  dambreak8k_160x90.npz  a small frame: input particles, Frame geometry, occupancy flags, neighbour
                         lists at sample points, and the reference marcher's positions/normals for the
                         camera_close_16x9 camera and default settings (depth input = the oracle's depth pre-pass)
Everything the reference produced is stored bit-exactly (float32 arrays).
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import importlib  # noqa: E402

import oracle_lib  # noqa: E402

scenes = importlib.import_module("bachelor-thesis_b200.scenes")


def main():
    ref = oracle_lib.Ref()
    orc = oracle_lib.Oracle()

    # ---- cameras ---------------------------------------------------------------------------------
    cams = {
        "camera_default_16x9": dict(fov=math.radians(60.0), aspect=16.0 / 9.0, near=0.1, far=1000.0, R=10.0, rot_x=0.0, rot_y=0.0),
        "camera_orbit_a_16x9": dict(fov=math.radians(60.0), aspect=16.0 / 9.0, near=0.1, far=1000.0, R=9.0, rot_x=0.6, rot_y=0.35),
        "camera_orbit_b_16x9": dict(fov=math.radians(60.0), aspect=16.0 / 9.0, near=0.1, far=1000.0, R=12.0, rot_x=-2.4, rot_y=-0.5),
        "camera_close_16x9": dict(fov=math.radians(60.0), aspect=16.0 / 9.0, near=0.1, far=1000.0, R=2.5, rot_x=0.3, rot_y=0.2),
    }
    for name, kw in cams.items():
        c = ref.camera(**kw)
        out = dict(params=kw, float32={k: [float(x) for x in v] for k, v in c.items()},
                   hex={k: [np.float32(x).tobytes().hex() for x in v] for k, v in c.items()},
                   source="reference Camera3D.cpp / CameraController3D.cpp via oracle/_ref/libfluidref.so")
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f, indent=1)

    # ---- kernel known answers -----------------------------------------------------------------------
    rng = np.random.default_rng(7)
    h = np.float32(0.1)
    r = (rng.uniform(-0.12, 0.12, size=(512, 3))).astype(np.float32)
    r[0] = (0.0, 0.0, 0.0)
    r[1] = (0.05, 0.0, 0.0)          # q = 0.5 branch point
    r[2] = (0.1, 0.0, 0.0)           # q = 1 cut-off
    r[3] = (0.0999999, 0.0, 0.0)
    W = np.array([ref.W(float(h), x) for x in r], np.float32)
    with np.errstate(all="ignore"):
        G = np.array([ref.gradW(float(h), x) for x in r], np.float32)
    o = rng.uniform(-1, 1, size=(256, 3)).astype(np.float32)
    d = rng.normal(0, 0.01, size=(256, 3)).astype(np.float32)
    d[:8, 0] = 0.0                   # axis-parallel rays: +-inf slabs
    d[8:12, 1] = 0.0
    bmin = (np.floor(o * 10) / 10).astype(np.float32)
    bmax = (bmin + np.float32(0.1)).astype(np.float32)
    with np.errstate(all="ignore"):
        X = np.array([ref.intersect_aabb(o[i], d[i], bmin[i], bmax[i]) for i in range(len(o))], np.float32)
    np.savez_compressed(os.path.join(HERE, "kernel_vectors.npz"), h=h, r=r, W=W, gradW=G, W0=np.float32(ref.W0(float(h))),
                        aabb_o=o, aabb_d=d, aabb_min=bmin, aabb_max=bmax, aabb_out=X)

    # ---- a small frame -------------------------------------------------------------------------------
    Wd, Hd = 160, 90
    xyz = scenes.dam_break(8000)
    ds = ref.dataset(xyz, 0.1, 2.0)
    counts, flags = ds.grid()
    cam = ref.camera(**cams["camera_close_16x9"])
    of = orc.frame(xyz, 0.1, 2.0)
    depth = of.depth_prepass(Wd, Hd, cam["view"], cam["proj"])
    s = oracle_lib.Settings()
    pos, nrm, _ = ds.march(Wd, Hd, s, cam["inv_proj_view"], cam["position"], depth, threads=1)
    pts = (xyz[rng.integers(0, len(xyz), 64)] + rng.normal(0, 0.04, size=(64, 3))).astype(np.float32)
    perm = ds.particles()
    nb = [perm[ds.neighbors(p)] for p in pts]            # neighbour POSITIONS in result order
    nb_len = np.array([len(x) for x in nb], np.int32)
    nb_flat = np.concatenate(nb, axis=0).astype(np.float32) if nb_len.sum() else np.zeros((0, 3), np.float32)
    np.savez_compressed(os.path.join(HERE, "dambreak8k_160x90.npz"), xyz=xyz, h=np.float32(0.1), mult=np.float32(2.0),
                        W=Wd, H=Hd, frame_min=ds.min, frame_max=ds.max, grid_dims=ds.dims, grid_counts=counts,
                        grid_flags=flags, particles_sorted=perm, depth=depth, positions=pos, normals=nrm,
                        query_points=pts, neighbour_len=nb_len, neighbour_xyz=nb_flat,
                        box_mode=np.int32(0))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
