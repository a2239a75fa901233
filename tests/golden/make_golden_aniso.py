"""Generates the anisotropic-path golden fixture FROM THE REFERENCE ITSELF (oracle/_ref/libfluidref.so).

Run in the build container (needs /root/reference at build time):
    python tests/golden/make_golden_aniso.py
Output (committed): aniso8k_160x90.npz
  eig_C / eig_vals / eig_vecs      Eigen::SelfAdjointEigenSolver<Matrix3f>::computeDirect known answers
                                   (vendor/eigen/.../SelfAdjointEigenSolver.h:583-741) on covariance-like matrices
  wpca_p / wpca_len / wpca_xyz / wpca_G   RayMarcher::WPCA (RayMarcher.cpp:114-254): query point, its 2h neighbour
                                   positions in the reference's list order, and the resulting G
  aw_G / aw_det / aw_r / aw_W / aw_gradW  AnisotropicKernel::W / gradW (Kernel.cpp:63-107), glm::determinant
  cubic_r / cubic_W                CubicKernel::W (Kernel.cpp:117-125) at h_ext = 0.2
  depth / positions / normals      PerPixel_Anisotropic (RayMarcher.cpp:346-423) over the dambreak8k frame,
                                   camera_close_16x9, default settings (depth input = the oracle's depth pre-pass)
The reference build links this image's glibc (std::atan2/sin/cos inside computeDirect); the values are stored bit-exactly.
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import oracle_lib  # noqa: E402

scenes = importlib.import_module("bachelor-thesis_b200.scenes")


def camera(name):
    with open(os.path.join(HERE, name + ".json")) as f:
        d = json.load(f)
    return {k: np.array(v, dtype=np.float32) for k, v in d["float32"].items()}


def main():
    ref = oracle_lib.Ref()
    orc = oracle_lib.Oracle()
    assert ref.has_aniso
    rng = np.random.default_rng(11)

    # ---- computeDirect ------------------------------------------------------------------------------
    mats = []
    for i in range(256):
        a = rng.normal(0, 1, size=(3, 3)) * rng.uniform(1e-3, 0.2)
        c = (a @ a.T)
        if i % 8 == 1:
            c = np.diag(np.diag(c))                      # diagonal
        if i % 8 == 2:
            v = rng.normal(size=3); c = np.outer(v, v) * 1e-3   # rank 1
        if i % 8 == 3:
            c = np.eye(3) * rng.uniform(1e-4, 1e-2)      # isotropic: evals[2] - evals[0] <= eps branch
        if i % 8 == 4:
            c[2, :] = 0; c[:, 2] = 0                     # rank 2 / planar
        mats.append(c.T.reshape(9))                      # column-major
    mats.append(np.zeros(9))
    eig_C = np.array(mats, np.float32)
    ev, vec = [], []
    for c in eig_C:
        a, b = ref.eigen3(c)
        ev.append(a); vec.append(b)

    # ---- frame, WPCA ------------------------------------------------------------------------------------
    xyz = scenes.dam_break(8000)
    ds = ref.dataset(xyz, 0.1, 2.0)
    s = oracle_lib.Settings(anisotropic=1)
    perm_ext = ds.particles_ext()
    pts = (xyz[rng.integers(0, len(xyz), 96)] + rng.normal(0, 0.05, size=(96, 3))).astype(np.float32)
    pts[0] = xyz.min(axis=0) - np.float32(0.15)          # one or zero neighbours: the N <= N_eps branch
    pts[1] = xyz.max(axis=0) + np.float32(0.12)
    nb = [perm_ext[ds.neighbors(p, ext=1)] for p in pts]
    nb_len = np.array([len(x) for x in nb], np.int32)
    G = np.array([ds.wpca(s, pts[i], nb[i]) if nb_len[i] else np.zeros(9, np.float32) for i in range(len(pts))], np.float32)
    nb_flat = np.concatenate([x for x in nb if len(x)], axis=0).astype(np.float32)

    # ---- AnisotropicKernel / CubicKernel ---------------------------------------------------------------
    keep = [i for i in range(len(pts)) if nb_len[i] > 1][:48]
    aw_G = G[keep]
    aw_det = np.array([ref.det3(g) for g in aw_G], np.float32)
    aw_r = rng.uniform(-0.1, 0.1, size=(len(keep), 8, 3)).astype(np.float32)
    aw_r[:, 0] = 0.0
    with np.errstate(all="ignore"):
        aw_W = np.array([[ref.aniso_W(0.1, aw_G[i], aw_det[i], r) for r in aw_r[i]] for i in range(len(keep))], np.float32)
        aw_g = np.array([[ref.aniso_gradW(0.1, aw_G[i], aw_det[i], r) for r in aw_r[i]] for i in range(len(keep))], np.float32)
    cubic_r = rng.uniform(-0.15, 0.15, size=(256, 3)).astype(np.float32)
    cubic_r[0] = 0.0
    cubic_r[1] = (0.2, 0.0, 0.0)
    cubic_W = np.array([ref.cubic_W(0.2, r) for r in cubic_r], np.float32)

    # ---- the march ----------------------------------------------------------------------------------------
    Wd, Hd = 160, 90
    cam = camera("camera_close_16x9")
    of = orc.frame(xyz, 0.1, 2.0)
    depth = of.depth_prepass(Wd, Hd, cam["view"], cam["proj"])
    out = ds.march(Wd, Hd, s, cam["inv_proj_view"], cam["position"], depth, threads=1)
    pos, nrm = out[0], out[1]

    np.savez_compressed(os.path.join(HERE, "aniso8k_160x90.npz"), h=np.float32(0.1), mult=np.float32(2.0),
                        eig_C=eig_C, eig_vals=np.array(ev, np.float32), eig_vecs=np.array(vec, np.float32),
                        wpca_p=pts, wpca_len=nb_len, wpca_xyz=nb_flat, wpca_G=G,
                        aw_G=aw_G, aw_det=aw_det, aw_r=aw_r, aw_W=aw_W, aw_gradW=aw_g,
                        cubic_r=cubic_r, cubic_W=cubic_W,
                        W=Wd, H=Hd, depth=depth, positions=pos, normals=nrm)
    print("hits", int((pos[..., 3] == 1).sum()), "of", int((depth != 1).sum()), "covered")


if __name__ == "__main__":
    main()
