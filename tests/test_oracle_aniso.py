"""Pins the anisotropic path of the CPU oracle (oracle/fluid_oracle_aniso.inc): against the golden fixture generated
FROM THE REFERENCE (tests/golden/make_golden_aniso.py -> aniso8k_160x90.npz) and, where the reference build is present,
against live calls into the reference's own code (oracle/_ref/libfluidref.so), bit for bit.  No GPU needed."""
import importlib
import os

import numpy as np
import pytest

import oracle_lib
from conftest import GOLDEN, golden_camera

scenes = importlib.import_module("bachelor-thesis_b200.scenes")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "aniso8k_160x90.npz"))


ANISO = oracle_lib.Settings(anisotropic=1)


# ---- Eigen computeDirect (SelfAdjointEigenSolver.h:583-741) ---------------------------------------------
def test_eigen3_golden(oracle, g):
    for c, ev, vec in zip(g["eig_C"], g["eig_vals"], g["eig_vecs"]):
        got_ev, got_vec = oracle.eigen3(c)
        assert np.array_equal(bits(got_ev), bits(ev))
        assert np.array_equal(bits(got_vec), bits(vec))


def test_eigen3_is_an_eigendecomposition(oracle):
    rng = np.random.default_rng(5)
    for _ in range(64):
        a = rng.normal(size=(3, 3)) * 0.05
        c = (a @ a.T).astype(np.float32)
        ev, vec = oracle.eigen3(c.T.reshape(9))
        V = vec.reshape(3, 3).T.astype(np.float64)          # columns = eigenvectors
        assert ev[0] <= ev[1] <= ev[2]
        assert np.abs(V.T @ V - np.eye(3)).max() < 1e-4
        assert np.abs(c.astype(np.float64) @ V - V * ev).max() < 1e-5 * max(1.0, float(ev[2]) / 1e-3)


# ---- glibc restatements used inside computeDirect -----------------------------------------------------------
def test_trig_restatements_equal_libm_on_the_solver_ranges(oracle):
    # theta = atan2(sqrt(q), half_b) / 3 lies in [0, pi/3]: every float in [2^-14, 1.0472], sin and cos
    assert oracle.lib.fo_trig_selftest(0, 6.1e-5, 1.0472, 0, 0) == 0
    assert oracle.lib.fo_trig_selftest(1, 6.1e-5, 1.0472, 0, 0) == 0
    assert oracle.lib.fo_trig_selftest(2, 0.0, 0.0, 4_000_000, 123) == 0
    assert oracle.lib.fo_sinf(0.0) == 0.0 and oracle.lib.fo_cosf(0.0) == 1.0


# ---- CubicKernel / AnisotropicKernel / determinant -------------------------------------------------------------
def test_cubic_kernel_golden(oracle, g):
    got = np.array([oracle.cubic_W(0.2, r) for r in g["cubic_r"]], np.float32)
    assert np.array_equal(bits(got), bits(g["cubic_W"]))
    assert got[0] == 1.0 and got[1] == 0.0              # W(0) = 1, cut-off at d >= h


def test_aniso_kernel_golden(oracle, g):
    with np.errstate(all="ignore"):
        for G, det, rs, Ws, gs in zip(g["aw_G"], g["aw_det"], g["aw_r"], g["aw_W"], g["aw_gradW"]):
            assert bits(oracle.det3(G)) == bits(det)
            for r, w, gw in zip(rs, Ws, gs):
                assert bits(oracle.aniso_W(0.1, G, det, r)) == bits(w)
                assert np.array_equal(bits(oracle.aniso_gradW(0.1, G, det, r)), bits(gw))   # NaN at r = 0, bit for bit


# ---- WPCA (RayMarcher.cpp:114-254) ---------------------------------------------------------------------------------
def test_wpca_golden(oracle, g):
    off = 0
    n_small = 0
    for p, n, G in zip(g["wpca_p"], g["wpca_len"], g["wpca_G"]):
        nb = g["wpca_xyz"][off:off + n]
        off += n
        if n == 0:
            continue
        n_small += n <= ANISO.n_eps
        got = oracle.wpca(0.1, 0.2, ANISO, p, nb)
        assert np.array_equal(bits(got), bits(G))
    assert off == len(g["wpca_xyz"])


def test_ext_neighbour_lists_golden(oracle, g):
    """Dataset::GetNeighborsExt (Dataset.cpp:282-290): the oracle's 2h search returns the reference's positions in the
    reference's order, and exactly the brute-force set."""
    xyz = scenes.dam_break(8000)
    f = oracle.frame(xyz, 0.1, 2.0)
    perm = f.particles_ext()
    off = 0
    r2 = np.float32(0.2) * np.float32(0.2)
    for p, n in zip(g["wpca_p"], g["wpca_len"]):
        ids = f.neighbors(p, ext=1)
        assert len(ids) == n
        assert np.array_equal(bits(perm[ids]), bits(g["wpca_xyz"][off:off + n]))
        off += n
        d = p[None, :] - perm
        l2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        assert set(ids.tolist()) == set(np.nonzero(l2 < r2)[0].tolist())


# ---- PerPixel_Anisotropic (RayMarcher.cpp:346-423) --------------------------------------------------------------------
def test_march_anisotropic_golden(oracle, g):
    xyz = scenes.dam_break(8000)
    cam = golden_camera("camera_close_16x9")
    W, H = int(g["W"]), int(g["H"])
    f = oracle.frame(xyz, 0.1, 2.0)
    depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
    assert np.array_equal(bits(depth), bits(g["depth"]))
    pos, nrm, band, steps, cnt = f.march(W, H, ANISO, cam["inv_proj_view"], cam["position"], depth)
    assert np.array_equal(bits(pos), bits(g["positions"]))
    assert np.array_equal(bits(nrm), bits(g["normals"]))
    assert cnt["hit_rays"] == int((g["positions"][..., 3] == 1).sum()) > 1500
    # the anisotropic surface differs from the isotropic one (the fixture is not the isotropic image by accident)
    pos_iso, *_ = f.march(W, H, oracle_lib.Settings(), cam["inv_proj_view"], cam["position"], depth)
    assert not np.array_equal(bits(pos), bits(pos_iso))


def test_march_anisotropic_thread_count_independent(oracle, g):
    xyz = scenes.dam_break(8000)
    cam = golden_camera("camera_close_16x9")
    f = oracle.frame(xyz, 0.1, 2.0)
    a = f.march(80, 45, ANISO, cam["inv_proj_view"], cam["position"], g["depth"][::2, ::2].copy(), threads=1)
    b = f.march(80, 45, ANISO, cam["inv_proj_view"], cam["position"], g["depth"][::2, ::2].copy(), threads=4)
    assert np.array_equal(bits(a[0]), bits(b[0])) and np.array_equal(bits(a[1]), bits(b[1]))


# ---- live against the reference build (only where oracle/_ref exists) ---------------------------------------------------
def test_march_anisotropic_vs_reference_build(oracle, ref):
    if not ref.has_aniso:
        pytest.skip("oracle/_ref predates the anisotropic harness")
    xyz = scenes.dam_break(6000, seed=3)
    cam = golden_camera("camera_orbit_a_16x9")
    W, H = 128, 72
    f = oracle.frame(xyz, 0.1, 2.0)
    depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
    for s in (ANISO, oracle_lib.Settings(anisotropic=1, k_n=0.4, k_r=4.0, k_s=1400.0, n_eps=25, step_size=0.02, max_steps=40)):
        pos, nrm, *_ = f.march(W, H, s, cam["inv_proj_view"], cam["position"], depth)
        ds = ref.dataset(xyz, 0.1, 2.0)
        out = ds.march(W, H, s, cam["inv_proj_view"], cam["position"], depth, threads=0)
        ds.close()
        assert np.array_equal(bits(pos), bits(out[0]))
        assert np.array_equal(bits(nrm), bits(out[1]))


def test_eigen3_and_wpca_vs_reference_build(oracle, ref):
    if not ref.has_aniso:
        pytest.skip("oracle/_ref predates the anisotropic harness")
    rng = np.random.default_rng(99)
    for _ in range(300):
        a = rng.normal(size=(3, 3)) * rng.uniform(1e-4, 0.3)
        c = (a @ a.T).astype(np.float32).T.reshape(9)
        e0, v0 = oracle.eigen3(c)
        e1, v1 = ref.eigen3(c)
        assert np.array_equal(bits(e0), bits(e1)) and np.array_equal(bits(v0), bits(v1))
