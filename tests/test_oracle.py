"""Pins the CPU oracle (oracle/fluid_oracle.c): against the golden fixtures generated FROM THE REFERENCE
(tests/golden/make_golden.py) and, where the reference build is present, against the reference's own
code (oracle/_ref/libfluidref.so) bit for bit.  No GPU needed."""
import importlib
import math
import os

import numpy as np
import pytest

import oracle_lib
from conftest import GOLDEN, golden_camera

scenes = importlib.import_module("bachelor-thesis_b200.scenes")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def small():
    return np.load(os.path.join(GOLDEN, "dambreak8k_160x90.npz"))


@pytest.fixture(scope="module")
def kv():
    return np.load(os.path.join(GOLDEN, "kernel_vectors.npz"))


# ---- Kernel.cpp / intersectAABB known answers (golden, from the reference) ----------------------------
def test_kernel_W_golden(oracle, kv):
    h = float(kv["h"])
    assert np.float32(oracle.W0(h)) == kv["W0"]
    got = np.array([oracle.W(h, r) for r in kv["r"]], np.float32)
    assert np.array_equal(bits(got), bits(kv["W"]))
    # documented special values: W(0) = W0, cut-off at q = 1, branch at q = 0.5
    assert got[0] == kv["W0"] and got[2] == 0.0


def test_kernel_gradW_golden(oracle, kv):
    h = float(kv["h"])
    got = np.array([oracle.gradW(h, r) for r in kv["r"]], np.float32)
    assert np.array_equal(bits(got), bits(kv["gradW"]))          # NaN at r = 0 included, bit for bit


def test_intersect_aabb_golden(oracle, kv):
    got = np.array([oracle.intersect_aabb(kv["aabb_o"][i], kv["aabb_d"][i], kv["aabb_min"][i], kv["aabb_max"][i])
                    for i in range(len(kv["aabb_o"]))], np.float32)
    assert np.array_equal(bits(got), bits(kv["aabb_out"]))


def test_cos_half_pi_accuracy(oracle):
    s = np.linspace(0, 1, 2001)
    got = np.array([oracle.cos_half_pi(x) for x in s], np.float64)
    assert np.abs(got - np.cos(np.pi / 2 * s)).max() < 2e-7


# ---- a whole small frame (golden, from the reference) ---------------------------------------------------
def test_frame_geometry_golden(oracle, small):
    f = oracle.frame(small["xyz"], float(small["h"]), float(small["mult"]), count_mode=1)
    assert np.array_equal(bits(f.min), bits(small["frame_min"]))
    assert np.array_equal(bits(f.max), bits(small["frame_max"]))
    assert np.array_equal(f.dims, small["grid_dims"])
    assert np.array_equal(bits(f.particles()), bits(small["particles_sorted"]))   # Morton permutation
    counts, flags = f.grid()
    assert np.array_equal(counts, small["grid_counts"])
    assert np.array_equal(flags, small["grid_flags"])


def test_count_modes_differ_only_at_cell_faces(oracle, small):
    """cell-exact (FR_COUNT_CELL_EXACT) vs centre-box (the _ref stand-in, the CUDA default): same total, few cells differ"""
    a = oracle.frame(small["xyz"], 0.1, 2.0, count_mode=0).grid()[0]
    b = oracle.frame(small["xyz"], 0.1, 2.0, count_mode=1).grid()[0]
    assert int(a.sum()) == len(small["xyz"])
    assert (a != b).sum() <= 8


def test_neighbour_lists_golden(oracle, small):
    f = oracle.frame(small["xyz"], 0.1, 2.0)
    perm = f.particles()
    off = 0
    for p, n in zip(small["query_points"], small["neighbour_len"]):
        ids = f.neighbors(p)
        assert len(ids) == n
        assert np.array_equal(bits(perm[ids]), bits(small["neighbour_xyz"][off:off + n]))   # same order too
        off += n


def test_neighbour_sets_match_brute_force(oracle, small):
    f = oracle.frame(small["xyz"], 0.1, 2.0)
    perm = f.particles()
    h2 = np.float32(0.1) * np.float32(0.1)
    for p in small["query_points"][:32]:
        d = (p[None, :] - perm).astype(np.float32)
        l2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + d[:, 2] * d[:, 2]
        want = set(np.nonzero(l2 < h2)[0].tolist())
        assert set(f.neighbors(p).tolist()) == want


def test_march_golden(oracle, small):
    cam = golden_camera("camera_close_16x9")
    W, H = int(small["W"]), int(small["H"])
    f = oracle.frame(small["xyz"], 0.1, 2.0, count_mode=1)
    depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
    assert np.array_equal(bits(depth), bits(small["depth"]))     # regression pin of the depth restatement
    pos, nrm, band, steps, cnt = f.march(W, H, oracle_lib.Settings(), cam["inv_proj_view"], cam["position"], depth)
    assert np.array_equal(bits(pos), bits(small["positions"]))
    assert np.array_equal(bits(nrm), bits(small["normals"]))
    assert cnt["hit_rays"] == int(small["positions"][..., 3].sum()) > 500


def test_depth_format_matches_reference_dump():
    """tools/DEPTH of the reference (1600x900 float32): depth in (0.9, 1), cleared to exactly 1.0.
    The restated pre-pass must produce the same format/range."""
    small = np.load(os.path.join(GOLDEN, "dambreak8k_160x90.npz"))
    d = small["depth"]
    assert d.dtype == np.float32 and d.max() == 1.0 and 0.5 < d.min() < 1.0
    assert 0.05 < (d < 1).mean() < 0.5


# ---- against the reference's own code, larger cases -------------------------------------------------------
@pytest.mark.parametrize("n,W,H,cam_name", [(20000, 320, 180, "camera_default_16x9"),
                                            (20000, 256, 144, "camera_orbit_a_16x9"),
                                            (64000, 320, 180, "camera_orbit_b_16x9")])
def test_oracle_equals_reference_build(oracle, ref, n, W, H, cam_name):
    xyz = scenes.dam_break(n)
    cam = golden_camera(cam_name)
    f = oracle.frame(xyz, 0.1, 2.0, count_mode=1)
    ds = ref.dataset(xyz, 0.1, 2.0)
    try:
        assert np.array_equal(bits(f.min), bits(ds.min)) and np.array_equal(f.dims, ds.dims)
        assert np.array_equal(bits(f.particles()), bits(ds.particles()))
        oc, of_ = f.grid()
        rc, rf = ds.grid()
        assert np.array_equal(oc, rc) and np.array_equal(of_, rf)
        depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
        s = oracle_lib.Settings()
        pos, nrm, *_ = f.march(W, H, s, cam["inv_proj_view"], cam["position"], depth)
        rpos, rnrm, _ = ds.march(W, H, s, cam["inv_proj_view"], cam["position"], depth)
        assert np.array_equal(bits(pos), bits(rpos))
        assert np.array_equal(bits(nrm), bits(rnrm))
        rng = np.random.default_rng(3)
        for p in xyz[rng.integers(0, len(xyz), 50)] + rng.normal(0, 0.05, (50, 3)).astype(np.float32):
            assert np.array_equal(f.neighbors(p), ds.neighbors(p))
    finally:
        ds.close()


def test_reference_pool_skips_last_pixel(ref, oracle):
    """ThreadPool.cpp:50: the reference's own pool never runs index W*H-1"""
    xyz = scenes.random_block(3000, 0.4)
    cam = golden_camera("camera_close_16x9")
    W, H = 64, 36
    f = oracle.frame(xyz, 0.1, 2.0, count_mode=1)
    depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
    depth[-1, -1] = depth[H // 2, W // 2]          # make the last pixel covered
    ds = ref.dataset(xyz, 0.1, 2.0)
    s = oracle_lib.Settings()
    pos = np.full((H, W, 4), 7.0, np.float32)
    full, _, _ = ds.march(W, H, s, cam["inv_proj_view"], cam["position"], depth)
    L = ds.L
    import ctypes as C
    f32p = C.POINTER(C.c_float)
    nrm = np.full((H, W, 4), 7.0, np.float32)
    ipv, cp, dd = (np.ascontiguousarray(x, np.float32) for x in (cam["inv_proj_view"], cam["position"], depth))
    L.ref_march(ds.ptr, 0, W, H, s.max_steps, s.step_size, s.iso_density, 0, s.k_n, s.k_r, s.k_s, s.n_eps,
                ipv.ctypes.data_as(f32p), cp.ctypes.data_as(f32p), dd.ctypes.data_as(f32p),
                pos.ctypes.data_as(f32p), nrm.ctypes.data_as(f32p), 0, 1)
    assert np.all(pos[-1, -1] == 7.0)                           # untouched by the reference pool
    assert np.array_equal(bits(pos.reshape(-1, 4)[:-1]), bits(full.reshape(-1, 4)[:-1]))
