"""GPU parity tests proper: the CUDA path, called through the C ABI (libfluidmarch.so), against the CPU
oracle on the same inputs and against the golden fixtures generated from the reference.

Parity bar (bit-exact wherever the work is integer/index work or a fixed FP32 sequence):
  * frame geometry, occupancy counts/flags, per-cell particle sets, neighbour lists: bit-exact
  * depth pre-pass: bit-exact (fixed FP32 op sequence, order-independent min)
  * hit mask, hit positions, normals: bit-exact -- the kernel accumulates in the reference's neighbour order,
    so no epsilon band around the iso threshold is needed (the band allowed by the spec is empty)
  * RGBA8: +-1 code value on a small fraction of pixels (powf of the sRGB curve is the only non-IEEE op)
"""
import importlib
import os

import numpy as np
import pytest

import oracle_lib
from conftest import GOLDEN, golden_camera

pytestmark = pytest.mark.gpu
scenes = importlib.import_module("bachelor-thesis_b200.scenes")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def set_cam(ctx, cam):
    ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])


def search_key(xyz, h, info):
    inv = np.float32(1.0) / np.float32(h)
    t = (inv * xyz.astype(np.float32)).astype(np.int32)          # truncation toward zero, like (int)
    k = np.where(xyz >= 0, t, t - 1) - info["search_min"][None, :]
    kd = info["search_dims"].astype(np.int64)
    return (k[:, 0].astype(np.int64) * kd[1] + k[:, 1]) * kd[2] + k[:, 2]


@pytest.fixture(scope="module")
def small():
    return np.load(os.path.join(GOLDEN, "dambreak8k_160x90.npz"))


# ---- (a1-a4) grid build ---------------------------------------------------------------------------------
@pytest.mark.parametrize("n,gen", [(8000, "dam"), (64000, "dam"), (5000, "rand"), (1, "rand"), (37, "rand")])
def test_grid_build_parity(fm, oracle, gpu_ctx_factory, n, gen):
    xyz = scenes.dam_break(n) if gen == "dam" else scenes.random_block(n, 0.7)
    ctx = gpu_ctx_factory(64, 64)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    g = ctx.download_frame(0)
    info = g["info"]
    f = oracle.frame(xyz, 0.1, 2.0)
    assert info["num_particles"] == len(xyz)
    assert np.array_equal(bits(info["min"]), bits(f.min)) and np.array_equal(bits(info["max"]), bits(f.max))
    assert np.array_equal(info["grid_dims"], f.dims)
    counts, flags = f.grid()
    assert np.array_equal(g["grid_counts"], counts)
    assert np.array_equal(g["grid_flags"], flags)
    assert info["occupied_cells"] == int(flags.sum())
    # counting sort: a permutation of the input, grouped by cell key, ascending original index inside a cell
    idx = g["sorted_index"].astype(np.int64)
    assert np.array_equal(np.sort(idx), np.arange(len(xyz)))
    assert np.array_equal(bits(g["sorted_xyz"]), bits(xyz[idx]))
    key = search_key(xyz, 0.1, info)
    ks = key[idx]
    assert np.all(np.diff(ks) >= 0)
    same = np.diff(ks) == 0
    assert np.all(np.diff(idx)[same] > 0)
    cells = int(np.prod(info["search_dims"].astype(np.int64)))
    want_start = np.concatenate([[0], np.cumsum(np.bincount(key, minlength=cells))])
    assert np.array_equal(g["cell_start"].astype(np.int64), want_start)


def test_grid_matches_golden_reference_frame(fm, gpu_ctx_factory, small):
    """against the Frame the reference itself built (tests/golden, counts in the reference's box reading)"""
    ctx = gpu_ctx_factory(64, 64)
    ctx.upload_frame(0, small["xyz"], float(small["h"]), float(small["mult"]))
    g = ctx.download_frame(0)
    assert np.array_equal(bits(g["info"]["min"]), bits(small["frame_min"]))
    assert np.array_equal(bits(g["info"]["max"]), bits(small["frame_max"]))
    assert np.array_equal(g["info"]["grid_dims"], small["grid_dims"])
    # FR_COUNT_CENTRE_BOX (the default) is the reading of find_neighbors_box the reference build used: identical
    assert np.array_equal(g["grid_counts"], small["grid_counts"])
    assert np.array_equal(g["grid_flags"], small["grid_flags"])
    # cell-exact reading: only particles within an ulp of a face may move
    ctx.set_count_mode(fm.FR_COUNT_CELL_EXACT)
    ctx.upload_frame(0, small["xyz"], float(small["h"]), float(small["mult"]))
    g = ctx.download_frame(0)
    ctx.set_count_mode(fm.FR_COUNT_CENTRE_BOX)
    assert (g["grid_counts"] != small["grid_counts"]).sum() <= 8
    assert (g["grid_flags"] != small["grid_flags"]).sum() <= 4
    assert int(g["grid_counts"].sum()) == len(small["xyz"])


@pytest.mark.parametrize("mode", [0, 1])
def test_count_modes_match_oracle(fm, oracle, gpu_ctx_factory, mode):
    """both readings of find_neighbors_box, on particles that sit on and next to cell faces"""
    rng = np.random.default_rng(5)
    h = np.float32(0.1)
    lattice = (np.stack(np.meshgrid(*[np.arange(-4, 5)] * 3, indexing="ij"), -1).reshape(-1, 3) * h).astype(np.float32)
    near = np.nextafter(lattice[:300], np.float32(10.0)).astype(np.float32)
    xyz = np.concatenate([scenes.dam_break(20000), lattice, near, rng.uniform(-0.4, 0.4, (2000, 3)).astype(np.float32)])
    ctx = gpu_ctx_factory(64, 64)
    ctx.set_count_mode(mode)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    g = ctx.download_frame(0)
    ctx.set_count_mode(fm.FR_COUNT_CENTRE_BOX)
    counts, flags = oracle.frame(xyz, 0.1, 2.0, count_mode=mode).grid()
    assert np.array_equal(g["grid_counts"], counts)
    assert np.array_equal(g["grid_flags"], flags)
    assert g["info"]["occupied_cells"] == int(flags.sum())


# ---- (a5) neighbour sets ---------------------------------------------------------------------------------
def test_neighbour_lists_bit_exact(fm, oracle, gpu_ctx_factory, small):
    xyz = small["xyz"]
    ctx = gpu_ctx_factory(64, 64)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    pts = small["query_points"]
    counts, ids = ctx.query_neighbors(0, pts, cap=256)
    off = 0
    for i, n in enumerate(small["neighbour_len"]):
        assert counts[i] == n
        got = xyz[ids[i, :n]]
        assert np.array_equal(bits(got), bits(small["neighbour_xyz"][off:off + n]))   # reference order, bit for bit
        off += n
    # and against brute force as sets, on fresh points (incl. points outside the particle AABB)
    rng = np.random.default_rng(11)
    pts = np.concatenate([xyz[rng.integers(0, len(xyz), 300)] + rng.normal(0, 0.06, (300, 3)),
                          rng.uniform(-3, 3, (50, 3))]).astype(np.float32)
    counts, ids = ctx.query_neighbors(0, pts, cap=256)
    h2 = np.float32(0.1) * np.float32(0.1)
    for i, p in enumerate(pts):
        d = (p[None, :] - xyz).astype(np.float32)
        l2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + d[:, 2] * d[:, 2]
        want = set(np.nonzero(l2 < h2)[0].tolist())
        assert counts[i] == len(want)
        assert set(ids[i, :counts[i]].tolist()) == want


def test_density_query_matches_oracle_kernels(fm, oracle, gpu_ctx_factory, small):
    xyz = small["xyz"]
    ctx = gpu_ctx_factory(64, 64)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    f = oracle.frame(xyz, 0.1, 2.0)
    perm = f.particles()
    pts = small["query_points"][:24]
    rho, grad = ctx.query_density(0, pts)
    for i, p in enumerate(pts):
        ids = f.neighbors(p)
        want = np.float32(0)
        g = np.zeros(3, np.float32)
        for j in ids:
            r = (perm[j] - p).astype(np.float32)
            want = np.float32(want + oracle.W(0.1, r))
            g = (g + oracle.gradW(0.1, r)).astype(np.float32)
        assert bits(rho[i:i + 1])[0] == bits(np.array([want]))[0]
        assert np.array_equal(bits(grad[i]), bits(g))


def test_collisions_and_cell_boundaries(fm, oracle, gpu_ctx_factory):
    """the domain's degenerate inputs: duplicated particles, query points that coincide with particles (r = 0: W = W0,
    gradW = NaN through normalize(0), Kernel.cpp:43), particles and queries exactly on cell faces and at negative
    coordinates (CompactNSearch's floor-like cell index, Dataset.cpp:28's multiply-by-reciprocal)"""
    h = np.float32(0.1)
    rng = np.random.default_rng(17)
    base = rng.uniform(-0.35, 0.35, size=(3000, 3)).astype(np.float32)
    lattice = (np.stack(np.meshgrid(*[np.arange(-3, 4)] * 3, indexing="ij"), -1).reshape(-1, 3) * h).astype(np.float32)   # on cell faces
    xyz = np.concatenate([base, base[:200], base[:50], lattice, lattice[:40]]).astype(np.float32)      # duplicates and triplicates
    ctx = gpu_ctx_factory(64, 64)
    ctx.upload_frame(0, xyz, float(h), 2.0)
    f = oracle.frame(xyz, float(h), 2.0)
    # grid structures first
    got, info = ctx.download_frame(0), ctx.frame_info(0)
    assert np.array_equal(info["grid_dims"], f.dims) and np.array_equal(bits(info["min"]), bits(f.min))
    assert np.array_equal(got["grid_counts"], f.grid()[0]) and np.array_equal(got["grid_flags"], f.grid()[1])
    # queries: on particles, on cell faces, just beside them
    eps = np.float32(1e-7)
    pts = np.concatenate([base[:40], lattice[100:140], lattice[100:140] + eps, lattice[100:140] - eps,
                          np.array([[0, 0, 0], [-0.0, 0.1, -0.1], [0.35, -0.35, 0.0]], np.float32)]).astype(np.float32)
    counts, ids = ctx.query_neighbors(0, pts, cap=512)
    perm = f.particles()
    rho, grad = ctx.query_density(0, pts)
    for i, p in enumerate(pts):
        want_ids = f.neighbors(p)
        assert counts[i] == len(want_ids)
        # same particles in the same order (the GPU reports original indices, the oracle its sorted ones)
        assert np.array_equal(bits(xyz[ids[i, :counts[i]]]), bits(perm[want_ids]))
        w = np.float32(0)
        g = np.zeros(3, np.float32)
        with np.errstate(all="ignore"):
            for j in want_ids:
                r = (perm[j] - p).astype(np.float32)
                w = np.float32(w + oracle.W(float(h), r))
                g = (g + oracle.gradW(float(h), r)).astype(np.float32)
        assert bits(rho[i:i + 1])[0] == bits(np.array([w]))[0]
        assert np.array_equal(np.isnan(grad[i]), np.isnan(g))
        assert np.array_equal(bits(grad[i])[~np.isnan(g)], bits(g)[~np.isnan(g)])
    assert np.isnan(grad[:40]).all()                  # a query on a particle: normalize(0) in gradW


# ---- (a11) depth pre-pass ----------------------------------------------------------------------------------
@pytest.mark.parametrize("n,W,H,cam_name", [(8000, 160, 90, "camera_close_16x9"), (64000, 1280, 720, "camera_default_16x9"),
                                            (20000, 333, 187, "camera_orbit_a_16x9"), (20000, 320, 180, "camera_orbit_b_16x9")])
def test_depth_prepass_bit_exact(fm, oracle, gpu_ctx_factory, n, W, H, cam_name):
    xyz = scenes.dam_break(n)
    cam = golden_camera(cam_name)
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    set_cam(ctx, cam)
    ctx.render(fm.FR_PASS_DEPTH)
    depth, *_ = ctx.download(True, False, False, False)
    want = oracle.frame(xyz, 0.1, 2.0).depth_prepass(W, H, cam["view"], cam["proj"])
    assert (want < 1).sum() > 50
    assert np.array_equal(bits(depth), bits(want))


# ---- (a10, a6-a8) the march ---------------------------------------------------------------------------------
def test_march_matches_golden_reference_output(fm, gpu_ctx_factory, small):
    """positions / normals the REFERENCE's PerPixel_Isotropic produced (tests/golden), bit for bit"""
    cam = golden_camera("camera_close_16x9")
    W, H = int(small["W"]), int(small["H"])
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, small["xyz"], 0.1, 2.0)
    set_cam(ctx, cam)
    ctx.set_settings(fm.VisualizationSettings())
    ctx.set_depth(small["depth"])
    ctx.render(fm.FR_PASS_MARCH)
    _, pos, nrm, _ = ctx.download(False, True, True, False)
    assert np.array_equal(bits(pos), bits(small["positions"]))
    assert np.array_equal(bits(nrm), bits(small["normals"]))
    c = ctx.counters()
    assert c["hit_rays"] == int(small["positions"][..., 3].sum())
    assert c["neighbour_overflow"] == 0


@pytest.mark.parametrize("n,W,H,cam_name,step,iso", [
    (64000, 1280, 720, "camera_default_16x9", 0.009, 1.0),        # BASELINE config C1
    (20000, 320, 180, "camera_orbit_a_16x9", 0.009, 1.0),
    (20000, 333, 187, "camera_orbit_b_16x9", 0.02, 1.0),
    (8000, 160, 90, "camera_close_16x9", 0.005, 400.0),           # threshold deep inside: many steps per ray
    (8000, 160, 90, "camera_close_16x9", 0.009, 5000.0),          # never reached: all covered rays miss
])
def test_march_bit_exact_vs_oracle(fm, oracle, gpu_ctx_factory, n, W, H, cam_name, step, iso):
    xyz = scenes.dam_break(n)
    cam = golden_camera(cam_name)
    f = oracle.frame(xyz, 0.1, 2.0)
    depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
    s = oracle_lib.Settings(step_size=step, iso_density=iso)
    pos, nrm, band, steps, cnt = f.march(W, H, s, cam["inv_proj_view"], cam["position"], depth)
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    set_cam(ctx, cam)
    ctx.set_settings(fm.VisualizationSettings(StepSize=step, IsoDensity=iso))
    ctx.set_depth(depth)
    ctx.render(fm.FR_PASS_MARCH)
    _, gpos, gnrm, _ = ctx.download(False, True, True, False)
    c = ctx.counters()
    assert c["covered_rays"] == cnt["covered_rays"] and c["hit_rays"] == cnt["hit_rays"]
    assert np.array_equal(gpos[..., 3], pos[..., 3])                     # hit / miss mask
    assert np.array_equal(bits(gpos), bits(pos))
    assert np.array_equal(bits(gnrm), bits(nrm))
    # the early exit only drops samples that lie outside the grid (density 0)
    assert c["ray_steps"] <= cnt["ray_steps"]
    assert c["ray_steps"] >= cnt["ray_steps"] - cnt["steps_outside_grid"]
    assert c["skip_iterations"] == cnt["skip_iterations"]


# BASELINE config C5: smoothing-radius sweep (neighbour count ~20 .. ~120 at fixed dx) and step-size sweep with
# MaxSteps scaled to keep MaxSteps * StepSize = 1.152, at a size the oracle finishes in seconds
@pytest.mark.parametrize("ratio,step_div", [(1.68, 11), (1.93, 20), (2.29, 5), (2.67, 11), (3.06, 11), (3.06, 5), (3.9, 11)])
def test_march_sweep_c5_bit_exact(fm, oracle, gpu_ctx_factory, ratio, step_div):
    dx = 0.05
    h = float(np.float32(ratio * dx))
    step = float(np.float32(h / step_div))
    max_steps = int(round(1.152 / step))
    W, H = 240, 135
    xyz = scenes.dam_break(24000, h=h, dx=dx)
    cam = golden_camera("camera_close_16x9")
    f = oracle.frame(xyz, h, 2.0)
    depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
    s = oracle_lib.Settings(step_size=step, max_steps=max_steps)
    pos, nrm, band, steps, cnt = f.march(W, H, s, cam["inv_proj_view"], cam["position"], depth)
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, h, 2.0)
    set_cam(ctx, cam)
    ctx.set_settings(fm.VisualizationSettings(StepSize=step, MaxSteps=max_steps))
    ctx.render(fm.FR_PASS_DEPTH | fm.FR_PASS_MARCH)
    gdepth, gpos, gnrm, _ = ctx.download(True, True, True, False)
    c = ctx.counters()
    assert np.array_equal(bits(gdepth), bits(depth))
    assert c["covered_rays"] == cnt["covered_rays"] and c["hit_rays"] == cnt["hit_rays"] and c["hit_rays"] > 500
    assert np.array_equal(bits(gpos), bits(pos))
    assert np.array_equal(bits(gnrm), bits(nrm))
    assert c["neighbour_overflow"] == 0
    # the sweep is about the neighbour count: ~ (4/3) pi ratio^3 in the interior, about half of it at the surface
    per_sample = c["neighbours"] / max(c["ray_steps"], 1)
    assert 0.2 * 4.19 * ratio ** 3 < per_sample < 1.1 * 4.19 * ratio ** 3
    if ratio >= 3.0:
        assert per_sample > 40          # more than one shared-memory list (32 entries) per sample: the resume path ran


def test_march_edge_cases(fm, oracle, gpu_ctx_factory):
    cam = golden_camera("camera_close_16x9")
    W, H = 61, 35                                                  # not a multiple of the 32x8 CTA footprint
    xyz = scenes.random_block(4000, 0.5)
    f = oracle.frame(xyz, 0.1, 2.0)
    depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    set_cam(ctx, cam)
    # empty depth image: nothing marched, outputs all zero
    ctx.set_depth(np.ones((H, W), np.float32))
    ctx.render(fm.FR_PASS_MARCH)
    _, p, n, _ = ctx.download(False, True, True, False)
    assert not p.any() and not n.any() and ctx.counters()["covered_rays"] == 0
    # ragged size, MaxSteps = 0 and 1
    for ms in (0, 1, 128):
        s = oracle_lib.Settings(max_steps=ms)
        pos, nrm, *_ = f.march(W, H, s, cam["inv_proj_view"], cam["position"], depth)
        ctx.set_settings(fm.VisualizationSettings(MaxSteps=ms))
        ctx.set_depth(depth)
        ctx.render(fm.FR_PASS_MARCH)
        _, p, n, _ = ctx.download(False, True, True, False)
        assert np.array_equal(bits(p), bits(pos)) and np.array_equal(bits(n), bits(nrm))
    # the reference pool never touches the last pixel (ThreadPool.cpp:50)
    d2 = depth.copy()
    d2[-1, -1] = d2[d2 < 1].min()
    ctx.set_settings(fm.VisualizationSettings(SkipLastPixel=True))
    ctx.set_depth(d2)
    ctx.render(fm.FR_PASS_MARCH)       # previous outputs stay in place for the skipped pixel
    _, p2, _, _ = ctx.download(False, True, True, False)
    assert np.array_equal(bits(p2[-1, -1]), bits(p[-1, -1]))


def test_fast_normals_within_tolerance(fm, oracle, gpu_ctx_factory):
    """fr_settings.fast_normals: hit mask and positions stay bit-exact; normals agree with the reference's
    per-component IEEE evaluation to 5e-6 per component (99.9 % of the hits to 2e-6): the stated FP32 tolerance"""
    xyz = scenes.dam_break(64000)
    cam = golden_camera("camera_default_16x9")
    W, H = 1280, 720
    f = oracle.frame(xyz, 0.1, 2.0)
    depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
    pos, nrm, *_ = f.march(W, H, oracle_lib.Settings(), cam["inv_proj_view"], cam["position"], depth)
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    set_cam(ctx, cam)
    ctx.set_settings(fm.VisualizationSettings(FastNormals=True))
    ctx.set_depth(depth)
    ctx.render(fm.FR_PASS_MARCH)
    _, gpos, gnrm, _ = ctx.download(False, True, True, False)
    assert np.array_equal(bits(gpos), bits(pos))
    hit = pos[..., 3] == 1
    assert np.array_equal(gnrm[..., 3], nrm[..., 3])
    err = np.abs(gnrm[hit] - nrm[hit]).max(axis=1)
    assert err.max() <= 5e-6 and np.percentile(err, 99.9) <= 2e-6 and np.median(err) <= 3e-7
    assert not np.array_equal(bits(gnrm), bits(nrm))          # it really is the other code path


def test_bisection_refines_toward_the_camera(fm, oracle, gpu_ctx_factory):
    """north_star item 3 (not in the reference): with bisection_steps > 0 the hit moves from the first sample at
    or above the threshold toward the last sample below it -- never past it, density still >= iso there"""
    xyz = scenes.dam_break(20000)
    cam = golden_camera("camera_orbit_a_16x9")
    W, H = 320, 180
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    set_cam(ctx, cam)
    ctx.set_settings(fm.VisualizationSettings())
    ctx.render(fm.FR_PASS_DEPTH | fm.FR_PASS_MARCH)
    _, p0, n0, _ = ctx.download(False, True, True, False)
    ctx.set_settings(fm.VisualizationSettings(BisectionSteps=8))
    ctx.render(fm.FR_PASS_MARCH)
    _, p1, n1, _ = ctx.download(False, True, True, False)
    assert np.array_equal(p0[..., 3], p1[..., 3])                         # same hit mask
    hit = p0[..., 3] == 1
    campos = cam["position"].astype(np.float64)
    d0 = np.linalg.norm(p0[hit][:, :3] - campos, axis=1)
    d1 = np.linalg.norm(p1[hit][:, :3] - campos, axis=1)
    assert np.all(d1 <= d0 + 1e-5)                                         # moved toward the camera (or stayed)
    # ... by at most the last advance: one StepSize, or StepSize + the empty cell the ray skipped right before the
    # hit (the reference's skip is not conservative: neighbours of an empty cell reach into it, RayMarcher.cpp:282-306)
    assert np.max(d0 - d1) <= 0.009 + 0.1 * np.sqrt(3.0) + 1e-4
    moved = (d0 - d1) > 1e-6                                               # rays with a bracketed sign change
    assert moved.sum() > 20
    rho, _ = ctx.query_density(0, p1[hit][:, :3], want_grad=False)
    rho0, _ = ctx.query_density(0, p0[hit][:, :3], want_grad=False)
    assert np.all(rho >= 1.0)                                             # still on / inside the iso-surface
    assert np.median(rho[moved]) < 1.05 and np.median(rho0[moved]) > 2.0   # and now right at it
    assert np.array_equal(bits(p1[hit][~moved]), bits(p0[hit][~moved]))   # no bracket: the reference's hit, bit for bit
    assert np.all(np.abs(np.linalg.norm(n1[hit][:, :3], axis=1) - 1) < 1e-3)


# ---- (a12) shading and the whole pipeline ------------------------------------------------------------------
def test_full_pipeline_vs_oracle(fm, oracle, gpu_ctx_factory):
    xyz = scenes.dam_break(64000)
    cam = golden_camera("camera_default_16x9")
    W, H = 1280, 720
    f = oracle.frame(xyz, 0.1, 2.0)
    depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
    pos, nrm, *_ = f.march(W, H, oracle_lib.Settings(), cam["inv_proj_view"], cam["position"], depth)
    cam_dir = cam["system"].reshape(3, 3)[2]
    color, rgba = oracle.shade(W, H, pos, nrm, cam["inv_proj_view"], cam["position"], cam_dir)
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    set_cam(ctx, cam)
    ctx.render(fm.FR_PASS_ALL)
    gd, gp, gn, gc = ctx.download()
    assert np.array_equal(bits(gd), bits(depth))
    assert np.array_equal(bits(gp), bits(pos)) and np.array_equal(bits(gn), bits(nrm))
    diff = np.abs(gc.astype(np.int16) - rgba.astype(np.int16))
    assert diff.max() <= 1                                         # tolerance: one 8-bit code value
    assert (diff > 0).mean() < 0.01
    assert len(np.unique(gc.reshape(-1, 4), axis=0)) > 50          # not a constant image


def test_tile_partition_union_is_bit_identical(fm, gpu_ctx_factory):
    """tile-parallel multi-GPU = pure partitioning: the union of the ranks' tiles equals the 1-GPU image"""
    xyz = scenes.dam_break(20000)
    cam = golden_camera("camera_close_16x9")
    W, H = 640, 360
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    set_cam(ctx, cam)
    ctx.render(fm.FR_PASS_ALL)
    full = ctx.download()
    world = 4
    acc = [np.zeros_like(a) for a in full]
    ty, tx = np.meshgrid(np.arange(H) // 64, np.arange(W) // 64, indexing="ij")
    tile = ty * ((W + 63) // 64) + tx
    for rank in range(world):
        c2 = gpu_ctx_factory(W, H)
        c2.upload_frame(0, xyz, 0.1, 2.0)
        set_cam(c2, cam)
        c2.set_tile_partition(rank, world, 64, 64)
        c2.render(fm.FR_PASS_ALL)
        part = c2.download()
        m = (tile % world) == rank
        for a, p in zip(acc, part):            # depth too: the pre-pass of a rank is complete on the pixels it owns
            a[m] = p[m]
        c2.close()
    for a, b in zip(acc, full):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


# ---- (b) the drop-in seam -------------------------------------------------------------------------------------
def test_raymarcher_drop_in_protocol(fm, small):
    """Prepare / Start / IsDone / Exit with caller-owned buffers, as AdvancedRenderer::Render drives it
    (AdvancedRenderer.cpp:257-298)"""
    W, H = int(small["W"]), int(small["H"])
    g = golden_camera("camera_close_16x9")

    class Cam:    # what the engine hands over: CameraController3D with its Camera3D
        pass
    ctl, cam = Cam(), Cam()
    cam.View, cam.Projection, cam.InvProjectionView = g["view"], g["proj"], g["inv_proj_view"]
    ctl.Camera, ctl.Position, ctl.System = cam, g["position"], g["system"]

    ds = fm.Dataset([small["xyz"]], particleRadius=0.1, particleRadiusMultiplier=2.0)
    rm = fm.RayMarcher((W, H))
    positions = np.full((H, W, 4), 9.0, np.float32)
    normals = np.full((H, W, 4), 9.0, np.float32)
    rm.Prepare(fm.VisualizationSettings(Frame=0), ctl, ds, positions, normals, small["depth"])
    rm.Start()
    import time
    t0 = time.time()
    while not rm.IsDone():
        assert time.time() - t0 < 30
        time.sleep(0.001)
    assert np.array_equal(bits(positions), bits(small["positions"]))
    assert np.array_equal(bits(normals), bits(small["normals"]))
    rm.Exit()


def test_cpp_shim_drop_in_matches_golden(fm, small, tmp_path):
    """the C++ class of host/RayMarcher.h (the header that replaces the reference's RayMarcher.h), driven by
    host/shim_driver.cpp the way AdvancedRenderer::Render drives the reference: Prepare / Start / poll IsDone"""
    import struct
    import subprocess
    drv = os.path.join(os.path.dirname(fm.LIB_PATH), "host", "shim_driver")
    assert os.path.exists(drv), "host/shim_driver is not built (python -c 'import __graft_entry__ as g; g.build()')"
    W, H = int(small["W"]), int(small["H"])
    g = golden_camera("camera_close_16x9")
    xyz = np.ascontiguousarray(small["xyz"], np.float32)
    frame = 2                                                       # any frame index of the dataset
    blob = struct.pack("<4i", W, H, len(xyz), frame + 1)
    blob += struct.pack("<iiffi3fi", frame, 128, 0.009, 1.0, 0, 0.5, 2.0, 2000.0, 1)   # VisualizationSettings (bool padded to 4)
    for k in ("view", "proj", "inv_proj_view", "position", "system"):
        blob += np.ascontiguousarray(g[k], np.float32).tobytes()
    blob += struct.pack("<2f", 0.1, 2.0) + xyz.tobytes() + np.ascontiguousarray(small["depth"], np.float32).tobytes()
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    fin.write_bytes(blob)
    r = subprocess.run([drv, str(fin), str(fout)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = np.frombuffer(fout.read_bytes()[:-4], np.float32).reshape(2, H, W, 4)
    assert np.array_equal(bits(out[0]), bits(small["positions"]))
    assert np.array_equal(bits(out[1]), bits(small["normals"]))
    # the interop-style call (depth made by the CUDA pre-pass, caller's depth pointer ignored) gives the same image
    r = subprocess.run([drv, str(fin), str(fout), "gpu_depth"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out2 = np.frombuffer(fout.read_bytes()[:-4], np.float32).reshape(2, H, W, 4)
    assert np.array_equal(bits(out2[0]), bits(small["positions"]))


# ---- size-independent properties at the full BASELINE size (C2: 1M particles, 1920x1080) ---------------------
def test_full_size_properties(fm, gpu_ctx_factory):
    xyz = scenes.dam_break(1_000_000)
    cam = golden_camera("camera_default_16x9")
    W, H = 1920, 1080
    ctx = gpu_ctx_factory(W, H)
    ctx.set_count_mode(fm.FR_COUNT_CELL_EXACT)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    g = ctx.download_frame(0)
    assert int(g["grid_counts"].sum()) == len(xyz)                       # cell-exact reading: every particle counted once
    ctx.set_count_mode(fm.FR_COUNT_CENTRE_BOX)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    g = ctx.download_frame(0)
    assert abs(int(g["grid_counts"].sum()) - len(xyz)) < 100             # centre boxes overlap / leave gaps by an ulp
    assert int(g["cell_start"][-1]) == len(xyz)
    idx = g["sorted_index"].astype(np.int64)
    assert np.array_equal(np.sort(idx), np.arange(len(xyz)))             # a permutation
    ks = search_key(xyz, 0.1, g["info"])[idx]
    assert np.all(np.diff(ks) >= 0)                                      # sortedness
    assert np.array_equal(g["grid_flags"], (g["grid_counts"] > 0).astype(np.uint8))
    set_cam(ctx, cam)
    ctx.render(fm.FR_PASS_ALL)
    a = ctx.download()
    c1 = ctx.counters()
    ctx.upload_frame(0, xyz[::-1].copy(), 0.1, 2.0)                       # input order must not matter ...
    ctx.render(fm.FR_PASS_ALL)
    b = ctx.download()
    assert np.array_equal(bits(a[0]), bits(b[0]))                        # ... for depth (exact min)
    assert np.array_equal(a[1][..., 3], b[1][..., 3]) or (a[1][..., 3] != b[1][..., 3]).mean() < 1e-5
    ctx.upload_frame(0, xyz, 0.1, 2.0)                                   # idempotence: same input, same bits
    ctx.render(fm.FR_PASS_ALL)
    c = ctx.download()
    for x, y in zip(a, c):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    assert c1["covered_rays"] > 100_000 and c1["hit_rays"] > 0.8 * c1["covered_rays"]
    assert c1["neighbour_overflow"] == 0
    # hits lie on the far side of the seed depth: unprojected hit is never in front of the pre-pass surface
    hit = a[1][..., 3] == 1
    assert np.all(np.isfinite(a[1][hit])) and np.all(np.abs(np.linalg.norm(a[2][hit][:, :3], axis=1) - 1) < 1e-3)


def test_ieee_shortcuts_selftest(fm, gpu_ctx_factory):
    """the sequences that claim the bits of IEEE division / square root / reciprocal without the library's range checks:
    shared-reciprocal quotients on random operands, sqrt_rn_normal / rcp_rn_normal on EVERY float in [2^-100, 2^100]"""
    ctx = gpu_ctx_factory(64, 64)
    assert ctx.selftest_division(1 << 22, 3) == 0


def _look_at(eye, target, fov_deg, aspect):
    """the matrices the reference's camera classes produce (glm::lookAtLH-style view, perspectiveLH_ZO with y flipped)"""
    eye, target = np.asarray(eye, np.float64), np.asarray(target, np.float64)
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross([0.0, 1.0, 0.0], fwd)
    right /= np.linalg.norm(right)
    up = np.cross(fwd, right)
    view = np.eye(4)
    view[0, :3], view[1, :3], view[2, :3] = right, up, fwd
    view[:3, 3] = -view[:3, :3] @ eye
    n, f_ = 0.1, 1000.0
    t = 1.0 / np.tan(np.radians(fov_deg) / 2)
    proj = np.zeros((4, 4))
    proj[0, 0], proj[1, 1], proj[2, 2], proj[2, 3], proj[3, 2] = t / aspect, -t, f_ / (f_ - n), -(f_ * n) / (f_ - n), 1.0
    ipv = np.linalg.inv(proj @ view)
    colmajor = lambda m: np.ascontiguousarray(m.T, dtype=np.float32).reshape(-1)
    return colmajor(view), colmajor(proj), colmajor(ipv), eye.astype(np.float32), fwd.astype(np.float32)


@pytest.mark.parametrize("seed", [0, 1])
def test_floor_squares_found_approximately_equal_the_exact_ones(fm, gpu_ctx_factory, seed, monkeypatch):
    """k_classify's background_fast (the checker square of an uncovered pixel from an approximate ray + error bound)
    against shade_pixel everywhere (FLUIDMARCH_BGFAST=0), for cameras all around the scene, near the floor, looking up"""
    rng = np.random.default_rng(seed)
    xyz = scenes.dam_break(8000, t=0.6)
    W, H = 640, 360
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    for k in range(10):
        ang, rad, hgt = rng.uniform(0, 2 * np.pi), rng.uniform(2.0, 40.0), rng.uniform(-0.9, 25.0)
        eye = [rad * np.cos(ang), hgt, rad * np.sin(ang)]
        target = [rng.uniform(-1, 1), rng.uniform(-1.5, 3.0), rng.uniform(-1, 1)]
        cam = _look_at(eye, target, rng.uniform(25, 100), W / H)
        ctx.set_camera(*cam)
        images = []
        for mode in ("0", "1"):
            monkeypatch.setenv("FLUIDMARCH_BGFAST", mode)
            ctx.render(fm.FR_PASS_ALL)
            images.append(ctx.download()[3])
        assert np.array_equal(images[0], images[1]), (k, eye, target)
        assert len(np.unique(images[0].reshape(-1, 4), axis=0)) >= 2
