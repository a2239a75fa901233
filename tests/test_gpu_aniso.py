"""GPU parity of the anisotropic path: PerPixel_Anisotropic (RayMarcher.cpp:346-423), WPCA
(:114-254), Eigen computeDirect, AnisotropicKernel (Kernel.cpp:55-107) and the r = h_ext search
(Dataset.cpp:65-75,282-290) on the CUDA path, through the C ABI, against the CPU oracle and against the golden fixture
the reference itself produced (tests/golden/aniso8k_160x90.npz).

Parity bar: bit-exact -- 2h neighbour lists (ids and order), G, densities, hit mask, hit positions, normals.  The
device runs the reference's FP32 operations in the reference's order, including glibc's atan2f/sinf/cosf inside the
eigen solver (fm_aniso.cuh), so no tolerance is needed."""
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


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "aniso8k_160x90.npz"))


ANISO = oracle_lib.Settings(anisotropic=1)


def test_ext_search_structure(fm, gpu_ctx_factory):
    """the r = h_ext counting sort: a permutation grouped by 2h cell key (z fastest), ascending original id per cell"""
    for n, gen in ((8000, "dam"), (5000, "rand"), (1, "rand")):
        xyz = scenes.dam_break(n) if gen == "dam" else scenes.random_block(n, 0.7)
        ctx = gpu_ctx_factory(64, 64)
        ctx.upload_frame(0, xyz, 0.1, 2.0)
        srt, cell_start, kmin, kdim = ctx.download_frame_ext(0)
        idx = srt[:, 3].copy().view(np.uint32).astype(np.int64)
        assert np.array_equal(np.sort(idx), np.arange(len(xyz)))
        assert np.array_equal(bits(srt[:, :3]), bits(xyz[idx]))
        inv = np.float32(1.0) / (np.float32(2.0) * np.float32(0.1))
        t = (inv * xyz.astype(np.float32)).astype(np.int32)
        k = np.where(xyz >= 0, t, t - 1) - kmin[None, :]
        assert (k >= 0).all() and (k < kdim[None, :]).all()
        assert (k.min(axis=0) == 0).all() and (k.max(axis=0) == kdim - 1).all()
        key = (k[:, 0].astype(np.int64) * kdim[1] + k[:, 1]) * kdim[2] + k[:, 2]
        ks = key[idx]
        assert np.all(np.diff(ks) >= 0)
        assert np.all(np.diff(idx)[np.diff(ks) == 0] > 0)
        cells = int(np.prod(kdim.astype(np.int64)))
        assert np.array_equal(cell_start.astype(np.int64), np.concatenate([[0], np.cumsum(np.bincount(key, minlength=cells))]))


def test_ext_neighbour_lists_bit_exact(fm, oracle, gpu_ctx_factory, g):
    """Dataset::GetNeighborsExt: same particles in the same order as the oracle and as the reference (golden)"""
    xyz = scenes.dam_break(8000)
    ctx = gpu_ctx_factory(64, 64)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    pts = g["wpca_p"]
    counts, ids = ctx.query_neighbors(0, pts, cap=1024, ext=True)
    assert np.array_equal(counts.astype(np.int64), g["wpca_len"].astype(np.int64))
    off = 0
    for i, n in enumerate(g["wpca_len"]):
        assert np.array_equal(bits(xyz[ids[i, :n]]), bits(g["wpca_xyz"][off:off + n]))
        off += n
    # brute force: exactly the set {j : |x_j - p|^2 < h_ext^2} with the FP32 expression of the search
    r2 = np.float32(0.2) * np.float32(0.2)
    for i in range(0, len(pts), 7):
        d = pts[i][None, :] - xyz
        l2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        assert set(ids[i, :counts[i]].tolist()) == set(np.nonzero(l2 < r2)[0].tolist())
    # the r = h query is unaffected by the second structure
    c1, _ = ctx.query_neighbors(0, pts, cap=0)
    f = oracle.frame(xyz, 0.1, 2.0)
    assert np.array_equal(c1, np.array([len(f.neighbors(p)) for p in pts], np.uint32))


def _oracle_point(oracle, f, perm_ext, p, s):
    ids = f.neighbors(p, ext=1)
    nb = perm_ext[ids]
    if len(nb) == 0:
        return None, np.float32(0), np.zeros(3, np.float32)
    G = oracle.wpca(0.1, 0.2, s, p, nb)
    det = oracle.det3(G)
    rho = np.float32(0)
    grad = np.zeros(3, np.float32)
    h2 = np.float32(0.1) * np.float32(0.1)
    for x in nb:
        r = (x - p).astype(np.float32)
        if np.float32(np.float32(r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]) < h2:
            rho = np.float32(rho + oracle.aniso_W(0.1, G, det, r))
            with np.errstate(all="ignore"):
                grad = (grad + oracle.aniso_gradW(0.1, G, det, r)).astype(np.float32)
    return G, rho, grad


@pytest.mark.parametrize("settings", [ANISO, oracle_lib.Settings(anisotropic=1, k_n=0.4, k_r=4.0, k_s=1400.0, n_eps=25)])
def test_wpca_and_kernel_sums_bit_exact(fm, oracle, gpu_ctx_factory, g, settings):
    """G = WPCA(p, GetNeighborsExt(p)), sum W and sum gradW per query point against the oracle's functions; for the
    default settings G is also the REFERENCE's own G (golden)"""
    xyz = scenes.dam_break(8000)
    ctx = gpu_ctx_factory(64, 64)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    ctx.set_settings(fm.VisualizationSettings(EnableAnisotropy=True, k_n=settings.k_n, k_r=settings.k_r, k_s=settings.k_s,
                                              N_eps=settings.n_eps))
    pts = g["wpca_p"]
    rho, grad, g9 = ctx.query_anisotropic(0, pts)
    f = oracle.frame(xyz, 0.1, 2.0)
    perm_ext = f.particles_ext()
    nonzero = 0
    for i, p in enumerate(pts):
        G, want_rho, want_grad = _oracle_point(oracle, f, perm_ext, p, settings)
        if G is None:
            assert rho[i] == 0.0
            continue
        assert np.array_equal(bits(g9[i]), bits(G)), i
        if settings is ANISO:
            assert np.array_equal(bits(g9[i]), bits(g["wpca_G"][i])), i
        assert bits(rho[i]) == bits(want_rho), i
        assert np.array_equal(bits(grad[i]), bits(want_grad)), i
        nonzero += rho[i] > 0
    assert nonzero > 40


def test_march_matches_golden_reference_output(fm, gpu_ctx_factory, g):
    """positions / normals the REFERENCE's PerPixel_Anisotropic produced (tests/golden), bit for bit"""
    xyz = scenes.dam_break(8000)
    cam = golden_camera("camera_close_16x9")
    W, H = int(g["W"]), int(g["H"])
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    set_cam(ctx, cam)
    ctx.set_settings(fm.VisualizationSettings(EnableAnisotropy=True))
    ctx.set_depth(g["depth"])
    ctx.render(fm.FR_PASS_MARCH)
    _, pos, nrm, _ = ctx.download(False, True, True, False)
    assert np.array_equal(pos[..., 3], g["positions"][..., 3])
    assert np.array_equal(bits(pos), bits(g["positions"]))
    assert np.array_equal(bits(nrm), bits(g["normals"]))
    c = ctx.counters()
    assert c["hit_rays"] == int(g["positions"][..., 3].sum()) > 1500
    assert c["neighbour_overflow"] == 0
    # the depth pre-pass of the device is the fixture's depth image, so the whole pipeline reproduces it too
    ctx.render(fm.FR_PASS_ALL)
    d2, p2, n2, rgba = ctx.download()
    assert np.array_equal(bits(d2), bits(g["depth"]))
    assert np.array_equal(bits(p2), bits(g["positions"])) and np.array_equal(bits(n2), bits(g["normals"]))
    assert rgba.any()


@pytest.mark.parametrize("n,W,H,cam_name,kw", [
    (20000, 320, 180, "camera_orbit_a_16x9", {}),
    (20000, 333, 187, "camera_orbit_b_16x9", dict(step_size=0.02, max_steps=60)),
    (8000, 160, 90, "camera_close_16x9", dict(k_n=0.4, k_r=4.0, k_s=1400.0, n_eps=25)),
    (8000, 160, 90, "camera_close_16x9", dict(iso_density=3.0, step_size=0.005)),      # deeper threshold: many steps per ray
    (8000, 160, 90, "camera_close_16x9", dict(iso_density=1e6)),                         # never reached
    (64000, 1280, 720, "camera_default_16x9", {}),                                       # BASELINE config C1, anisotropic
])
def test_march_bit_exact_vs_oracle(fm, oracle, gpu_ctx_factory, n, W, H, cam_name, kw):
    xyz = scenes.dam_break(n)
    cam = golden_camera(cam_name)
    f = oracle.frame(xyz, 0.1, 2.0)
    depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
    s = oracle_lib.Settings(anisotropic=1, **kw)
    pos, nrm, band, steps, cnt = f.march(W, H, s, cam["inv_proj_view"], cam["position"], depth)
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    set_cam(ctx, cam)
    ctx.set_settings(fm.VisualizationSettings(EnableAnisotropy=True, MaxSteps=s.max_steps, StepSize=s.step_size,
                                              IsoDensity=s.iso_density, k_n=s.k_n, k_r=s.k_r, k_s=s.k_s, N_eps=s.n_eps))
    ctx.set_depth(depth)
    ctx.render(fm.FR_PASS_MARCH)
    _, gpos, gnrm, _ = ctx.download(False, True, True, False)
    c = ctx.counters()
    assert c["covered_rays"] == cnt["covered_rays"] and c["hit_rays"] == cnt["hit_rays"]
    assert np.array_equal(gpos[..., 3], pos[..., 3])
    assert np.array_equal(bits(gpos), bits(pos))
    assert np.array_equal(bits(gnrm), bits(nrm))
    assert c["ray_steps"] <= cnt["ray_steps"]
    assert c["ray_steps"] >= cnt["ray_steps"] - cnt["steps_outside_grid"]
    assert c["skip_iterations"] == cnt["skip_iterations"]
    assert c["neighbour_overflow"] == 0


def test_anisotropic_and_isotropic_share_a_context(fm, oracle, gpu_ctx_factory):
    """switching EnableAnisotropy back and forth on one context / one uploaded frame gives each path's own image"""
    xyz = scenes.dam_break(8000)
    cam = golden_camera("camera_close_16x9")
    W, H = 160, 90
    f = oracle.frame(xyz, 0.1, 2.0)
    depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
    want = {a: f.march(W, H, oracle_lib.Settings(anisotropic=a), cam["inv_proj_view"], cam["position"], depth)[:2] for a in (0, 1)}
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    set_cam(ctx, cam)
    for a in (1, 0, 1, 0):
        ctx.set_settings(fm.VisualizationSettings(EnableAnisotropy=bool(a)))
        ctx.render(fm.FR_PASS_ALL)
        _, p, n, _ = ctx.download(False, True, True, False)
        assert np.array_equal(bits(p), bits(want[a][0])) and np.array_equal(bits(n), bits(want[a][1]))
    # a re-uploaded frame rebuilds its r = h_ext search
    xyz2 = scenes.dam_break(6000, seed=5)
    f2 = oracle.frame(xyz2, 0.1, 2.0)
    depth2 = f2.depth_prepass(W, H, cam["view"], cam["proj"])
    p2, n2, *_ = f2.march(W, H, ANISO, cam["inv_proj_view"], cam["position"], depth2)
    ctx.upload_frame(0, xyz2, 0.1, 2.0)
    ctx.set_settings(fm.VisualizationSettings(EnableAnisotropy=True))
    ctx.render(fm.FR_PASS_ALL)
    _, p, n, _ = ctx.download(False, True, True, False)
    assert np.array_equal(bits(p), bits(p2)) and np.array_equal(bits(n), bits(n2))


def test_anisotropic_bisection_and_edge_cases(fm, oracle, gpu_ctx_factory):
    cam = golden_camera("camera_close_16x9")
    W, H = 61, 35
    xyz = scenes.random_block(4000, 0.5)
    f = oracle.frame(xyz, 0.1, 2.0)
    depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    set_cam(ctx, cam)
    for ms in (0, 1, 128):
        s = oracle_lib.Settings(anisotropic=1, max_steps=ms)
        pos, nrm, *_ = f.march(W, H, s, cam["inv_proj_view"], cam["position"], depth)
        ctx.set_settings(fm.VisualizationSettings(EnableAnisotropy=True, MaxSteps=ms))
        ctx.set_depth(depth)
        ctx.render(fm.FR_PASS_MARCH)
        _, p, n, _ = ctx.download(False, True, True, False)
        assert np.array_equal(bits(p), bits(pos)) and np.array_equal(bits(n), bits(nrm))
    # bisection (not in the reference): same hit mask, hits move toward the camera by at most one step
    ctx.set_settings(fm.VisualizationSettings(EnableAnisotropy=True, BisectionSteps=6))
    ctx.render(fm.FR_PASS_MARCH)
    _, pb, nb, _ = ctx.download(False, True, True, False)
    assert np.array_equal(pb[..., 3], p[..., 3])
    hit = p[..., 3] == 1
    cam_pos = cam["position"]
    d0 = np.linalg.norm(p[hit][:, :3] - cam_pos, axis=1)
    d1 = np.linalg.norm(pb[hit][:, :3] - cam_pos, axis=1)
    assert np.all(d1 <= d0 + 1e-5)
    assert np.all(d0 - d1 <= 0.25)            # never farther back than the bracket (an empty-cell jump can be longer than a step)
    assert (d0 - d1 > 1e-4).any()
    nl = np.linalg.norm(nb[hit][:, :3], axis=1)
    assert np.nanmax(np.abs(nl - 1)) < 1e-5


def test_tile_partition_union_is_bit_identical(fm, gpu_ctx_factory):
    xyz = scenes.dam_break(20000)
    cam = golden_camera("camera_close_16x9")
    W, H = 640, 360
    s = fm.VisualizationSettings(EnableAnisotropy=True)
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    set_cam(ctx, cam)
    ctx.set_settings(s)
    ctx.render(fm.FR_PASS_ALL)
    full = ctx.download()
    world = 2
    acc = [np.zeros_like(a) for a in full]
    ty, tx = np.meshgrid(np.arange(H) // 64, np.arange(W) // 64, indexing="ij")
    tile = ty * ((W + 63) // 64) + tx
    for rank in range(world):
        c2 = gpu_ctx_factory(W, H)
        c2.upload_frame(0, xyz, 0.1, 2.0)
        set_cam(c2, cam)
        c2.set_settings(s)
        c2.set_tile_partition(rank, world, 64, 64)
        c2.render(fm.FR_PASS_ALL)
        part = c2.download()
        m = (tile % world) == rank
        for a, p in zip(acc[1:], part[1:]):
            a[m] = p[m]
        c2.close()
    for a, b in zip(acc[1:], full[1:]):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_full_size_properties(fm, gpu_ctx_factory):
    """BASELINE config C2 (1 M particles, 1920x1080), anisotropic: size-independent properties"""
    xyz = scenes.dam_break(1_000_000)
    cam = golden_camera("camera_default_16x9")
    W, H = 1920, 1080
    ctx = gpu_ctx_factory(W, H)
    ctx.upload_frame(0, xyz, 0.1, 2.0)
    set_cam(ctx, cam)
    ctx.set_settings(fm.VisualizationSettings(EnableAnisotropy=True))
    ctx.render(fm.FR_PASS_ALL)
    d, p, n, c = ctx.download()
    cnt = ctx.counters()
    covered = d != 1.0
    hit = p[..., 3] == 1.0
    assert cnt["covered_rays"] == int(covered.sum()) and cnt["hit_rays"] == int(hit.sum())
    assert cnt["neighbour_overflow"] == 0
    assert not hit[~covered].any() and hit.sum() > 0.8 * covered.sum()
    assert set(np.unique(p[..., 3])) <= {0.0, 1.0}
    assert not p[~hit].any() and not n[~hit].any()
    nl = np.linalg.norm(n[hit][:, :3], axis=1)
    assert np.abs(nl - 1).max() < 1e-5
    # hits lie inside the particle AABB padded by h and behind the depth seed
    info = ctx.frame_info(0)
    assert (p[hit][:, :3] >= info["min"] - 1e-4).all() and (p[hit][:, :3] <= info["max"] + 1e-4).all()
    # idempotence: the same frame rendered again gives the same bits (no order dependence on scheduling)
    ctx.render(fm.FR_PASS_ALL)
    d2, p2, n2, c2 = ctx.download()
    assert np.array_equal(bits(p), bits(p2)) and np.array_equal(bits(n), bits(n2)) and np.array_equal(c, c2)
    # the anisotropic surface is not the isotropic one
    ctx.set_settings(fm.VisualizationSettings())
    ctx.render(fm.FR_PASS_ALL)
    _, p3, _, _ = ctx.download()
    assert not np.array_equal(bits(p), bits(p3))
    t = ctx.timings()
    assert t["march_ms"] > 0
