"""Frame builds that do not wait for the host (builds into the tables of an earlier frame of the same slot read the grid
parameters from device memory; the host picks them up from a side stream when it launches the march), and what the host
used to check in its round trip: non-finite particle coordinates, tables too small for the new bounds (rebuild),
collapsed cells.  Also fr_seq_wait's
per-ticket status.  Everything through the C ABI."""
import importlib
import os

import numpy as np
import pytest

from conftest import golden_camera

pytestmark = pytest.mark.gpu
scenes = importlib.import_module("bachelor-thesis_b200.scenes")
W, H = 320, 180


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def set_cam(ctx, cam):
    ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])


def render_fresh(fm, xyz, cam):
    c = fm.Context(W, H)
    try:
        set_cam(c, cam)
        c.upload_frame(0, xyz, 0.1, 2.0)                 # first build of the slot: host-sized tables
        c.render(fm.FR_PASS_ALL)
        return c.download(), c.download_frame(0), c.counters()
    finally:
        c.close()


def test_build_into_existing_tables_is_bit_identical(fm, gpu_ctx_factory):
    cam = golden_camera("camera_close_16x9")
    a, b = scenes.dam_break(20000, t=0.5), scenes.dam_break(20000, t=0.55)
    want_img, want_grid, want_cnt = render_fresh(fm, b, cam)
    ctx = gpu_ctx_factory(W, H)
    set_cam(ctx, cam)
    ctx.upload_frame(0, a, 0.1, 2.0)
    ctx.render(fm.FR_PASS_ALL)
    ctx.upload_frame(0, b, 0.1, 2.0)                     # same slot: no host wait inside this build
    ctx.render(fm.FR_PASS_ALL)
    got_img, got_grid, got_cnt = ctx.download(), ctx.download_frame(0), ctx.counters()
    for x, y in zip(got_img, want_img):
        assert np.array_equal(bits(x), bits(y))
    for k in ("sorted_xyz", "sorted_index", "cell_start", "grid_counts", "grid_flags"):
        assert np.array_equal(bits(got_grid[k]), bits(want_grid[k])), k
    for k in ("min", "max", "grid_dims", "search_min", "search_dims"):
        assert np.array_equal(bits(got_grid["info"][k]), bits(want_grid["info"][k])), k
    assert got_grid["info"]["occupied_cells"] == want_grid["info"]["occupied_cells"]
    for k in ("covered_rays", "hit_rays", "ray_steps", "skip_iterations", "candidates", "neighbours"):
        assert got_cnt[k] == want_cnt[k], k
    # the round-1 behaviour (every build waits once) is still there, same bits
    ctx.set_async_build(False)
    ctx.upload_frame(0, b, 0.1, 2.0)
    ctx.render(fm.FR_PASS_ALL)
    for x, y in zip(ctx.download(), want_img):
        assert np.array_equal(bits(x), bits(y))
    ctx.set_async_build(True)


def test_tables_too_small_are_rebuilt_and_the_render_repeated(fm, gpu_ctx_factory):
    cam = golden_camera("camera_default_16x9")
    small = scenes.random_block(2000, 0.3)               # 8 x 8 x 8 cells
    big = scenes.dam_break(64000)                        # a few thousand cells: the slot's tables cannot hold it
    want_img, want_grid, _ = render_fresh(fm, big, cam)
    ctx = gpu_ctx_factory(W, H)
    set_cam(ctx, cam)
    ctx.upload_frame(0, small, 0.1, 2.0)
    ctx.render(fm.FR_PASS_ALL)
    ctx.upload_frame(0, big, 0.1, 2.0)                   # queued against the small tables: FM_GRID_OVERFLOW on the device
    ctx.render_async(fm.FR_PASS_ALL)                     # picks the parameters up, rebuilds the frame, repeats the pre-pass
    got_img = ctx.download()
    for x, y in zip(got_img, want_img):
        assert np.array_equal(bits(x), bits(y))
    got_grid = ctx.download_frame(0)
    assert np.array_equal(got_grid["cell_start"], want_grid["cell_start"])
    assert np.array_equal(bits(got_grid["sorted_xyz"]), bits(want_grid["sorted_xyz"]))
    # and back to a small frame in the now large tables
    want_small, _, _ = render_fresh(fm, small, cam)
    ctx.upload_frame(0, small, 0.1, 2.0)
    ctx.render(fm.FR_PASS_ALL)
    for x, y in zip(ctx.download(), want_small):
        assert np.array_equal(bits(x), bits(y))


@pytest.mark.parametrize("poison", [np.nan, np.inf, -np.inf])
def test_non_finite_particles_are_rejected(fm, gpu_ctx_factory, poison, tmp_path):
    """one NaN coordinate used to index the cell histogram out of bounds"""
    cam = golden_camera("camera_close_16x9")
    good = scenes.dam_break(8000)
    bad = good.copy()
    bad[1234, 1] = poison
    ctx = gpu_ctx_factory(W, H)
    set_cam(ctx, cam)
    with pytest.raises(fm.FluidMarchError, match="NaN or infinite"):       # first build of the slot: reported at once
        ctx.upload_frame(0, bad, 0.1, 2.0)
    ctx.upload_frame(0, good, 0.1, 2.0)
    ctx.render(fm.FR_PASS_ALL)
    want = ctx.download()
    ctx.upload_frame(0, bad, 0.1, 2.0)                    # tables exist: the build is only queued ...
    with pytest.raises(fm.FluidMarchError, match="NaN or infinite"):       # ... and the render that needs the frame reports it
        ctx.render_async(fm.FR_PASS_ALL)
    ctx.wait()
    path = str(tmp_path / "bad.bgeo")
    fm.bgeo_write(path, bad)
    ctx.upload_frame_bgeo(0, path, 0.1, 2.0)
    with pytest.raises(fm.FluidMarchError, match="NaN or infinite"):
        ctx.frame_info(0)
    # the context is still usable and nothing was corrupted
    ctx.upload_frame(0, good, 0.1, 2.0)
    ctx.render(fm.FR_PASS_ALL)
    for x, y in zip(ctx.download(), want):
        assert np.array_equal(bits(x), bits(y))
    other = gpu_ctx_factory(W, H)
    set_cam(other, cam)
    with pytest.raises(fm.FluidMarchError, match="NaN or infinite"):
        other.upload_frame_bgeo(0, path, 0.1, 2.0)


@pytest.mark.parametrize("crowd", [1500, 5000, 40000])
def test_collapsed_cell_is_ordered_by_a_sort(fm, oracle, gpu_ctx_factory, crowd):
    """the in-cell ranking is quadratic; a cell with more than 2048 particles (a collapsed simulation) is
    sorted by one CTA instead -- same order (ascending original index inside every cell, both searches), same sums"""
    rng = np.random.default_rng(crowd)
    blob = (np.float32(0.01234) + rng.uniform(0.0, 0.02, (crowd, 3))).astype(np.float32)      # inside one h-cell
    xyz = np.concatenate([scenes.dam_break(8000), blob])
    xyz = xyz[rng.permutation(len(xyz))]
    ctx = gpu_ctx_factory(64, 64)
    for _ in range(2):                                   # host-sized tables, then a build that does not wait
        ctx.upload_frame(0, xyz, 0.1, 2.0)
        g = ctx.download_frame(0)
        idx = g["sorted_index"].astype(np.int64)
        assert np.array_equal(np.sort(idx), np.arange(len(xyz)))
        cs = g["cell_start"].astype(np.int64)
        assert (np.diff(cs) > 2048).any() == (crowd > 2048)
        inner = np.ones(len(idx), bool)
        inner[cs[1:-1][cs[1:-1] < len(idx)]] = False     # first slot of every later cell
        inner[0] = False
        assert (np.diff(idx)[inner[1:]] > 0).all(), "indices ascend inside every cell"
        assert np.array_equal(xyz[idx], g["sorted_xyz"])
    # the ordered sums over the blob (up to MAX_NEIGHBORS = 8192 of its particles) against the oracle
    pts = np.array([[0.02, 0.02, 0.02], [0.05, 0.0, 0.03], [-0.3, -0.8, 0.1]], np.float32)
    rho, grad = ctx.query_density(0, pts)
    f = oracle.frame(xyz, 0.1, 2.0)
    perm = f.particles()
    for i, p in enumerate(pts):
        want, gw = np.float32(0), np.zeros(3, np.float32)
        for j in f.neighbors(p, cap=8192):               # MAX_NEIGHBORS (RayMarcher.cpp:14): the list is cut there
            r = (perm[j] - p).astype(np.float32)
            want = np.float32(want + oracle.W(0.1, r))
            gw = (gw + oracle.gradW(0.1, r)).astype(np.float32)
        assert bits(rho[i:i + 1])[0] == bits(np.array([want]))[0], (i, rho[i], want)
        assert np.array_equal(bits(grad[i]), bits(gw)), i
    # the r = 2h search
    srt, cs_ext, _, _ = ctx.download_frame_ext(0)
    ids = srt[:, 3].view(np.uint32).astype(np.int64)
    assert np.array_equal(np.sort(ids), np.arange(len(xyz)))
    cs_ext = cs_ext.astype(np.int64)
    inner = np.ones(len(ids), bool)
    inner[cs_ext[1:-1][cs_ext[1:-1] < len(ids)]] = False
    inner[0] = False
    assert (np.diff(ids)[inner[1:]] > 0).all()
    assert np.array_equal(xyz[ids], srt[:, :3])


def test_many_search_cells_scan_in_one_pass(fm, oracle, gpu_ctx_factory):
    """sparse particles over a large extent: ~1.3 M search cells = 300+ scan tiles through the decoupled look-back"""
    rng = np.random.default_rng(3)
    xyz = rng.uniform(-5.5, 5.5, (30000, 3)).astype(np.float32)
    ctx = gpu_ctx_factory(64, 64)
    for _ in range(2):                                    # host-sized, then device-checked
        ctx.upload_frame(0, xyz, 0.1, 2.0)
        g = ctx.download_frame(0)
        info = g["info"]
        inv = np.float32(1.0) / np.float32(0.1)
        t = (inv * xyz).astype(np.int32)
        k = np.where(xyz >= 0, t, t - 1) - info["search_min"][None, :]
        kd = info["search_dims"].astype(np.int64)
        key = (k[:, 0].astype(np.int64) * kd[1] + k[:, 1]) * kd[2] + k[:, 2]
        cells = int(np.prod(kd))
        assert cells > 1_000_000
        want = np.concatenate([[0], np.cumsum(np.bincount(key, minlength=cells))])
        assert np.array_equal(g["cell_start"].astype(np.int64), want)
        counts, flags = oracle.frame(xyz, 0.1, 2.0).grid()
        assert np.array_equal(g["grid_counts"], counts) and np.array_equal(g["grid_flags"], flags)


def test_seq_wait_reports_the_status_of_that_ticket(fm):
    """a failed frame must not read as FR_OK once a later frame of its lane has finished"""
    cam = golden_camera("camera_close_16x9")
    seq = fm.Sequence(W, H, lanes=2)
    try:
        seq.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
        seq.set_settings(fm.VisualizationSettings())
        good = scenes.dam_break(8000)
        bad = good.copy()
        bad[5, 0] = np.nan
        tickets = []
        for k in range(8):                                # tickets 2 and 5 fail; later frames of their lanes succeed
            t, out = seq.submit(bad if k in (2, 5) else good)
            tickets.append((t, out))
        seq.wait(tickets[-1][0])
        seq.wait(tickets[-2][0])                          # both lanes are past the failed tickets now
        for k, (t, out) in enumerate(tickets):
            if k in (2, 5):
                with pytest.raises(fm.FluidMarchError, match="NaN or infinite"):
                    seq.wait(t)
            else:
                seq.wait(t)
                assert out["rgba"].any()
        with pytest.raises(fm.FluidMarchError):
            seq.drain()
        seq.drain()
    finally:
        seq.close()


@pytest.mark.parametrize("cam_name", ["camera_close_16x9", "camera_orbit_a_16x9"])
def test_region_partition_is_bit_identical_and_filters_particles(fm, gpu_ctx_factory, cam_name):
    """fr_set_region_partition: a context that renders one pixel rectangle builds its frame from the particles that can
    influence it -- same grid geometry, so the rectangle's pixels equal the unpartitioned render's bit for bit"""
    Wr, Hr = 640, 384
    cam = golden_camera(cam_name)
    xyz = scenes.dam_break(64000)
    full = gpu_ctx_factory(Wr, Hr)
    set_cam(full, cam)
    full.upload_frame(0, xyz, 0.1, 2.0)
    full.render(fm.FR_PASS_ALL)
    want = full.download()
    want_cnt = full.counters()
    part = gpu_ctx_factory(Wr, Hr)
    set_cam(part, cam)
    regions = [(0, 0, 256, 384), (256, 0, 448, 192), (256, 192, 448, 384), (448, 0, 640, 384)]
    covered = 0
    steps = 0
    for (x0, y0, x1, y1) in regions:
        part.set_region_partition(x0, y0, x1, y1)
        for _ in range(2):                                   # host-sized tables, then the build that does not wait
            part.upload_frame(0, xyz, 0.1, 2.0)
            part.render(fm.FR_PASS_ALL)
            got = part.download()
            for a, b in zip(got, want):
                assert np.array_equal(bits(a[y0:y1, x0:x1]), bits(b[y0:y1, x0:x1])), (x0, y0, x1, y1)
        c = part.counters()
        covered += c["covered_rays"]
        steps += c["ray_steps"]
    assert covered == want_cnt["covered_rays"] and steps == want_cnt["ray_steps"]
    part.set_region_partition()
    part.upload_frame(0, xyz, 0.1, 2.0)
    part.render(fm.FR_PASS_ALL)
    for a, b in zip(part.download(), want):
        assert np.array_equal(bits(a), bits(b))
    with pytest.raises(fm.FluidMarchError, match="multiples of 64"):
        part.set_region_partition(10, 0, 200, 100)


@pytest.mark.parametrize("shuffle", [False, True])
def test_prepass_beside_the_build_is_bit_identical(fm, gpu_ctx_factory, shuffle):
    """With stage timing off a render queued behind a frame build runs the depth pre-pass on a second stream from the
    raw particle array (render_depth): same depth image, same march, whatever the particle order."""
    cam = golden_camera("camera_close_16x9")
    xyz = scenes.dam_break(30000, t=0.6)
    if shuffle:
        xyz = xyz[np.random.default_rng(5).permutation(len(xyz))]
    want, _, want_counters = render_fresh(fm, xyz, cam)               # stage timing on: one stream
    c = gpu_ctx_factory(W, H)
    set_cam(c, cam)
    c.set_stage_timing(False)
    for k in range(3):                                                # first build (host-sized), then builds that do not wait
        c.upload_frame(0, xyz, 0.1, 2.0)
        c.render_async(fm.FR_PASS_ALL)
        got = c.download()
        for name, g, w in zip(("depth", "positions", "normals", "rgba"), got, want):
            assert np.array_equal(bits(g), bits(w)), (name, k)
    assert c.counters()["hit_rays"] == want_counters["hit_rays"]
    # a second render of the same frame (camera moved, no new build): the pre-pass reads the sorted array again
    cam2 = golden_camera("camera_default_16x9")
    set_cam(c, cam2)
    c.render_async(fm.FR_PASS_ALL)
    got2 = c.download()
    c2 = gpu_ctx_factory(W, H)
    set_cam(c2, cam2)
    c2.upload_frame(0, xyz, 0.1, 2.0)
    c2.render(fm.FR_PASS_ALL)
    for name, g, w in zip(("depth", "positions", "normals", "rgba"), got2, c2.download()):
        assert np.array_equal(bits(g), bits(w)), name


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_call_sequences_render_what_a_fresh_context_renders(fm, gpu_ctx_factory, seed):
    """uploads into two frame slots, camera changes, renders with and without per-stage events (one stream / pre-pass
    beside the build + uncovered pixels beside the long rays), host waits in between, in random order: every image that
    is downloaded equals the one a fresh context renders for that (particles, camera)"""
    rng = np.random.default_rng(seed)
    second = scenes.dam_break(26000, t=0.65)
    data = [scenes.dam_break(20000, t=0.5), second[rng.permutation(len(second))]]
    cams = [golden_camera("camera_close_16x9"), golden_camera("camera_default_16x9")]
    want = {}
    for d in range(2):
        for k in range(2):
            c = gpu_ctx_factory(W, H)
            set_cam(c, cams[k])
            c.upload_frame(0, data[d], 0.1, 2.0)
            c.render(fm.FR_PASS_ALL)
            want[d, k] = c.download()
    ctx = gpu_ctx_factory(W, H)
    slot_data = {}
    cam_now, frame_now = 0, 0
    set_cam(ctx, cams[0])
    checked = 0
    for step in range(140):
        op = rng.integers(0, 7)
        if op == 0 or not slot_data:
            slot, d = int(rng.integers(0, 2)), int(rng.integers(0, 2))
            ctx.upload_frame(slot, data[d], 0.1, 2.0)
            slot_data[slot] = d
        elif op == 1:
            cam_now = int(rng.integers(0, 2))
            set_cam(ctx, cams[cam_now])
        elif op == 2:
            ctx.set_stage_timing(bool(rng.integers(0, 2)))
        elif op == 3:
            ctx.frame_info(int(rng.choice(list(slot_data))))            # a host wait between a build and its render
        elif op == 4:
            ctx.wait()
        else:
            frame_now = int(rng.choice(list(slot_data)))
            s = fm.VisualizationSettings()
            s.Frame = frame_now
            ctx.set_settings(s)
            ctx.render_async(fm.FR_PASS_ALL)
            if rng.integers(0, 3):
                got = ctx.download()
                for name, g, w in zip(("depth", "positions", "normals", "rgba"), got, want[slot_data[frame_now], cam_now]):
                    assert np.array_equal(bits(g), bits(w)), (name, step, slot_data[frame_now], cam_now)
                checked += 1
    assert checked > 15
