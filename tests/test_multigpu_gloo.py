"""world_size-2 (and 3) gloo tests of the multi-GPU host logic (bachelor-thesis_b200/multigpu.py): the tile-parallel
gather and the frame-parallel collection.  The per-rank "render" is a deterministic stand-in that only fills the
pixels the rank owns (what fr_set_tile_partition makes the kernels do); the GPU version of the same property is
tests/test_gpu_parity.py::test_tile_partition_union_is_bit_identical."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

mg = importlib.import_module("bachelor-thesis_b200.multigpu")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _full_image(W, H, frame=0):
    y, x = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    img = np.stack([(x * 7 + y * 13 + frame) % 251, (x ^ y) % 256, (x + 3 * y + 5 * frame) % 256, np.full_like(x, 255)], -1)
    return torch.from_numpy(img.astype(np.uint8))


def _worker(rank, world, port, W, H, n_frames, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = importlib.import_module("bachelor-thesis_b200.multigpu")
        # ---- tile-parallel: each rank fills only its own tiles, everything else is garbage
        owner = torch.from_numpy(m.tile_owner_map(W, H, world, 64, 64))
        full = _full_image(W, H)
        local = torch.full_like(full, 77)
        local[owner == rank] = full[owner == rank]
        merged = m.gather_tiles(local, rank, world, owner, dst=0)
        ok_tiles = (rank != 0 and merged is None) or (rank == 0 and torch.equal(merged, full))
        # ---- frame-parallel: every rank renders its frames, rank 0 ends up with all of them in order
        rendered = []

        def render(f):
            rendered.append(f)
            return _full_image(W, H, f)
        frames = m.collect_frames(render, n_frames, rank, world, like=full, dst=0)
        ok_frames = rendered == m.frames_of_rank(n_frames, rank, world)
        if rank == 0:
            ok_frames = ok_frames and all(torch.equal(frames[f], _full_image(W, H, f)) for f in range(n_frames))
        else:
            ok_frames = ok_frames and frames is None
        q.put((rank, bool(ok_tiles), bool(ok_frames)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tile_gather_and_frame_collection_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 200, 136, 7, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(r, True, True) for r in range(world)]


def test_partition_rules():
    assert mg.frames_of_rank(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((mg.frames_of_rank(240, r, 8) for r in range(8)), [])) == list(range(240))
    own = mg.tile_owner_map(1920, 1080, 8)
    assert own.shape == (1080, 1920) and own.min() == 0 and own.max() == 7
    # interleaving: every rank owns between 1/8 -15 % and 1/8 +15 % of a 1080p frame
    frac = np.bincount(own.ravel(), minlength=8) / own.size
    assert np.all(np.abs(frac - 0.125) < 0.02)
    # same rule as the device code: tile index row-major over 64x64 tiles
    assert own[0, 0] == 0 and own[0, 64] == 1 and own[64, 0] == (30 % 8)
    with pytest.raises(ValueError):
        mg.tile_owner_map(100, 100, 2, tile_w=48)
    with pytest.raises(ValueError):
        mg.frames_of_rank(4, 4, 4)


def test_balanced_strips_partition_the_image():
    """bench.py's region partition for the tile-parallel record: strips cover [0, W) without gaps, bounds are multiples of
    64 pixels (fr_set_region_partition's rule), every rank gets at least one block, the weights are split about evenly"""
    import importlib
    import sys
    from conftest import ROOT
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    rng = np.random.default_rng(2)
    for W in (3840, 1920, 1000):
        nb = (W + 63) // 64
        w = np.zeros(nb)
        w[nb // 5: nb // 2] = rng.uniform(1.0, 5.0, nb // 2 - nb // 5)       # the fluid covers part of the image only
        w += 1e-3
        for world in (1, 2, 3, 4, 8):
            strips = bench.balanced_strips(w, W, world)
            assert len(strips) == world and strips[0][0] == 0 and strips[-1][1] == W
            for (a0, a1), (b0, b1) in zip(strips, strips[1:]):
                assert a1 == b0
            for x0, x1 in strips:
                assert x1 > x0 and x0 % 64 == 0 and (x1 % 64 == 0 or x1 == W)
            if world in (2, 4):
                loads = [w[x0 // 64:(x1 + 63) // 64].sum() for x0, x1 in strips]
                assert max(loads) <= w.sum() / world + w.max() * 1.01        # within one block of the ideal split
