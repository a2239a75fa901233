"""Pixel-level parity at the BASELINE headline sizes: the CUDA path against the reference's own translation units
(oracle/_ref, RayMarcher.cpp:256-344 + Dataset.cpp) on the frame bench.py reports -- C2 (1M particles, 1920x1080) and
C3 (4M particles, 3840x2160), default camera.  Where oracle/_ref is not built the C port (validated against _ref at
these sizes in the container that has /root/reference) stands in, on all host cores.

Bar: grid geometry, occupancy counts, depth image, hit mask, hit positions and normals bit for bit; RGBA +-1 code."""
import importlib

import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu
scenes = importlib.import_module("bachelor-thesis_b200.scenes")
camera = importlib.import_module("bachelor-thesis_b200.camera")

SIZES = {
    "C2": dict(n=1_000_000, W=1920, H=1080, h=0.1, dx=None),
    "C3": dict(n=4_000_000, W=3840, H=2160, h=0.063, dx=0.0315),
}


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def cpu_reference(xyz, h, W, H, cam, depth):
    """(positions, normals, geometry, counts, kind) from the reference's own TUs when they are built, else the C port"""
    s = oracle_lib.Settings()
    if oracle_lib.ref_available():
        ds = oracle_lib.Ref().dataset(xyz, h, 2.0)
        pos, nrm, _ = ds.march(W, H, s, cam["inv_proj_view"], cam["position"], depth)
        geo = (ds.min.copy(), ds.max.copy(), ds.dims.copy())
        counts, flags = ds.grid()
        ds.close()
        return pos, nrm, geo, counts, flags, "reference"
    f = oracle_lib.Oracle().frame(xyz, h, 2.0)
    pos, nrm, *_ = f.march(W, H, s, cam["inv_proj_view"], cam["position"], depth, want_band=False)
    counts, flags = f.grid()
    return pos, nrm, (f.min, f.max, f.dims), counts, flags, "port"


@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_frame_bit_exact_vs_reference(fm, oracle, gpu_ctx_factory, cfg):
    c = SIZES[cfg]
    W, H, h = c["W"], c["H"], c["h"]
    xyz = scenes.dam_break(c["n"], h=h, dx=c["dx"])          # t = 0.6: the frame bench.py renders in every arm
    cam = camera.reference_default_camera()
    ctx = gpu_ctx_factory(W, H)
    ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
    ctx.set_settings(fm.VisualizationSettings())
    ctx.upload_frame(0, xyz, h, 2.0)
    ctx.render(fm.FR_PASS_ALL)
    depth, pos, nrm, rgba = ctx.download()
    cnt = ctx.counters()
    g = ctx.download_frame(0)

    # a11: the depth image against the restated rasteriser (the reference makes it on its GPU)
    f = oracle.frame(xyz, h, 2.0)
    want_depth = f.depth_prepass(W, H, cam["view"], cam["proj"])
    assert np.array_equal(bits(depth), bits(want_depth)), f"{cfg}: {(bits(depth) != bits(want_depth)).sum()} depth pixels differ"

    want_pos, want_nrm, geo, counts, flags, kind = cpu_reference(xyz, h, W, H, cam, want_depth)
    info = g["info"]
    assert np.array_equal(bits(info["min"]), bits(geo[0])) and np.array_equal(bits(info["max"]), bits(geo[1]))
    assert np.array_equal(info["grid_dims"], geo[2])
    assert np.array_equal(g["grid_counts"], counts), f"{cfg}: {(g['grid_counts'] != counts).sum()} cell counts differ ({kind})"
    assert np.array_equal(g["grid_flags"], flags)

    # a10: hit mask, positions, normals -- bit for bit (NaN normals compare as bits too)
    mask_diff = int((pos[..., 3] != want_pos[..., 3]).sum())
    pos_diff = int((bits(pos) != bits(want_pos)).any(-1).sum())
    nrm_diff = int((bits(nrm) != bits(want_nrm)).any(-1).sum())
    assert (mask_diff, pos_diff, nrm_diff) == (0, 0, 0), f"{cfg} vs {kind}: mask {mask_diff}, positions {pos_diff}, normals {nrm_diff} pixels differ"
    assert cnt["covered_rays"] == int((want_depth != 1.0).sum())
    assert cnt["hit_rays"] == int((want_pos[..., 3] == 1.0).sum())
    assert cnt["neighbour_overflow"] == 0

    # a12: colour, +-1 code (powf of the sRGB transfer curve is the only non-IEEE operation)
    _, want_rgba = oracle.shade(W, H, want_pos, want_nrm, cam["inv_proj_view"], cam["position"],
                                cam["system"].reshape(3, 3)[2], want_color=False)
    d = np.abs(rgba.astype(np.int16) - want_rgba.astype(np.int16))
    assert d.max() <= 1 and (d > 0).mean() < 0.01
