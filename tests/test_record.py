"""Recording: fr_encode_bmp / fr_write_bmp against the file stb_image_write (the reference's vendored copy, called as in
Renderer::_Screenshot, Renderer.cpp:400-409) produces for the same pixels.  tests/golden/bmp_16x9_stb.bmp was written
by that stb through oracle/_ref (see make_golden_bgeo.py's sibling snippet in DESIGN.md section 5)."""
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN, golden_camera


def bmp_bytes(rgba):
    """restatement of stbi_write_bmp_core for comp == 4 (stb_image_write.h:501-509)"""
    h, w = rgba.shape[:2]
    hdr = (b"BM" + struct.pack("<IHHI", 122 + w * h * 4, 0, 0, 122)
           + struct.pack("<IiiHHIIiiII", 108, w, h, 1, 32, 3, 0, 0, 0, 0, 0)
           + struct.pack("<IIII", 0xff0000, 0xff00, 0xff, 0xff000000) + struct.pack("<I", 0) + b"\0" * 48)
    return hdr + np.ascontiguousarray(rgba[::-1, :, [2, 1, 0, 3]]).tobytes()


def test_restatement_matches_the_stb_golden_file():
    img = np.load(os.path.join(GOLDEN, "bmp_16x9_rgba.npy"))
    assert bmp_bytes(img) == open(os.path.join(GOLDEN, "bmp_16x9_stb.bmp"), "rb").read()


def test_restatement_matches_live_stb(ref, tmp_path):
    if not ref.has_bmp:
        pytest.skip("oracle/_ref predates the stb hook")
    rng = np.random.default_rng(11)
    for (h, w) in ((1, 1), (3, 5), (37, 64), (180, 321)):
        img = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        p = str(tmp_path / f"{w}x{h}.bmp")
        assert ref.write_bmp(p, img) == 1
        assert open(p, "rb").read() == bmp_bytes(img)


@pytest.mark.gpu
def test_encode_bmp_matches_the_stb_golden_file(fm, tmp_path):
    import torch
    img = np.load(os.path.join(GOLDEN, "bmp_16x9_rgba.npy"))
    ctx = fm.Context(16, 9)
    try:
        t = torch.from_numpy(img).cuda()
        ctx.set_color_target(t.data_ptr())                      # the colour image the recording reads
        want = open(os.path.join(GOLDEN, "bmp_16x9_stb.bmp"), "rb").read()
        assert ctx.encode_bmp() == want
        p = str(tmp_path / "shot.bmp")
        ctx.write_bmp(p)
        assert open(p, "rb").read() == want
    finally:
        ctx.close()


@pytest.mark.gpu
def test_recorded_sequence_frames(fm, tmp_path):
    """g_Recording: every frame of a sequence is saved as screenshot_<k>.bmp (AdvancedRenderer.cpp:283-297)"""
    cam = golden_camera("camera_close_16x9")
    W, H = 321, 180                                             # odd width: no row padding games in 32 bpp
    seq = fm.Sequence(W, H, lanes=2)
    try:
        seq.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
        seq.set_settings(fm.VisualizationSettings())
        frames = [fm.scenes.dam_break(5000 + 500 * i, t=0.3 + 0.1 * i) for i in range(4)]
        outs = []
        for k, xyz in enumerate(frames):
            rgba = np.zeros((H, W, 4), np.uint8)
            seq.submit_ptrs(xyz.ctypes.data, len(xyz), 0.1, 2.0, rgba=rgba.ctypes.data, bmp_path=str(tmp_path / f"screenshot_{k}.bmp"))
            outs.append(rgba)
        seq.drain()
        for k, rgba in enumerate(outs):
            assert rgba.any()
            assert open(tmp_path / f"screenshot_{k}.bmp", "rb").read() == bmp_bytes(rgba)
    finally:
        seq.close()
