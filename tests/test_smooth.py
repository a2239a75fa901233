"""the Gauss pass over the depth image (GaussRenderPass.cpp:15-66, gauss.frag:28-47 -- it runs in the reference) and
the Sobel screen normals of the unprojected smoothed depth (composition.frag:50-57,87-104 -- `#if 0` there).  CPU tests pin the
oracle's restatement by construction properties; the GPU tests compare the CUDA pass with it bit for bit."""
import importlib
import math

import numpy as np
import pytest

from conftest import golden_camera

scenes = importlib.import_module("bachelor-thesis_b200.scenes")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("n", [0, 1, 8, 31])
def test_oracle_gauss_kernel(oracle, n):
    k = oracle.gauss_kernel(n).astype(np.float64)
    full = np.array([[k[abs(i), abs(j)] for j in range(-n, n + 1)] for i in range(-n, n + 1)])
    assert abs(full.sum() - 1.0) < 1e-5                     # normalised over the (2N+1)^2 taps
    assert np.allclose(k, k.T, rtol=1e-6)                   # e^(-r^2/2) is symmetric in (i, j)
    want = np.array([[math.exp(-0.5 * (i * i + j * j)) for j in range(n + 1)] for i in range(n + 1)])
    assert np.allclose(k / k[0, 0], want, rtol=1e-5, atol=1e-30)


def test_oracle_gauss_depth_properties(oracle):
    rng = np.random.default_rng(1)
    d = rng.uniform(0.9, 1.0, (40, 56)).astype(np.float32)
    assert np.abs(oracle.gauss_depth(np.full((40, 56), 0.75, np.float32), 8) - 0.75).max() < 2e-6      # weights sum to 1, edges clamp
    assert np.array_equal(bits(oracle.gauss_depth(d, 0)), bits(d))                                      # N = 0: the single tap has weight 1
    s = oracle.gauss_depth(d, 3)
    assert s.min() >= d.min() - 1e-6 and s.max() <= d.max() + 1e-6                                      # a convex combination
    # separable reference in float64 (the Gaussian factorises), interior pixels
    k = oracle.gauss_kernel(3).astype(np.float64)
    full = np.array([[k[abs(i), abs(j)] for j in range(-3, 4)] for i in range(-3, 4)])
    y, x = 20, 30
    want = sum(full[i + 3, j + 3] * float(d[y + j, x + i]) for i in range(-3, 4) for j in range(-3, 4))
    assert abs(float(s[y, x]) - want) < 1e-6


def test_oracle_sobel_normal_of_a_plane(oracle, default_camera):
    """a constant depth image unprojects to a plane z = const in view space: the Sobel normal is +-z everywhere"""
    W, H = 64, 36
    sm = np.full((H, W), 0.99, np.float32)
    n = oracle.sobel_normals(sm, default_camera["inv_proj"])
    inner = n[2:-2, 2:-2]
    assert np.all(np.abs(np.abs(inner[..., 2]) - 1.0) < 1e-4) and np.all(np.abs(inner[..., :2]) < 1e-2)
    assert np.all(inner[..., 3] == 1.0)


def test_library_gauss_kernel_matches_oracle(fm, oracle):
    for n in (0, 3, 8, 31):
        assert np.array_equal(bits(fm.gauss_kernel(n)), bits(oracle.gauss_kernel(n)))      # same libm powf on both sides
    with pytest.raises(fm.FluidMarchError):
        fm.gauss_kernel(32)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 2, 8])
def test_gpu_smoothing_matches_oracle_bit_for_bit(fm, oracle, gpu_ctx_factory, n):
    W, H = 320, 180
    cam = golden_camera("camera_close_16x9")
    ctx = gpu_ctx_factory(W, H)
    ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
    ctx.upload_frame(0, scenes.dam_break(8000), 0.1, 2.0)
    ctx.render(fm.FR_PASS_DEPTH)
    depth = ctx.download(True, False, False, False)[0]
    sm, nrm = ctx.smooth_depth(n, cam["inv_proj"])
    want_sm = oracle.gauss_depth(depth, n)
    assert np.array_equal(bits(sm), bits(want_sm))
    want_n = oracle.sobel_normals(want_sm, cam["inv_proj"])
    assert np.array_equal(bits(nrm), bits(want_n))                # NaN normals (flat regions give cross = 0) compare as bits too
    sm2, none = ctx.smooth_depth(n)
    assert none is None and np.array_equal(bits(sm2), bits(sm))
