"""Host camera mirror vs matrices dumped from the reference's own Camera3D / CameraController3D."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.mark.parametrize("name", ["camera_default_16x9", "camera_orbit_a_16x9", "camera_orbit_b_16x9", "camera_close_16x9"])
def test_camera_mirror_matches_reference(fm, name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        g = json.load(f)
    p = g["params"]
    cam = fm.Camera3D(p["fov"], p["aspect"], p["near"], p["far"])
    ctl = fm.CameraController3D(cam, p["R"], p["rot_x"], p["rot_y"])
    want = {k: np.array(v, np.float32) for k, v in g["float32"].items()}
    for key, mine in (("view", cam.View), ("proj", cam.Projection), ("inv_proj_view", cam.InvProjectionView),
                      ("position", ctl.Position), ("system", ctl.System)):
        a, b = want[key], np.asarray(mine, np.float32).reshape(-1)
        tol = 1e-4 if key == "inv_proj_view" else 1e-5   # glm::inverse runs in float32 on an ill-conditioned matrix
        assert np.abs(a - b).max() <= tol * max(1.0, np.abs(a).max()), key


def test_default_camera_contract():
    """Position (0,0,-10), System = identity, P[1][1] < 0 (y flipped), LH, depth 0..1"""
    with open(os.path.join(GOLDEN, "camera_default_16x9.json")) as f:
        g = json.load(f)["float32"]
    assert g["position"] == [0.0, 0.0, -10.0]
    assert np.allclose(np.array(g["system"]).reshape(3, 3), np.eye(3))
    proj = np.array(g["proj"], np.float32)
    assert proj[5] < 0 and proj[11] == 1.0 and abs(proj[10] - 1.0001) < 1e-6


def test_package_camera_table_is_the_golden_dump(fm):
    """bachelor-thesis_b200/data holds a copy of the matrices dumped from the reference's camera TUs (bench.py and
    smoke() read it, so that neither depends on tests/)"""
    from conftest import golden_camera
    for name in ("camera_default_16x9", "camera_close_16x9"):
        a, b = fm.camera.reference_default_camera(name), golden_camera(name)
        assert set(a) == set(b)
        for k in a:
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), (name, k)
