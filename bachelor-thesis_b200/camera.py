"""Host mirror of the reference camera (src/engine/camera/Camera3D.{h,cpp},
CameraController3D.{h,cpp}) in numpy float32.

The ray marcher takes the camera as INPUT (RayMarcher::Prepare reads camera.Position and
camera.Camera.GetInvProjectionView(), RayMarcher.cpp:95-96), so these matrices are part of the input
contract, not of the kernels.  The formulas below follow glm 0.9.9.8 with the reference's
GLM_FORCE_LEFT_HANDED / DEPTH_ZERO_TO_ONE / RADIANS defines (src/engine/hzpch.h:22-24); the inverse is
taken in float64 and rounded, which agrees with glm::inverse to ~1e-7 relative
(tests/test_camera.py checks it against matrices dumped from the reference's own camera code,
tests/golden/camera_*.json).  Matrices are stored column-major like glm: M[col, row]; use
``.reshape(16)`` for the C ABI.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def perspective_lh_zo(fovy, aspect, z_near, z_far):
    """glm::perspectiveLH_ZO (vendor/glm/glm/ext/matrix_clip_space.inl:265-278)."""
    fovy, aspect, z_near, z_far = F(fovy), F(aspect), F(z_near), F(z_far)
    t = F(math.tan(float(fovy / F(2))))
    m = np.zeros((4, 4), F)
    m[0, 0] = F(1) / (aspect * t)
    m[1, 1] = F(1) / t
    m[2, 2] = z_far / (z_far - z_near)
    m[2, 3] = F(1)
    m[3, 2] = -(z_far * z_near) / (z_far - z_near)
    return m


def _quat_from_euler(pitch, yaw, roll):
    """glm::quat(vec3 eulerAngles) (vendor/glm/glm/detail/type_quat.inl): (w, x, y, z)."""
    c = np.cos(np.array([pitch, yaw, roll], np.float64) * 0.5)
    s = np.sin(np.array([pitch, yaw, roll], np.float64) * 0.5)
    w = c[0] * c[1] * c[2] + s[0] * s[1] * s[2]
    x = s[0] * c[1] * c[2] - c[0] * s[1] * s[2]
    y = c[0] * s[1] * c[2] + s[0] * c[1] * s[2]
    z = c[0] * c[1] * s[2] - s[0] * s[1] * c[2]
    return np.array([w, x, y, z], np.float64)


def _quat_rotate(q, v):
    w, u = q[0], q[1:]
    uv = np.cross(u, v)
    uuv = np.cross(u, uv)
    return v + ((uv * w) + uuv) * 2.0


def _quat_to_mat3(q):
    w, x, y, z = q
    m = np.empty((3, 3), np.float64)   # m[col, row]
    m[0] = [1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y)]
    m[1] = [2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x)]
    m[2] = [2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)]
    return m


def _mat3_to_quat(m):
    """glm::quat_cast(mat3), m[col, row]."""
    fx = m[0, 0] - m[1, 1] - m[2, 2]
    fy = m[1, 1] - m[0, 0] - m[2, 2]
    fz = m[2, 2] - m[0, 0] - m[1, 1]
    fw = m[0, 0] + m[1, 1] + m[2, 2]
    idx, big = 0, fw
    for i, f in enumerate((fx, fy, fz), start=1):
        if f > big:
            idx, big = i, f
    bv = math.sqrt(big + 1.0) * 0.5
    mult = 0.25 / bv
    if idx == 0:
        return np.array([bv, (m[1, 2] - m[2, 1]) * mult, (m[2, 0] - m[0, 2]) * mult, (m[0, 1] - m[1, 0]) * mult])
    if idx == 1:
        return np.array([(m[1, 2] - m[2, 1]) * mult, bv, (m[0, 1] + m[1, 0]) * mult, (m[2, 0] + m[0, 2]) * mult])
    if idx == 2:
        return np.array([(m[2, 0] - m[0, 2]) * mult, (m[0, 1] + m[1, 0]) * mult, bv, (m[1, 2] + m[2, 1]) * mult])
    return np.array([(m[0, 1] - m[1, 0]) * mult, (m[2, 0] + m[0, 2]) * mult, (m[1, 2] + m[2, 1]) * mult, bv])


def _quat_look_at_lh(direction, up):
    """glm::quatLookAtLH (vendor/glm/glm/gtc/quaternion.inl)."""
    m = np.empty((3, 3), np.float64)
    m[2] = direction
    right = np.cross(up, m[2])
    m[0] = right / math.sqrt(max(1e-5, float(np.dot(right, right))))
    m[1] = np.cross(m[2], m[0])
    return _mat3_to_quat(m)


class Camera3D:
    """Camera3D (src/engine/camera/Camera3D.cpp:5-20)."""

    def __init__(self, fov=math.radians(60.0), aspect=16.0 / 9.0, near=0.1, far=1000.0):
        self.FOV, self.Aspect, self.Near, self.Far = fov, aspect, near, far
        p = perspective_lh_zo(fov, aspect, near, far)
        p[1, :] = -p[1, :]                      # glm::scale(P, {1,-1,1}) scales column 1
        self.Projection = p
        self.InvProjection = np.linalg.inv(p.astype(np.float64).T).T.astype(F)
        self.View = np.eye(4, dtype=F)
        self.ComputeMatrices()

    def ComputeMatrices(self):
        v = self.View.astype(np.float64).T       # to row-major maths
        p = self.Projection.astype(np.float64).T
        pv = (p.astype(F) @ v.astype(F)).astype(F)
        self.InvView = np.linalg.inv(v).T.astype(F)
        self.ProjectionView = pv.T.copy()
        self.InvProjectionView = np.linalg.inv(pv.astype(np.float64)).T.astype(F)

    def GetInvProjectionView(self):
        return self.InvProjectionView

    def GetProjection(self):
        return self.Projection

    def GetView(self):
        return self.View


class CameraController3D:
    """Orbit controller (src/engine/camera/CameraController3D.cpp:70-84): Position = rotate(quat(RotationY,
    RotationX, 0), (0, 0, -R)); View = translate(mat4(inverse(quatLookAt(-normalize(Position), up))), -Position);
    System = transpose(mat3(View))."""

    def __init__(self, camera: Camera3D, R=10.0, RotationX=0.0, RotationY=0.0):
        self.Camera = camera
        self.R, self.RotationX, self.RotationY = R, RotationX, RotationY
        self.ComputeMatrices()

    def ComputeMatrices(self):
        pi = math.pi
        self.RotationY = min(max(self.RotationY, -0.49 * pi), 0.49 * pi)
        q = _quat_from_euler(self.RotationY, self.RotationX, 0.0)
        pos = _quat_rotate(q, np.array([0.0, 0.0, -self.R]))
        self.Position = pos.astype(F)
        d = -pos / math.sqrt(float(np.dot(pos, pos)))
        orient = _quat_look_at_lh(d, np.array([0.0, 1.0, 0.0]))
        inv = np.array([orient[0], -orient[1], -orient[2], -orient[3]]) / float(np.dot(orient, orient))
        rot = _quat_to_mat3(inv)                # m[col, row]
        view = np.eye(4, dtype=np.float64)      # [col, row]
        view[:3, :3] = rot
        # glm::translate(m, v): m[3] = m[0]*v0 + m[1]*v1 + m[2]*v2 + m[3]
        t = -pos
        view[3, :] = view[0, :] * t[0] + view[1, :] * t[1] + view[2, :] * t[2] + view[3, :]
        self.Camera.View = view.astype(F)
        self.System = self.Camera.View[:3, :3].T.copy()   # transpose(mat3(View)), [col, row]
        self.Camera.ComputeMatrices()

    @property
    def Direction(self):
        """System[2], what CompositionRenderPass uploads as CameraDirection (:319)."""
        return self.System[2].copy()


def reference_default_camera(name: str = "camera_default_16x9") -> dict:
    """The reference's default camera (orbit R = 10, fov 60 deg, 16:9, near 0.1, far 1000) as float32 arrays with the
    exact bits the reference's own Camera3D / CameraController3D produce (dumped by tests/golden/make_golden.py from
    the reference TUs; the numpy mirror above agrees to ~1e-7 relative, this table is what parity runs use so that the
    CUDA path and the CPU reference see the same input bits).  Keys: view, proj, inv_proj, inv_proj_view, position,
    system."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", name + ".json")) as f:
        d = json.load(f)
    return {k: np.array(v, dtype=np.float32) for k, v in d["float32"].items()}
