"""An INDEPENDENT restatement of the reference's depth pre-pass by rasteriser rules (test infrastructure).

oracle/fluid_oracle.c::fo_depth_prepass -- the checker the CUDA pre-pass is bit-identical to -- evaluates the pass
analytically: the impostor quad lies in the plane view-z = z_c, so u and v at a pixel centre follow from one ray / plane
intersection.  Nothing the reference produced pins that formulation (its depth image is made by a Vulkan rasteriser,
absent here).  This module restates the pass the way a rasteriser executes it instead, to bound how far a real D32
pipeline can deviate from the analytic one:

  * CollectRenderData (src/app/AdvancedRenderer/AdvancedRenderer.cpp:447-485): per particle two triangles
    (TL, BL, TR), (TR, BL, BR) with corners p + R(-+x +-y), x / y = CameraController.System[0 / 1], UV in {-1, 1}^2, FP32
  * depth.vert:20-27: viewPosition = View * vec4(a_Position, 1); ViewPosition = xyz / w; gl_Position = Projection * vp, FP32
  * rasterisation (Vulkan spec 1.3, 27.7): viewport (0, 0, W, H); vertices optionally snapped to the 1/256 pixel grid
    (VkPhysicalDeviceLimits::subPixelPrecisionBits = 8 on NVIDIA); a pixel is covered when its CENTRE is inside the
    triangle, edges owned by the top-left rule; attributes interpolated perspective-correctly with barycentric
    weights (float64 here: the hardware's interpolation precision is not specified)
  * depth.frag:19-33 per covered sample in FP32: l2 = dot(UV, UV); discard if l2 > 1; off = cos(pi/2 sqrt(l2));
    z = (P * (ViewPosition - R (0, 0, off))).z / w; DepthRenderPass.cpp:54,155,177-179: clear 1.0, cull none, test Less
"""
from __future__ import annotations

import numpy as np

F = np.float32


def _mat_vec(m, x, y, z, w):
    """glm mat4 * vec4, column-major m[c*4+r]: (m0 v0 + m1 v1) + (m2 v2 + m3 v3), FP32 (type_mat4x4.inl:561-572)"""
    out = []
    for r in range(4):
        a = (F(m[0 + r]) * x + F(m[4 + r]) * y).astype(F)
        b = (F(m[8 + r]) * z + F(m[12 + r]) * w).astype(F)
        out.append((a + b).astype(F))
    return out


def rasterise_depth(xyz, R, W, H, view, proj, system, snap_bits=8, chunk=4096, collect_band=None, band_tol=1e-4):
    """depth image (H, W) float32 by rasteriser rules; `collect_band`: if a list, receives the pixel indices of every
    covered sample with |l2 - 1| < band_tol (the disc-edge band where discard decisions may legitimately differ)"""
    xyz = np.ascontiguousarray(xyz, F).reshape(-1, 3)
    view = np.asarray(view, F).reshape(16)
    proj = np.asarray(proj, F).reshape(16)
    sx = np.asarray(system, F).reshape(3, 3)[0]
    sy = np.asarray(system, F).reshape(3, 3)[1]
    R = F(R)
    depth = np.ones(W * H, F)
    one = F(1.0)
    # corner offsets R * (sgn_x * x + sgn_y * y), FP32, in the reference's expression order: R * (-x + y) etc.
    corners = {}
    for name, (cx, cy) in dict(TL=(-1, 1), TR=(1, 1), BL=(-1, -1), BR=(1, -1)).items():
        corners[name] = (R * ((F(cx) * sx).astype(F) + (F(cy) * sy).astype(F)).astype(F)).astype(F)
    uv = dict(TL=(-1.0, -1.0), BL=(-1.0, 1.0), TR=(1.0, -1.0), BR=(1.0, 1.0))
    tris = (("TL", "BL", "TR"), ("TR", "BL", "BR"))
    snap = float(1 << snap_bits) if snap_bits else None
    for c0 in range(0, len(xyz), chunk):
        P = xyz[c0:c0 + chunk]
        n = len(P)
        vert = {}
        for name, off in corners.items():
            wx, wy, wz = (P[:, 0] + off[0]).astype(F), (P[:, 1] + off[1]).astype(F), (P[:, 2] + off[2]).astype(F)
            v = _mat_vec(view, wx, wy, wz, np.full(n, one))
            vp = [(v[k] / v[3]).astype(F) for k in range(3)]                    # ViewPosition
            clip = _mat_vec(proj, v[0], v[1], v[2], v[3])
            w = clip[3].astype(np.float64)
            px = (clip[0].astype(np.float64) / w + 1.0) * (W / 2.0)             # viewport transform
            py = (clip[1].astype(np.float64) / w + 1.0) * (H / 2.0)
            if snap:
                px, py = np.round(px * snap) / snap, np.round(py * snap) / snap
            vert[name] = dict(px=px, py=py, w=w, vp=vp, ok=(clip[3] > 0) & (clip[2] >= 0) & (clip[2] <= clip[3]))
        # quads clipped by near / far are dropped whole, as fo_depth_prepass does (same test on the centre depth)
        x_lo = np.floor(np.minimum.reduce([vert[k]["px"] for k in vert]) - 0.5).astype(np.int64)
        x_hi = np.ceil(np.maximum.reduce([vert[k]["px"] for k in vert]) - 0.5).astype(np.int64)
        y_lo = np.floor(np.minimum.reduce([vert[k]["py"] for k in vert]) - 0.5).astype(np.int64)
        y_hi = np.ceil(np.maximum.reduce([vert[k]["py"] for k in vert]) - 0.5).astype(np.int64)
        span = int(max((x_hi - x_lo).max(), (y_hi - y_lo).max())) + 1
        ok = np.logical_and.reduce([vert[k]["ok"] for k in vert])
        gx = x_lo[:, None, None] + np.arange(span)[None, None, :]
        gy = y_lo[:, None, None] + np.arange(span)[None, :, None]
        gx, gy = np.broadcast_arrays(gx, gy)
        cxs, cys = gx + 0.5, gy + 0.5                                            # pixel centres
        inside_screen = (gx >= 0) & (gx < W) & (gy >= 0) & (gy < H) & ok[:, None, None]
        for tri in tris:
            a, b, c = (vert[k] for k in tri)
            ax, ay, bx, by, cx_, cy_ = (t[:, None, None] for t in (a["px"], a["py"], b["px"], b["py"], c["px"], c["py"]))
            area = (bx - ax) * (cy_ - ay) - (by - ay) * (cx_ - ax)
            sgn = np.where(area < 0, -1.0, 1.0)                                  # cull none: both windings

            def edge(x0, y0, x1, y1):
                e = ((x1 - x0) * (cys - y0) - (y1 - y0) * (cxs - x0)) * sgn
                dx, dy = (x1 - x0) * sgn, (y1 - y0) * sgn
                # top-left rule (y down): a top edge is horizontal with the interior below it, a left edge goes "up"
                top_left = ((dy == 0) & (dx > 0)) | (dy < 0)
                return e, (e > 0) | ((e == 0) & top_left)
            e0, in0 = edge(bx, by, cx_, cy_)
            e1, in1 = edge(cx_, cy_, ax, ay)
            e2, in2 = edge(ax, ay, bx, by)
            cov = in0 & in1 & in2 & inside_screen & (np.abs(area) > 0)
            if not cov.any():
                continue
            idx = np.nonzero(cov)
            pi = idx[0]
            tot = (e0 + e1 + e2)[idx]
            l0, l1, l2b = e0[idx] / tot, e1[idx] / tot, e2[idx] / tot          # screen-space barycentrics
            iw = l0 / a["w"][pi] + l1 / b["w"][pi] + l2b / c["w"][pi]           # perspective correction
            b0, b1, b2 = l0 / a["w"][pi] / iw, l1 / b["w"][pi] / iw, l2b / c["w"][pi] / iw
            u = (b0 * uv[tri[0]][0] + b1 * uv[tri[1]][0] + b2 * uv[tri[2]][0]).astype(F)
            v = (b0 * uv[tri[0]][1] + b1 * uv[tri[1]][1] + b2 * uv[tri[2]][1]).astype(F)
            vz = (b0 * a["vp"][2][pi] + b1 * b["vp"][2][pi] + b2 * c["vp"][2][pi]).astype(F)
            l2 = ((u * u).astype(F) + (v * v).astype(F)).astype(F)
            pix = (gy[idx] * W + gx[idx]).astype(np.int64)
            if collect_band is not None:
                near = np.abs(l2.astype(np.float64) - 1.0) < band_tol
                collect_band.append(pix[near])
            keep = ~(l2 > one)
            off = np.cos(np.float64(1.57079632679) * np.sqrt(l2.astype(np.float64))).astype(F)
            zf = (vz - (R * off).astype(F)).astype(F)
            zc = ((F(proj[10]) * zf).astype(F) + F(proj[14])).astype(F)          # (P * vec4(pView, 1)).z ; w = pView.z (P[2][3] = 1)
            d = (zc / zf).astype(F)
            d = np.clip(d, F(0.0), F(1.0))
            keep &= d < one
            np.minimum.at(depth, pix[keep], d[keep])
    return depth.reshape(H, W)
