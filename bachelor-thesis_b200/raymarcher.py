"""Host-side mirror of the reference's operator interface for the ray-march path, over the C ABI.

Names, argument meaning and call protocol follow the reference:
  * ``VisualizationSettings``  src/app/AdvancedRenderer/RayMarcher.h:12-26 (defaults AdvancedRenderer.cpp:18-28)
  * ``Dataset`` / frames       src/app/Dataset.h:67-105  (particleRadius / particleRadiusMultiplier: config.yml:19-20)
  * ``RayMarcher``             src/app/AdvancedRenderer/RayMarcher.h:28-44: Prepare -> Start (returns
                               immediately) -> poll IsDone -> read positions/normals; Exit at shutdown.
``Context`` is the thin object wrapper of the fr_* entry points the classes above are built on.
All compute happens in libfluidmarch.so on the GPU; nothing here has a CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _cabi as abi
from ._cabi import FR_PASS_ALL, FR_PASS_DEPTH, FR_PASS_MARCH, FR_PASS_SHADE, check


@dataclass
class VisualizationSettings:
    Frame: int = 0
    MaxSteps: int = 128
    StepSize: float = 0.009
    IsoDensity: float = 1.0
    EnableAnisotropy: bool = False     # the reference's default is True (AdvancedRenderer.cpp:23); north_star's headline path is the isotropic one
    k_n: float = 0.5
    k_r: float = 2.0
    k_s: float = 2000.0
    N_eps: int = 1
    # additions, 0 = reference behaviour
    BisectionSteps: int = 0
    SkipLastPixel: bool = False
    FastNormals: bool = False

    def to_c(self) -> abi.FrSettings:
        return abi.FrSettings(self.Frame, self.MaxSteps, self.StepSize, self.IsoDensity,
                              1 if self.EnableAnisotropy else 0, self.k_n, self.k_r, self.k_s, self.N_eps,
                              self.BisectionSteps, 1 if self.SkipLastPixel else 0, 1 if self.FastNormals else 0)


def _ptr(a, ctype=C.c_void_p):
    return None if a is None else a.ctypes.data_as(ctype)


def _camera_struct(view, projection, inv_projection_view, position, direction) -> abi.FrCamera:
    cam = abi.FrCamera()
    for name, src, n in (("view", view, 16), ("projection", projection, 16),
                         ("inv_projection_view", inv_projection_view, 16),
                         ("position", position, 3), ("direction", direction, 3)):
        arr = np.ascontiguousarray(src, dtype=np.float32).reshape(-1)
        if arr.size != n:
            raise ValueError(f"camera.{name}: expected {n} floats, got {arr.size}")
        getattr(cam, name)[:] = arr.tolist()
    return cam


class Context:
    """One GPU context (fr_create .. fr_destroy)."""

    def __init__(self, width: int, height: int, device: int = 0):
        self.lib = abi.load()
        self.width, self.height = int(width), int(height)
        h = C.c_void_p()
        check(self.lib.fr_create(device, self.width, self.height, C.byref(h)), "fr_create")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.fr_destroy(self.h)
            self.h = None

    __del__ = close

    def resize(self, width, height):
        check(self.lib.fr_resize(self.h, width, height), "fr_resize")
        self.width, self.height = int(width), int(height)

    # frames ---------------------------------------------------------------------------------------
    def upload_frame(self, frame: int, xyz, h: float = 0.1, h_ext_mult: float = 2.0):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        check(self.lib.fr_upload_frame(self.h, frame, _ptr(xyz), xyz.shape[0], h, h_ext_mult), "fr_upload_frame")

    def upload_frame_ptr(self, frame: int, host_ptr: int, n: int, h: float = 0.1, h_ext_mult: float = 2.0):
        check(self.lib.fr_upload_frame(self.h, frame, C.c_void_p(host_ptr), n, h, h_ext_mult), "fr_upload_frame")

    def upload_frame_bgeo(self, frame: int, path: str, h: float = 0.1, h_ext_mult: float = 2.0):
        """the frame straight from a classic .bgeo file (decoded on the GPU)"""
        check(self.lib.fr_upload_frame_bgeo(self.h, frame, os.fsencode(path), h, h_ext_mult), "fr_upload_frame_bgeo")

    def build_frame_device(self, frame: int, dev_ptr: int, n: int, h: float = 0.1, h_ext_mult: float = 2.0):
        check(self.lib.fr_build_frame_device(self.h, frame, C.c_void_p(dev_ptr), n, h, h_ext_mult),
              "fr_build_frame_device")

    def set_count_mode(self, mode: int):
        """reading of the fork-only find_neighbors_box for frames built afterwards: FR_COUNT_CENTRE_BOX (default, what the
        reference build under oracle/_ref computes) or FR_COUNT_CELL_EXACT"""
        check(self.lib.fr_set_count_mode(self.h, mode), "fr_set_count_mode")

    def set_async_build(self, on: bool):
        """frame builds into a slot that already has tables skip the host round trip (default on); errors of such a
        build surface at the next wait"""
        check(self.lib.fr_set_async_build(self.h, 1 if on else 0), "fr_set_async_build")

    def set_stage_timing(self, on: bool):
        """per-stage CUDA events (timings()); off: the frame's kernels chain by programmatic dependent launch"""
        check(self.lib.fr_set_stage_timing(self.h, 1 if on else 0), "fr_set_stage_timing")

    def frame_info(self, frame: int) -> dict:
        fi = abi.FrFrameInfo()
        check(self.lib.fr_get_frame_info(self.h, frame, C.byref(fi)), "fr_get_frame_info")
        return dict(num_particles=int(fi.num_particles), h=float(fi.h),
                    min=np.array(fi.min[:], np.float32), max=np.array(fi.max[:], np.float32),
                    grid_dims=np.array(fi.grid_dims[:], np.int32), occupied_cells=int(fi.occupied_cells),
                    search_min=np.array(fi.search_min[:], np.int32), search_dims=np.array(fi.search_dims[:], np.int32))

    def release_frame(self, frame: int):
        check(self.lib.fr_release_frame(self.h, frame), "fr_release_frame")

    def download_frame(self, frame: int):
        fi = self.frame_info(frame)
        n = fi["num_particles"]
        cells = int(np.prod(fi["search_dims"].astype(np.int64)))
        gcells = int(np.prod(fi["grid_dims"].astype(np.int64)))
        sorted_xyzi = np.zeros((n, 4), np.float32)
        cell_start = np.zeros(cells + 1, np.uint32)
        counts = np.zeros(gcells, np.uint32)
        flags = np.zeros(gcells, np.uint8)
        check(self.lib.fr_download_frame(self.h, frame, _ptr(sorted_xyzi, abi.f32p), _ptr(cell_start, abi.u32p),
                                         _ptr(counts, abi.u32p), _ptr(flags, abi.u8p)), "fr_download_frame")
        return dict(sorted_xyz=sorted_xyzi[:, :3].copy(), sorted_index=sorted_xyzi[:, 3].copy().view(np.uint32),
                    cell_start=cell_start, grid_counts=counts, grid_flags=flags, info=fi)

    # per-render state -----------------------------------------------------------------------------
    def set_settings(self, s: VisualizationSettings):
        cs = s.to_c()
        check(self.lib.fr_set_settings(self.h, C.byref(cs)), "fr_set_settings")

    def set_camera(self, view, projection, inv_projection_view, position, direction):
        cam = _camera_struct(view, projection, inv_projection_view, position, direction)
        check(self.lib.fr_set_camera(self.h, C.byref(cam)), "fr_set_camera")

    def set_camera_controller(self, controller):
        """controller: a CameraController3D-like object (Position, System, Camera.{View,Projection,InvProjectionView})."""
        cam = controller.Camera
        self.set_camera(cam.View, cam.Projection, cam.InvProjectionView, controller.Position,
                        np.asarray(controller.System).reshape(3, 3)[2])

    def set_depth(self, depth):
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        if depth.size != self.width * self.height:
            raise ValueError("depth must hold W*H floats")
        check(self.lib.fr_set_depth(self.h, _ptr(depth)), "fr_set_depth")

    def set_tile_partition(self, rank, world, tile_w=64, tile_h=64):
        check(self.lib.fr_set_tile_partition(self.h, rank, world, tile_w, tile_h), "fr_set_tile_partition")

    def set_region_partition(self, x0=0, y0=0, x1=0, y1=0):
        """render (and build frames for) the pixel rectangle [x0, x1) x [y0, y1) only; all zero = off"""
        check(self.lib.fr_set_region_partition(self.h, x0, y0, x1, y1), "fr_set_region_partition")

    # render ----------------------------------------------------------------------------------------
    def render_async(self, passes: int = FR_PASS_ALL):
        check(self.lib.fr_render_async(self.h, passes), "fr_render_async")

    def is_done(self) -> bool:
        rc = self.lib.fr_is_done(self.h)
        if rc < 0:
            check(rc, "fr_is_done")
        return rc == 1

    def wait(self):
        check(self.lib.fr_wait(self.h), "fr_wait")

    def render(self, passes: int = FR_PASS_ALL):
        self.render_async(passes)
        self.wait()

    # results ---------------------------------------------------------------------------------------
    def download(self, depth=True, positions=True, normals=True, rgba=True):
        H, W = self.height, self.width
        d = np.empty((H, W), np.float32) if depth else None
        p = np.empty((H, W, 4), np.float32) if positions else None
        n = np.empty((H, W, 4), np.float32) if normals else None
        c = np.empty((H, W, 4), np.uint8) if rgba else None
        check(self.lib.fr_download(self.h, _ptr(d), _ptr(p), _ptr(n), _ptr(c)), "fr_download")
        return d, p, n, c

    def download_into(self, depth=None, positions=None, normals=None, rgba=None):
        check(self.lib.fr_download(self.h, _ptr(depth), _ptr(positions), _ptr(normals), _ptr(rgba)), "fr_download")

    def download_ptrs(self, depth=0, positions=0, normals=0, rgba=0):
        vp = lambda a: C.c_void_p(a) if a else None
        check(self.lib.fr_download(self.h, vp(depth), vp(positions), vp(normals), vp(rgba)), "fr_download")

    def device_images(self) -> dict:
        d, p, n, c = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self.lib.fr_device_images(self.h, C.byref(d), C.byref(p), C.byref(n), C.byref(c)), "fr_device_images")
        return dict(depth=d.value, positions=p.value, normals=n.value, rgba=c.value)

    def encode_bmp(self) -> bytes:
        """the colour image of the last render as the reference's screenshot file (Renderer::_Screenshot)"""
        n = C.c_size_t()
        check(self.lib.fr_encode_bmp(self.h, None, 0, C.byref(n)), "fr_encode_bmp")
        buf = np.zeros(n.value, np.uint8)
        check(self.lib.fr_encode_bmp(self.h, buf.ctypes.data, n.value, C.byref(n)), "fr_encode_bmp")
        return buf.tobytes()

    def write_bmp(self, path: str):
        check(self.lib.fr_write_bmp(self.h, os.fsencode(path)), "fr_write_bmp")

    def set_color_target(self, dev_ptr: int | None):
        check(self.lib.fr_set_color_target(self.h, C.c_void_p(dev_ptr) if dev_ptr else None), "fr_set_color_target")

    def ipc_export_color(self) -> bytes:
        """64-byte handle of this context's colour image for the other processes of the box (fr_ipc_export_color)"""
        buf = (C.c_ubyte * 64)()
        check(self.lib.fr_ipc_export_color(self.h, buf), "fr_ipc_export_color")
        return bytes(buf)

    def ipc_open_color_target(self, handle: bytes):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        check(self.lib.fr_ipc_open_color_target(self.h, buf), "fr_ipc_open_color_target")

    def ipc_close_color_target(self):
        check(self.lib.fr_ipc_close_color_target(self.h), "fr_ipc_close_color_target")

    def counters(self) -> dict:
        c = abi.FrCounters()
        check(self.lib.fr_get_counters(self.h, C.byref(c)), "fr_get_counters")
        return c.as_dict()

    def timings(self) -> dict:
        t = abi.FrTimings()
        check(self.lib.fr_get_timings(self.h, C.byref(t)), "fr_get_timings")
        return t.as_dict()

    def stream(self) -> int:
        s = C.c_void_p()
        check(self.lib.fr_get_stream(self.h, C.byref(s)), "fr_get_stream")
        return s.value or 0

    # point queries ---------------------------------------------------------------------------------
    def query_neighbors(self, frame: int, points, cap: int = 256, ext: bool = False):
        """Dataset::GetNeighbors (ext=False, r = h) / GetNeighborsExt (ext=True, r = h_ext): counts and original ids"""
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        m = pts.shape[0]
        counts = np.zeros(m, np.uint32)
        ids = np.zeros((m, cap), np.uint32) if cap else None
        fn = self.lib.fr_query_neighbors_ext if ext else self.lib.fr_query_neighbors
        check(fn(self.h, frame, _ptr(pts, abi.f32p), m, _ptr(counts, abi.u32p), _ptr(ids, abi.u32p), cap),
              "fr_query_neighbors_ext" if ext else "fr_query_neighbors")
        return counts, ids

    def query_anisotropic(self, frame: int, points, want_grad=True):
        """per point: WPCA's G (9 floats, glm::mat3 column-major), the anisotropic density and gradient sums"""
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        m = pts.shape[0]
        rho = np.zeros(m, np.float32)
        grad = np.zeros((m, 3), np.float32) if want_grad else None
        g9 = np.zeros((m, 9), np.float32)
        check(self.lib.fr_query_anisotropic(self.h, frame, _ptr(pts, abi.f32p), m, _ptr(rho, abi.f32p),
                                            _ptr(grad, abi.f32p), _ptr(g9, abi.f32p)), "fr_query_anisotropic")
        return rho, grad, g9

    def download_frame_ext(self, frame: int):
        """the r = h_ext search (Frame::m_SearchExt): sorted (n, 4) xyz + original id bits, cell_start, kmin, kdim"""
        info = self.frame_info(frame)
        n = int(info["num_particles"])
        kmin = np.zeros(3, np.int32)
        kdim = np.zeros(3, np.int32)
        i32p = C.POINTER(C.c_int32)
        check(self.lib.fr_download_frame_ext(self.h, frame, None, None, kmin.ctypes.data_as(i32p),
                                             kdim.ctypes.data_as(i32p)), "fr_download_frame_ext")
        sorted_ = np.zeros((n, 4), np.float32)
        cell_start = np.zeros(int(np.prod(kdim.astype(np.int64))) + 1, np.uint32)
        check(self.lib.fr_download_frame_ext(self.h, frame, _ptr(sorted_, abi.f32p), _ptr(cell_start, abi.u32p),
                                             None, None), "fr_download_frame_ext")
        return sorted_, cell_start, kmin, kdim

    def query_density(self, frame: int, points, want_grad=True):
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        m = pts.shape[0]
        rho = np.zeros(m, np.float32)
        grad = np.zeros((m, 3), np.float32) if want_grad else None
        check(self.lib.fr_query_density(self.h, frame, _ptr(pts, abi.f32p), m, _ptr(rho, abi.f32p),
                                        _ptr(grad, abi.f32p)), "fr_query_density")
        return rho, grad


    def smooth_depth(self, gauss_n: int = 8, inv_projection=None):
        """GaussRenderPass on the context's depth image (GaussN = 8 by default, GaussRenderPass.h:41); with
        inv_projection (Camera3D::InvProjection) also the Sobel screen normals of composition.frag:87-104"""
        H, W = self.height, self.width
        sm = np.empty((H, W), np.float32)
        nrm = np.empty((H, W, 4), np.float32) if inv_projection is not None else None
        ip = None if inv_projection is None else np.ascontiguousarray(inv_projection, np.float32).reshape(16)
        check(self.lib.fr_smooth_depth(self.h, gauss_n, _ptr(ip, abi.f32p), _ptr(sm, abi.f32p), _ptr(nrm, abi.f32p)), "fr_smooth_depth")
        return sm, nrm

    def measure_l2_bandwidth(self, mbytes: int = 32, reps: int = 40) -> float:
        """GB/s of read-only streaming over an L2-resident buffer (the roofline denominator of the march's gathers)"""
        g = C.c_float()
        check(self.lib.fr_measure_l2_bandwidth(self.h, mbytes << 20, reps, C.byref(g)), "fr_measure_l2_bandwidth")
        return float(g.value)

    def selftest_division(self, n: int = 1 << 26, seed: int = 1) -> int:
        bad = C.c_uint64(0)
        check(self.lib.fr_selftest_division(self.h, n, seed, C.byref(bad)), "fr_selftest_division")
        return int(bad.value)


class DeviceBuffer:
    """a plain device allocation that other processes of the box can map (fr_device_alloc / fr_ipc_*): frame rings and
    completion flags on the presenting GPU of a frame-parallel run"""

    def __init__(self, nbytes: int = 0, device: int = 0, handle: bytes | None = None):
        self.lib = abi.load()
        self.device = device
        self.owner = handle is None
        p = C.c_void_p()
        if self.owner:
            check(self.lib.fr_device_alloc(device, nbytes, C.byref(p)), "fr_device_alloc")
        else:
            buf = (C.c_ubyte * 64).from_buffer_copy(handle)
            check(self.lib.fr_ipc_open_buffer(device, buf, C.byref(p)), "fr_ipc_open_buffer")
        self.ptr = p.value
        self.nbytes = nbytes

    def export(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        check(self.lib.fr_ipc_export_buffer(self.device, C.c_void_p(self.ptr), buf), "fr_ipc_export_buffer")
        return bytes(buf)

    def close(self):
        if getattr(self, "ptr", None):
            if self.owner:
                self.lib.fr_device_free(self.device, C.c_void_p(self.ptr))
            else:
                self.lib.fr_ipc_close_buffer(self.device, C.c_void_p(self.ptr))
            self.ptr = None

    __del__ = close


def gauss_kernel(gauss_n: int = 8) -> np.ndarray:
    """ComputeGaussKernel (reference GaussRenderPass.cpp:25-66): (N+1, N+1) weights, [i, j]"""
    out = np.zeros((gauss_n + 1) * (gauss_n + 1), np.float32)
    check(abi.load().fr_gauss_kernel(gauss_n, _ptr(out, abi.f32p)), "fr_gauss_kernel")
    return out.reshape(gauss_n + 1, gauss_n + 1)


def bgeo_probe(path: str) -> dict:
    """header of a classic .bgeo file (fr_bgeo_probe)"""
    info = abi.FrBgeoInfo()
    check(abi.load().fr_bgeo_probe(os.fsencode(path), C.byref(info)), "fr_bgeo_probe")
    return {n: int(getattr(info, n)) for n, _ in info._fields_}


def bgeo_read(path: str) -> np.ndarray:
    """positions of a classic .bgeo file as (N, 3) float32 -- what Dataset::ReadFile keeps (reference Dataset.cpp:292-306)"""
    lib = abi.load()
    n = C.c_uint64()
    check(lib.fr_bgeo_read(os.fsencode(path), None, 0, C.byref(n)), "fr_bgeo_read")
    xyz = np.zeros((int(n.value), 3), np.float32)
    if n.value:
        check(lib.fr_bgeo_read(os.fsencode(path), _ptr(xyz, abi.f32p), n.value, C.byref(n)), "fr_bgeo_read")
    return xyz


def bgeo_write(path: str, xyz, compressed: bool = False):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    check(abi.load().fr_bgeo_write(os.fsencode(path), _ptr(xyz, abi.f32p), xyz.shape[0], 1 if compressed else 0), "fr_bgeo_write")


def dataset_count(prefix: str, suffix: str, count: int = -1) -> int:
    """files <prefix>1<suffix>, <prefix>2<suffix>, ... that exist (Dataset::Dataset, reference Dataset.cpp:186-203)"""
    return int(abi.load().fr_dataset_count(os.fsencode(prefix), os.fsencode(suffix), count))


class _LaneContext(Context):
    """A lane's context inside a Sequence: owned by the sequence, read-only use (counters, timings, frame info)."""

    def __init__(self, lib, handle, width, height, device):
        self.lib, self.h, self.width, self.height, self.device = lib, handle, width, height, device

    def close(self):
        self.h = None

    __del__ = close


class Sequence:
    """fr_seq_*: `lanes` frames of an animation in flight on one GPU (the autoplay loop of AdvancedRenderer::Render,
    reference AdvancedRenderer.cpp:275-298, pipelined).  Frame k renders on lane k % lanes."""

    def __init__(self, width: int, height: int, lanes: int = 2, device: int = 0):
        self.lib = abi.load()
        self.width, self.height, self.device = int(width), int(height), device
        h = C.c_void_p()
        check(self.lib.fr_seq_create(device, self.width, self.height, lanes, C.byref(h)), "fr_seq_create")
        self.h = h
        self.lanes = lanes
        self._keep = {}          # ticket -> arrays that must outlive the job

    def close(self):
        if getattr(self, "h", None):
            self.lib.fr_seq_destroy(self.h)
            self.h = None

    __del__ = close

    def set_yielding(self, on: bool):
        """lane workers poll pinned memory and yield the core instead of spinning in the driver (oversubscribed hosts)"""
        check(self.lib.fr_seq_set_yielding(self.h, 1 if on else 0), "fr_seq_set_yielding")

    def context(self, lane: int) -> Context:
        c = C.c_void_p()
        check(self.lib.fr_seq_context(self.h, lane, C.byref(c)), "fr_seq_context")
        return _LaneContext(self.lib, c, self.width, self.height, self.device)

    def set_camera(self, view, projection, inv_projection_view, position, direction):
        cam = _camera_struct(view, projection, inv_projection_view, position, direction)
        check(self.lib.fr_seq_set_camera(self.h, C.byref(cam)), "fr_seq_set_camera")

    def set_settings(self, s: VisualizationSettings):
        cs = s.to_c()
        check(self.lib.fr_seq_set_settings(self.h, C.byref(cs)), "fr_seq_set_settings")

    def submit_ptrs(self, xyz_ptr: int, n: int, h: float = 0.1, h_ext_mult: float = 2.0, on_device: bool = False,
                    passes: int = FR_PASS_ALL, depth: int = 0, positions: int = 0, normals: int = 0, rgba: int = 0,
                    bgeo_path: str | None = None, bmp_path: str | None = None, rgba_device: int = 0,
                    done_flag_device: int = 0, done_value: int = 0) -> int:
        job = abi.FrSeqJob(xyz_ptr or None, n, h, h_ext_mult, 1 if on_device else 0, passes, depth or None, positions or None,
                           normals or None, rgba or None, os.fsencode(bgeo_path) if bgeo_path else None,
                           os.fsencode(bmp_path) if bmp_path else None, rgba_device or None, done_flag_device or None, done_value)
        t = self.lib.fr_seq_submit(self.h, C.byref(job))
        if t < 0:
            check(int(t), "fr_seq_submit")
        return int(t)

    def submit(self, xyz, h: float = 0.1, h_ext_mult: float = 2.0, want=("rgba",), passes: int = FR_PASS_ALL):
        """numpy in, numpy out: returns (ticket, dict of output arrays that are valid after wait(ticket))"""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        H, W = self.height, self.width
        shapes = dict(depth=((H, W), np.float32), positions=((H, W, 4), np.float32), normals=((H, W, 4), np.float32),
                      rgba=((H, W, 4), np.uint8))
        out = {k: np.zeros(*shapes[k]) for k in want}
        t = self.submit_ptrs(xyz.ctypes.data, xyz.shape[0], h, h_ext_mult, False, passes,
                             **{k: v.ctypes.data for k, v in out.items()})
        self._keep[t] = (xyz, out)
        return t, out

    def submit_file(self, path: str, h: float = 0.1, h_ext_mult: float = 2.0, want=("rgba",), passes: int = FR_PASS_ALL):
        """a frame from a .bgeo file, read and decoded on the lane that renders it"""
        H, W = self.height, self.width
        shapes = dict(depth=((H, W), np.float32), positions=((H, W, 4), np.float32), normals=((H, W, 4), np.float32),
                      rgba=((H, W, 4), np.uint8))
        out = {k: np.zeros(*shapes[k]) for k in want}
        t = self.submit_ptrs(0, 0, h, h_ext_mult, False, passes, bgeo_path=path, **{k: v.ctypes.data for k, v in out.items()})
        self._keep[t] = (None, out)
        return t, out

    def wait(self, ticket: int):
        check(self.lib.fr_seq_wait(self.h, ticket), "fr_seq_wait")
        self._keep.pop(ticket, None)

    def drain(self):
        check(self.lib.fr_seq_drain(self.h), "fr_seq_drain")
        self._keep.clear()

    def timer_begin(self):
        check(self.lib.fr_seq_timer_begin(self.h), "fr_seq_timer_begin")

    def timer_end(self) -> float:
        ms = C.c_float()
        check(self.lib.fr_seq_timer_end(self.h, C.byref(ms)), "fr_seq_timer_end")
        return float(ms.value)


class Dataset:
    """Mirror of the reference Dataset (src/app/Dataset.h:67-105): a sequence of particle frames sharing one
    support radius.  ``Frames[i]`` are (N_i, 3) float32 arrays; uploading a frame builds its search grid,
    AABB and occupancy grid on the GPU (Frame::Frame, Dataset.cpp:9-24)."""

    def __init__(self, frames, particleRadius: float = 0.1, particleRadiusMultiplier: float = 2.0):
        self.ParticleRadius = float(particleRadius)
        self.ParticleRadiusExt = float(particleRadiusMultiplier) * float(particleRadius)
        self.ParticleRadiusMultiplier = float(particleRadiusMultiplier)
        self.Frames = [np.ascontiguousarray(f, dtype=np.float32).reshape(-1, 3) for f in frames]
        if not self.Frames:
            raise ValueError("Dataset needs at least one frame")
        self.MaxParticles = max(f.shape[0] for f in self.Frames)
        self.Loaded = True
        self._resident = {}     # id(Context) -> set of uploaded frame indices

    def ensure_resident(self, ctx: Context, frame: int):
        have = self._resident.setdefault(id(ctx), set())
        if frame not in have:
            ctx.upload_frame(frame, self.Frames[frame], self.ParticleRadius, self.ParticleRadiusMultiplier)
            have.add(frame)


class RayMarcher:
    """Drop-in for the reference RayMarcher (RayMarcher.h:28-44) on one B200.

    Prepare(settings, camera, dataset, positions, normals, depth): positions / normals are caller-owned
    (H, W, 4) float32 arrays that receive (p, 1) / (n, 1) for hits and zeros otherwise; depth is the caller's
    (H, W) float32 depth image (1.0 = empty), exactly as in RayMarcher.cpp:76-100.
    Start() returns immediately; IsDone() polls; outputs are valid once IsDone() returned True."""

    def __init__(self, extent, device: int = 0):
        w, h = extent     # Vulkan.SwapchainExtent in the reference (RayMarcher.cpp:86-89)
        self.ctx = Context(w, h, device)
        self._out = None
        self._running = False

    def Exit(self):
        self.ctx.close()

    def Prepare(self, settings: VisualizationSettings, camera, dataset: Dataset, positions, normals, depth):
        ctx = self.ctx
        for name, a, shape in (("positions", positions, 4), ("normals", normals, 4)):
            if a.dtype != np.float32 or a.size != ctx.width * ctx.height * shape or not a.flags.c_contiguous:
                raise ValueError(f"{name} must be a C-contiguous float32 array of W*H*4 elements")
        dataset.ensure_resident(ctx, settings.Frame)
        ctx.set_settings(settings)
        ctx.set_camera_controller(camera)
        ctx.set_depth(depth)
        self._out = (positions, normals)

    def Start(self):
        if self._out is None:
            raise RuntimeError("RayMarcher.Start before Prepare")
        self.ctx.render_async(FR_PASS_MARCH)
        self._running = True

    def IsDone(self) -> bool:
        if not self._running:
            return True
        if not self.ctx.is_done():
            return False
        positions, normals = self._out
        self.ctx.download_into(positions=positions, normals=normals)
        self._running = False
        return True
