"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE, never the product path):

  * ``Oracle``  -> oracle/liboracle.so         plain-C restatement (oracle/fluid_oracle.c)
  * ``Ref``     -> oracle/_ref/libfluidref.so  the reference's own unmodified TUs + harness

Both are built by ``make -C oracle`` (``__graft_entry__.build()`` does it).  ``Ref`` exists only
where it was built in the container that has /root/reference; the .so travels to the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libfluidref.so")

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


def _fp(a):
    return None if a is None else a.ctypes.data_as(f32p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class FoSettings(C.Structure):
    _fields_ = [("max_steps", C.c_int32), ("step_size", C.c_float), ("iso_density", C.c_float),
                ("anisotropic", C.c_int32), ("k_n", C.c_float), ("k_r", C.c_float), ("k_s", C.c_float),
                ("n_eps", C.c_int32)]


class FoCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("pixels", "covered_rays", "hit_rays", "ray_steps",
                                          "skip_iterations", "candidates", "neighbours", "steps_outside_grid")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


@dataclass
class Settings:
    """VisualizationSettings defaults (src/app/AdvancedRenderer/AdvancedRenderer.cpp:18-28), isotropic."""
    max_steps: int = 128
    step_size: float = 0.009
    iso_density: float = 1.0
    anisotropic: int = 0
    k_n: float = 0.5
    k_r: float = 2.0
    k_s: float = 2000.0
    n_eps: int = 1

    def fo(self):
        return FoSettings(self.max_steps, self.step_size, self.iso_density, self.anisotropic,
                          self.k_n, self.k_r, self.k_s, self.n_eps)


# ----------------------------------------------------------------------------------------------
class Oracle:
    def __init__(self, path: str = ORACLE_SO):
        self.lib = L = C.CDLL(path)
        L.fo_W0.restype = C.c_float
        L.fo_W0.argtypes = [C.c_float]
        L.fo_W.restype = C.c_float
        L.fo_W.argtypes = [C.c_float, f32p]
        L.fo_gradW.argtypes = [C.c_float, f32p, f32p]
        L.fo_intersect_aabb.argtypes = [f32p] * 5
        L.fo_cos_half_pi.restype = C.c_float
        L.fo_cos_half_pi.argtypes = [C.c_float]
        L.fo_frame_create.restype = C.c_void_p
        L.fo_frame_create.argtypes = [f32p, C.c_size_t, C.c_float, C.c_float]
        L.fo_frame_destroy.argtypes = [C.c_void_p]
        L.fo_frame_num_particles.restype = C.c_size_t
        L.fo_frame_num_particles.argtypes = [C.c_void_p]
        L.fo_frame_info.argtypes = [C.c_void_p, f32p, f32p, C.POINTER(C.c_int32)]
        L.fo_frame_particles.argtypes = [C.c_void_p, f32p]
        L.fo_frame_grid.argtypes = [C.c_void_p, u32p, u8p]
        L.fo_query_cell.restype = C.c_int64
        L.fo_query_cell.argtypes = [C.c_void_p, f32p]
        L.fo_neighbors.restype = C.c_size_t
        L.fo_neighbors.argtypes = [C.c_void_p, f32p, C.c_int, u32p, C.c_size_t]
        L.fo_depth_prepass.argtypes = [C.c_void_p, C.c_int32, C.c_int32, f32p, f32p, f32p]
        L.fo_march.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(FoSettings), f32p, f32p, f32p,
                               f32p, f32p, f32p, u32p, C.POINTER(FoCounters), C.c_int]
        L.fo_shade.argtypes = [C.c_int32, C.c_int32, f32p, f32p, f32p, f32p, f32p, f32p, u8p]
        L.fo_gauss_kernel.argtypes = [C.c_int32, f32p]
        L.fo_gauss_depth.argtypes = [C.c_int32, C.c_int32, f32p, C.c_int32, f32p]
        L.fo_sobel_normals.argtypes = [C.c_int32, C.c_int32, f32p, f32p, f32p]
        L.fo_set_threads.argtypes = [C.c_int]
        L.fo_set_count_mode.argtypes = [C.c_int]
        L.fo_get_threads.restype = C.c_int
        # anisotropic path
        L.fo_frame_particles_ext.argtypes = [C.c_void_p, f32p]
        L.fo_eigen3.argtypes = [f32p, f32p, f32p]
        L.fo_wpca.argtypes = [C.c_float, C.c_float, C.POINTER(FoSettings), f32p, f32p, C.c_uint32, f32p]
        L.fo_det3.restype = C.c_float
        L.fo_det3.argtypes = [f32p]
        L.fo_aniso_W.restype = C.c_float
        L.fo_aniso_W.argtypes = [C.c_float, f32p, C.c_float, f32p]
        L.fo_aniso_gradW.argtypes = [C.c_float, f32p, C.c_float, f32p, f32p]
        L.fo_cubic_W.restype = C.c_float
        L.fo_cubic_W.argtypes = [C.c_float, f32p]
        for fn in (L.fo_sinf, L.fo_cosf):
            fn.restype = C.c_float
            fn.argtypes = [C.c_float]
        L.fo_atan2f.restype = C.c_float
        L.fo_atan2f.argtypes = [C.c_float, C.c_float]
        L.fo_set_trig_libm.argtypes = [C.c_int]
        L.fo_trig_selftest.restype = C.c_uint64
        L.fo_trig_selftest.argtypes = [C.c_int, C.c_float, C.c_float, C.c_uint64, C.c_uint32]

    # anisotropic path (RayMarcher.cpp:114-254, Kernel.cpp:55-125)
    def eigen3(self, c9):
        c9 = _f32(c9).reshape(9)
        ev, vec = np.zeros(3, np.float32), np.zeros(9, np.float32)
        self.lib.fo_eigen3(_fp(c9), _fp(ev), _fp(vec))
        return ev, vec

    def wpca(self, h, h_ext, settings, particle, nbr_xyz):
        particle, nbr_xyz = _f32(particle), _f32(nbr_xyz).reshape(-1, 3)
        g = np.zeros(9, np.float32)
        s = settings.fo()
        self.lib.fo_wpca(h, h_ext, C.byref(s), _fp(particle), _fp(nbr_xyz), nbr_xyz.shape[0], _fp(g))
        return g

    def det3(self, g9):
        g9 = _f32(g9)
        return np.float32(self.lib.fo_det3(_fp(g9)))

    def aniso_W(self, h, g9, det, r):
        g9, r = _f32(g9), _f32(r)
        return np.float32(self.lib.fo_aniso_W(h, _fp(g9), float(det), _fp(r)))

    def aniso_gradW(self, h, g9, det, r):
        g9, r = _f32(g9), _f32(r)
        out = np.zeros(3, np.float32)
        self.lib.fo_aniso_gradW(h, _fp(g9), float(det), _fp(r), _fp(out))
        return out

    def cubic_W(self, h, r):
        r = _f32(r)
        return np.float32(self.lib.fo_cubic_W(h, _fp(r)))

    # kernels
    def W0(self, h):
        return float(self.lib.fo_W0(h))

    def W(self, h, r):
        r = _f32(r)
        return np.float32(self.lib.fo_W(h, _fp(r)))

    def gradW(self, h, r):
        r = _f32(r)
        out = np.zeros(3, np.float32)
        self.lib.fo_gradW(h, _fp(r), _fp(out))
        return out

    def intersect_aabb(self, o, d, bmin, bmax):
        o, d, bmin, bmax = map(_f32, (o, d, bmin, bmax))
        out = np.zeros(3, np.float32)
        self.lib.fo_intersect_aabb(_fp(o), _fp(d), _fp(bmin), _fp(bmax), _fp(out))
        return out

    def cos_half_pi(self, s):
        return np.float32(self.lib.fo_cos_half_pi(float(s)))

    def frame(self, xyz, h=0.1, mult=2.0, count_mode=1):
        """count_mode 1 = centre box (== the oracle/_ref build, the CUDA default FR_COUNT_CENTRE_BOX), 0 = cell-exact
        (FR_COUNT_CELL_EXACT)"""
        self.lib.fo_set_count_mode(count_mode)
        try:
            return OracleFrame(self, xyz, h, mult)
        finally:
            self.lib.fo_set_count_mode(0)

    # screen-space smoothing (f5)
    def gauss_kernel(self, n):
        out = np.zeros((n + 1) * (n + 1), np.float32)
        self.lib.fo_gauss_kernel(n, _fp(out))
        return out.reshape(n + 1, n + 1)

    def gauss_depth(self, depth, n):
        depth = _f32(depth)
        H, W = depth.shape
        out = np.zeros((H, W), np.float32)
        self.lib.fo_gauss_depth(W, H, _fp(depth), n, _fp(out))
        return out

    def sobel_normals(self, smoothed, inv_proj):
        smoothed, inv_proj = _f32(smoothed), _f32(inv_proj)
        H, W = smoothed.shape
        out = np.zeros((H, W, 4), np.float32)
        self.lib.fo_sobel_normals(W, H, _fp(smoothed), _fp(inv_proj), _fp(out))
        return out

    def shade(self, W, H, pos4, nrm4, ipv, cam_pos, cam_dir, want_color=True):
        pos4, nrm4, ipv, cam_pos, cam_dir = map(_f32, (pos4, nrm4, ipv, cam_pos, cam_dir))
        color = np.zeros((H, W, 4), np.float32) if want_color else None
        rgba = np.zeros((H, W, 4), np.uint8)
        self.lib.fo_shade(W, H, _fp(pos4), _fp(nrm4), _fp(ipv), _fp(cam_pos), _fp(cam_dir),
                          _fp(color), rgba.ctypes.data_as(u8p))
        return color, rgba


class OracleFrame:
    def __init__(self, oracle: Oracle, xyz, h, mult):
        self.o = oracle
        self.L = oracle.lib
        xyz = _f32(xyz).reshape(-1, 3)
        self.h = float(h)
        self.ptr = self.L.fo_frame_create(_fp(xyz), xyz.shape[0], h, mult)
        if not self.ptr:
            raise RuntimeError("fo_frame_create failed")
        self.n = int(self.L.fo_frame_num_particles(self.ptr))
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        dims = np.zeros(3, np.int32)
        self.L.fo_frame_info(self.ptr, _fp(mn), _fp(mx), dims.ctypes.data_as(C.POINTER(C.c_int32)))
        self.min, self.max, self.dims = mn, mx, dims

    def __del__(self):
        if getattr(self, "ptr", None):
            self.L.fo_frame_destroy(self.ptr)
            self.ptr = None

    def particles(self):
        out = np.zeros((self.n, 3), np.float32)
        self.L.fo_frame_particles(self.ptr, _fp(out))
        return out

    def particles_ext(self):
        out = np.zeros((self.n, 3), np.float32)
        self.L.fo_frame_particles_ext(self.ptr, _fp(out))
        return out

    def grid(self):
        ncell = int(np.prod(self.dims.astype(np.int64)))
        counts = np.zeros(ncell, np.uint32)
        flags = np.zeros(ncell, np.uint8)
        self.L.fo_frame_grid(self.ptr, counts.ctypes.data_as(u32p), flags.ctypes.data_as(u8p))
        return counts, flags

    def query_cell(self, p):
        p = _f32(p)
        return int(self.L.fo_query_cell(self.ptr, _fp(p)))

    def neighbors(self, p, ext=0, cap=8192):
        p = _f32(p)
        out = np.zeros(cap, np.uint32)
        n = int(self.L.fo_neighbors(self.ptr, _fp(p), ext, out.ctypes.data_as(u32p), cap))
        return out[:min(n, cap)].copy()

    def depth_prepass(self, W, H, view, proj):
        view, proj = _f32(view), _f32(proj)
        depth = np.zeros((H, W), np.float32)
        rc = self.L.fo_depth_prepass(self.ptr, W, H, _fp(view), _fp(proj), _fp(depth))
        if rc != 0:
            raise RuntimeError(f"fo_depth_prepass rc={rc}")
        return depth

    def march(self, W, H, settings: Settings, ipv, cam_pos, depth, threads=0, want_band=True):
        ipv, cam_pos, depth = _f32(ipv), _f32(cam_pos), _f32(depth)
        pos = np.zeros((H, W, 4), np.float32)
        nrm = np.zeros((H, W, 4), np.float32)
        band = np.zeros((H, W), np.float32) if want_band else None
        steps = np.zeros((H, W), np.uint32) if want_band else None
        cnt = FoCounters()
        s = settings.fo()
        rc = self.L.fo_march(self.ptr, W, H, C.byref(s), _fp(ipv), _fp(cam_pos), _fp(depth), _fp(pos), _fp(nrm),
                             _fp(band), None if steps is None else steps.ctypes.data_as(u32p), C.byref(cnt), threads)
        if rc != 0:
            raise RuntimeError(f"fo_march rc={rc}")
        return pos, nrm, band, steps, cnt.as_dict()


# ----------------------------------------------------------------------------------------------
def ref_available() -> bool:
    return os.path.exists(REF_SO)


class Ref:
    """The reference's own code (oracle/_ref/libfluidref.so)."""

    def __init__(self, path: str = REF_SO):
        self.lib = L = C.CDLL(path)
        L.ref_W.restype = C.c_float
        L.ref_W.argtypes = [C.c_float, f32p]
        L.ref_W0.restype = C.c_float
        L.ref_W0.argtypes = [C.c_float]
        L.ref_gradW.argtypes = [C.c_float, f32p, f32p]
        L.ref_intersectAABB.argtypes = [f32p] * 5
        L.ref_camera.argtypes = [C.c_float] * 7 + [f32p] * 6
        L.ref_dataset_create.restype = C.c_void_p
        L.ref_dataset_create.argtypes = [f32p, C.c_size_t, C.c_float, C.c_float, C.c_int, C.POINTER(C.c_double)]
        L.ref_dataset_load.restype = C.c_void_p
        L.ref_dataset_load.argtypes = [C.c_char_p, C.c_char_p, C.c_float, C.c_float, C.c_int, C.c_int]
        L.ref_dataset_destroy.argtypes = [C.c_void_p]
        L.ref_num_frames.argtypes = [C.c_void_p]
        L.ref_frame_num_particles.restype = C.c_size_t
        L.ref_frame_num_particles.argtypes = [C.c_void_p, C.c_int]
        L.ref_frame_info.argtypes = [C.c_void_p, C.c_int, f32p, f32p, C.POINTER(C.c_int)]
        L.ref_frame_particles.argtypes = [C.c_void_p, C.c_int, f32p]
        L.ref_frame_grid.argtypes = [C.c_void_p, C.c_int, u32p, u8p]
        L.ref_query_cell.restype = C.c_int64
        L.ref_query_cell.argtypes = [C.c_void_p, C.c_int, f32p]
        L.ref_neighbors.restype = C.c_size_t
        L.ref_neighbors.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int, u32p, C.c_size_t]
        L.ref_march.restype = C.c_double
        L.ref_march.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                C.c_float, C.c_float, C.c_float, C.c_int, f32p, f32p, f32p, f32p, f32p,
                                C.c_int, C.c_int]
        L.ref_hardware_threads.restype = C.c_int
        L.ref_abi_version.restype = C.c_int
        self.has_aniso = L.ref_abi_version() >= 2
        self.has_partio = L.ref_abi_version() >= 3
        self.has_bmp = L.ref_abi_version() >= 4
        if self.has_bmp:
            L.ref_write_bmp.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        if self.has_partio:
            L.ref_partio_read.restype = C.c_long
            L.ref_partio_read.argtypes = [C.c_char_p, f32p, C.c_size_t]
            L.ref_partio_write.argtypes = [C.c_char_p, f32p, C.c_size_t, C.c_int, C.c_int]
        if self.has_aniso:
            L.ref_eigen3.argtypes = [f32p, f32p, f32p]
            L.ref_wpca.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, f32p, f32p, C.c_uint32, f32p]
            L.ref_det3.restype = C.c_float
            L.ref_det3.argtypes = [f32p]
            L.ref_aniso_W.restype = C.c_float
            L.ref_aniso_W.argtypes = [C.c_float, f32p, C.c_float, f32p]
            L.ref_aniso_gradW.argtypes = [C.c_float, f32p, C.c_float, f32p, f32p]
            L.ref_cubic_W.restype = C.c_float
            L.ref_cubic_W.argtypes = [C.c_float, f32p]
            L.ref_frame_particles_ext.argtypes = [C.c_void_p, C.c_int, f32p]

    def write_bmp(self, path, rgba):
        """stbi_write_bmp(path, W, H, 4, rgba) of the reference's vendored stb (Renderer.cpp:400-409)"""
        rgba = np.ascontiguousarray(rgba, np.uint8)
        h, w = rgba.shape[:2]
        return self.lib.ref_write_bmp(os.fsencode(path), w, h, rgba.ctypes.data)

    def partio_read(self, path):
        """positions of a particle file through the reference's vendored partio (Dataset.cpp:209-220, 292-306)"""
        n = self.lib.ref_partio_read(os.fsencode(path), None, 0)
        if n < 0:
            return None
        out = np.zeros((n, 3), np.float32)
        self.lib.ref_partio_read(os.fsencode(path), _fp(out), n)
        return out

    def partio_write(self, path, xyz, extras=0, compressed=False):
        xyz = _f32(xyz).reshape(-1, 3)
        self.lib.ref_partio_write(os.fsencode(path), _fp(xyz), xyz.shape[0], extras, 1 if compressed else 0)

    def eigen3(self, c9):
        c9 = _f32(c9).reshape(9)
        ev, vec = np.zeros(3, np.float32), np.zeros(9, np.float32)
        self.lib.ref_eigen3(_fp(c9), _fp(ev), _fp(vec))
        return ev, vec

    def det3(self, g9):
        g9 = _f32(g9)
        return np.float32(self.lib.ref_det3(_fp(g9)))

    def aniso_W(self, h, g9, det, r):
        g9, r = _f32(g9), _f32(r)
        return np.float32(self.lib.ref_aniso_W(h, _fp(g9), float(det), _fp(r)))

    def aniso_gradW(self, h, g9, det, r):
        g9, r = _f32(g9), _f32(r)
        out = np.zeros(3, np.float32)
        self.lib.ref_aniso_gradW(h, _fp(g9), float(det), _fp(r), _fp(out))
        return out

    def cubic_W(self, h, r):
        r = _f32(r)
        return np.float32(self.lib.ref_cubic_W(h, _fp(r)))

    def W(self, h, r):
        r = _f32(r)
        return np.float32(self.lib.ref_W(h, _fp(r)))

    def W0(self, h):
        return float(self.lib.ref_W0(h))

    def gradW(self, h, r):
        r = _f32(r)
        out = np.zeros(3, np.float32)
        self.lib.ref_gradW(h, _fp(r), _fp(out))
        return out

    def intersect_aabb(self, o, d, bmin, bmax):
        o, d, bmin, bmax = map(_f32, (o, d, bmin, bmax))
        out = np.zeros(3, np.float32)
        self.lib.ref_intersectAABB(_fp(o), _fp(d), _fp(bmin), _fp(bmax), _fp(out))
        return out

    def camera(self, fov, aspect, near, far, R=10.0, rot_x=0.0, rot_y=0.0):
        view, proj, ip, ipv = (np.zeros(16, np.float32) for _ in range(4))
        pos, system = np.zeros(3, np.float32), np.zeros(9, np.float32)
        self.lib.ref_camera(fov, aspect, near, far, R, rot_x, rot_y, _fp(view), _fp(proj), _fp(ip), _fp(ipv),
                            _fp(pos), _fp(system))
        return dict(view=view, proj=proj, inv_proj=ip, inv_proj_view=ipv, position=pos, system=system)

    def dataset(self, xyz, h=0.1, mult=2.0, box_mode=0):
        return RefDataset(self, xyz, h, mult, box_mode)


class RefDataset:
    def __init__(self, ref: Ref, xyz, h, mult, box_mode):
        self.L = ref.lib
        xyz = _f32(xyz).reshape(-1, 3)
        secs = C.c_double(0)
        self.ptr = self.L.ref_dataset_create(_fp(xyz), xyz.shape[0], h, mult, box_mode, C.byref(secs))
        self.build_seconds = secs.value
        self.n = int(self.L.ref_frame_num_particles(self.ptr, 0))
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        dims = np.zeros(3, np.int32)
        self.L.ref_frame_info(self.ptr, 0, _fp(mn), _fp(mx), dims.ctypes.data_as(C.POINTER(C.c_int)))
        self.min, self.max, self.dims = mn, mx, dims

    def close(self):
        if self.ptr:
            self.L.ref_dataset_destroy(self.ptr)
            self.ptr = None

    def particles(self):
        out = np.zeros((self.n, 3), np.float32)
        self.L.ref_frame_particles(self.ptr, 0, _fp(out))
        return out

    def particles_ext(self):
        out = np.zeros((self.n, 3), np.float32)
        self.L.ref_frame_particles_ext(self.ptr, 0, _fp(out))
        return out

    def wpca(self, settings, particle, nbr_xyz):
        particle, nbr_xyz = _f32(particle), _f32(nbr_xyz).reshape(-1, 3)
        g = np.zeros(9, np.float32)
        s = settings
        self.L.ref_wpca(self.ptr, s.k_n, s.k_r, s.k_s, s.n_eps, _fp(particle), _fp(nbr_xyz), nbr_xyz.shape[0], _fp(g))
        return g

    def grid(self):
        ncell = int(np.prod(self.dims.astype(np.int64)))
        counts = np.zeros(ncell, np.uint32)
        flags = np.zeros(ncell, np.uint8)
        self.L.ref_frame_grid(self.ptr, 0, counts.ctypes.data_as(u32p), flags.ctypes.data_as(u8p))
        return counts, flags

    def query_cell(self, p):
        p = _f32(p)
        return int(self.L.ref_query_cell(self.ptr, 0, _fp(p)))

    def neighbors(self, p, ext=0, cap=8192):
        p = _f32(p)
        out = np.zeros(cap, np.uint32)
        n = int(self.L.ref_neighbors(self.ptr, 0, _fp(p), ext, out.ctypes.data_as(u32p), cap))
        return out[:min(n, cap)].copy()

    def march(self, W, H, settings: Settings, ipv, cam_pos, depth, threads=0, use_ref_pool=0):
        ipv, cam_pos, depth = _f32(ipv), _f32(cam_pos), _f32(depth)
        pos = np.zeros((H, W, 4), np.float32)
        nrm = np.zeros((H, W, 4), np.float32)
        s = settings
        secs = self.L.ref_march(self.ptr, 0, W, H, s.max_steps, s.step_size, s.iso_density, s.anisotropic,
                                s.k_n, s.k_r, s.k_s, s.n_eps, _fp(ipv), _fp(cam_pos), _fp(depth), _fp(pos), _fp(nrm),
                                threads, use_ref_pool)
        return pos, nrm, float(secs)
