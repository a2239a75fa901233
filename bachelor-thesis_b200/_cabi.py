"""ctypes binding of libfluidmarch.so (include/fluidmarch.h).  Fails loudly when the CUDA library
is missing or no sm_100 device is present -- there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLUIDMARCH_LIB") or os.path.join(HERE, "libfluidmarch.so")   # override: A/B builds of the kernels

FR_OK = 0
FR_PASS_DEPTH, FR_PASS_MARCH, FR_PASS_SHADE, FR_PASS_ALL = 1, 2, 4, 7
FR_ERR_NO_DEVICE = -3
FR_ERR_INVALID = -1
FR_COUNT_CELL_EXACT, FR_COUNT_CENTRE_BOX = 0, 1

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
vpp = C.POINTER(C.c_void_p)


class FrSettings(C.Structure):
    """fr_settings: POD mirror of VisualizationSettings (reference RayMarcher.h:12-26)."""
    _fields_ = [("frame", C.c_int32), ("max_steps", C.c_int32), ("step_size", C.c_float),
                ("iso_density", C.c_float), ("enable_anisotropy", C.c_int32),
                ("k_n", C.c_float), ("k_r", C.c_float), ("k_s", C.c_float), ("n_eps", C.c_int32),
                ("bisection_steps", C.c_int32), ("skip_last_pixel", C.c_int32), ("fast_normals", C.c_int32)]


class FrCamera(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("projection", C.c_float * 16),
                ("inv_projection_view", C.c_float * 16), ("position", C.c_float * 3),
                ("direction", C.c_float * 3)]


class FrFrameInfo(C.Structure):
    _fields_ = [("num_particles", C.c_uint64), ("h", C.c_float), ("min", C.c_float * 3), ("max", C.c_float * 3),
                ("grid_dims", C.c_int32 * 3), ("occupied_cells", C.c_uint64),
                ("search_min", C.c_int32 * 3), ("search_dims", C.c_int32 * 3)]


class FrCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("pixels", "covered_rays", "hit_rays", "ray_steps", "skip_iterations",
                                          "candidates", "neighbours", "early_exits", "neighbour_overflow",
                                          "kernel_launches", "first_candidates", "queued_rays",
                                          "first_examined", "first_fallbacks")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class FrTimings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("upload_ms", "grid_ms", "depth_ms", "march_ms", "download_ms",
                                          "classify_ms", "march_first_ms", "march_long_ms")]

    def as_dict(self):
        return {n: float(getattr(self, n)) for n, _ in self._fields_}


class FrSeqJob(C.Structure):
    """fr_seq_job: one frame of a sequence (particles in, host images out)."""
    _fields_ = [("xyz", C.c_void_p), ("n", C.c_uint64), ("h", C.c_float), ("h_ext_mult", C.c_float),
                ("xyz_on_device", C.c_int32), ("passes", C.c_int32),
                ("depth", C.c_void_p), ("positions", C.c_void_p), ("normals", C.c_void_p), ("rgba", C.c_void_p),
                ("bgeo_path", C.c_char_p), ("bmp_path", C.c_char_p),
                ("rgba_device", C.c_void_p), ("done_flag_device", C.c_void_p), ("done_value", C.c_uint32)]


class FrBgeoInfo(C.Structure):
    _fields_ = [("num_particles", C.c_uint64), ("record_words", C.c_uint32), ("compressed", C.c_int32),
                ("data_offset", C.c_uint64), ("file_bytes", C.c_uint64)]


# every symbol include/fluidmarch.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("fr_abi_version", C.c_int, []),
    ("fr_last_error", C.c_char_p, []),
    ("fr_create", C.c_int, [C.c_int, C.c_int, C.c_int, vpp]),
    ("fr_resize", C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    ("fr_destroy", None, [C.c_void_p]),
    ("fr_host_alloc", C.c_int, [C.c_size_t, vpp]),
    ("fr_host_free", None, [C.c_void_p]),
    ("fr_upload_frame", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_float, C.c_float]),
    ("fr_build_frame_device", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_float, C.c_float]),
    ("fr_set_count_mode", C.c_int, [C.c_void_p, C.c_int]),
    ("fr_set_async_build", C.c_int, [C.c_void_p, C.c_int]),
    ("fr_set_stage_timing", C.c_int, [C.c_void_p, C.c_int]),
    ("fr_get_frame_info", C.c_int, [C.c_void_p, C.c_int, C.POINTER(FrFrameInfo)]),
    ("fr_release_frame", C.c_int, [C.c_void_p, C.c_int]),
    ("fr_download_frame", C.c_int, [C.c_void_p, C.c_int, f32p, u32p, u32p, u8p]),
    ("fr_set_settings", C.c_int, [C.c_void_p, C.POINTER(FrSettings)]),
    ("fr_set_camera", C.c_int, [C.c_void_p, C.POINTER(FrCamera)]),
    ("fr_set_depth", C.c_int, [C.c_void_p, C.c_void_p]),
    ("fr_set_tile_partition", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("fr_set_region_partition", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("fr_render_async", C.c_int, [C.c_void_p, C.c_int]),
    ("fr_is_done", C.c_int, [C.c_void_p]),
    ("fr_wait", C.c_int, [C.c_void_p]),
    ("fr_download", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("fr_device_images", C.c_int, [C.c_void_p, vpp, vpp, vpp, vpp]),
    ("fr_set_color_target", C.c_int, [C.c_void_p, C.c_void_p]),
    ("fr_ipc_export_color", C.c_int, [C.c_void_p, C.c_void_p]),
    ("fr_ipc_open_color_target", C.c_int, [C.c_void_p, C.c_void_p]),
    ("fr_ipc_close_color_target", C.c_int, [C.c_void_p]),
    ("fr_device_alloc", C.c_int, [C.c_int, C.c_size_t, vpp]),
    ("fr_device_free", C.c_int, [C.c_int, C.c_void_p]),
    ("fr_ipc_export_buffer", C.c_int, [C.c_int, C.c_void_p, C.c_void_p]),
    ("fr_ipc_open_buffer", C.c_int, [C.c_int, C.c_void_p, vpp]),
    ("fr_ipc_close_buffer", C.c_int, [C.c_int, C.c_void_p]),
    ("fr_get_counters", C.c_int, [C.c_void_p, C.POINTER(FrCounters)]),
    ("fr_get_timings", C.c_int, [C.c_void_p, C.POINTER(FrTimings)]),
    ("fr_get_stream", C.c_int, [C.c_void_p, vpp]),
    ("fr_query_neighbors", C.c_int, [C.c_void_p, C.c_int, f32p, C.c_size_t, u32p, u32p, C.c_size_t]),
    ("fr_query_density", C.c_int, [C.c_void_p, C.c_int, f32p, C.c_size_t, f32p, f32p]),
    ("fr_query_neighbors_ext", C.c_int, [C.c_void_p, C.c_int, f32p, C.c_size_t, u32p, u32p, C.c_size_t]),
    ("fr_query_anisotropic", C.c_int, [C.c_void_p, C.c_int, f32p, C.c_size_t, f32p, f32p, f32p]),
    ("fr_download_frame_ext", C.c_int, [C.c_void_p, C.c_int, f32p, u32p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("fr_measure_l2_bandwidth", C.c_int, [C.c_void_p, C.c_size_t, C.c_uint32, f32p]),
    ("fr_selftest_division", C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
    ("fr_import_vk_memory_fd", C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t]),
    ("fr_import_vk_images_fd", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t]),
    ("fr_import_vk_semaphores_fd", C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    ("fr_bgeo_probe", C.c_int, [C.c_char_p, C.POINTER(FrBgeoInfo)]),
    ("fr_bgeo_read", C.c_int, [C.c_char_p, f32p, C.c_uint64, C.POINTER(C.c_uint64)]),
    ("fr_bgeo_write", C.c_int, [C.c_char_p, f32p, C.c_uint64, C.c_int]),
    ("fr_dataset_count", C.c_int, [C.c_char_p, C.c_char_p, C.c_int]),
    ("fr_upload_frame_bgeo", C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_float, C.c_float]),
    ("fr_encode_bmp", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    ("fr_write_bmp", C.c_int, [C.c_void_p, C.c_char_p]),
    ("fr_gauss_kernel", C.c_int, [C.c_int, f32p]),
    ("fr_smooth_depth", C.c_int, [C.c_void_p, C.c_int, f32p, f32p, f32p]),
    ("fr_seq_create", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, vpp]),
    ("fr_seq_destroy", None, [C.c_void_p]),
    ("fr_seq_lanes", C.c_int, [C.c_void_p]),
    ("fr_seq_set_yielding", C.c_int, [C.c_void_p, C.c_int]),
    ("fr_seq_context", C.c_int, [C.c_void_p, C.c_int, vpp]),
    ("fr_seq_set_camera", C.c_int, [C.c_void_p, C.POINTER(FrCamera)]),
    ("fr_seq_set_settings", C.c_int, [C.c_void_p, C.POINTER(FrSettings)]),
    ("fr_seq_submit", C.c_int64, [C.c_void_p, C.POINTER(FrSeqJob)]),
    ("fr_seq_wait", C.c_int, [C.c_void_p, C.c_int64]),
    ("fr_seq_drain", C.c_int, [C.c_void_p]),
    ("fr_seq_timer_begin", C.c_int, [C.c_void_p]),
    ("fr_seq_timer_end", C.c_int, [C.c_void_p, f32p]),
]

_lib = None


class FluidMarchError(RuntimeError):
    pass


def load(path: str = LIB_PATH):
    """Loads the CUDA library; raises (never falls back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise FluidMarchError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, restype, argtypes in SYMBOLS:
        if os.environ.get("FLUIDMARCH_AB") and not hasattr(lib, name):
            continue              # A/B runs against an older build (tools/ab_bench.sh); never set in tests or the bench
        fn = getattr(lib, name)   # AttributeError here == header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != FR_OK:
        msg = load().fr_last_error()
        raise FluidMarchError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
