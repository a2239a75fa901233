"""fluidmarch: B200-native SPH iso-surface ray marcher (drop-in for the CPU ray-march path of
Fruup/bachelor-thesis).  The compute lives in libfluidmarch.so (hand-written sm_100a CUDA, C ABI in
include/fluidmarch.h); this package is the host-side mirror of the reference's operator interface.
The directory name has a hyphen: import it with ``importlib.import_module("bachelor-thesis_b200")``."""
from . import _cabi, camera, scenes
from ._cabi import (FR_COUNT_CELL_EXACT, FR_COUNT_CENTRE_BOX, FR_PASS_ALL, FR_PASS_DEPTH, FR_PASS_MARCH, FR_PASS_SHADE, FluidMarchError, LIB_PATH, load)
from .camera import Camera3D, CameraController3D
from .raymarcher import (Context, Dataset, DeviceBuffer, RayMarcher, Sequence, VisualizationSettings, bgeo_probe, bgeo_read, bgeo_write, gauss_kernel,
                         dataset_count)

__all__ = ["Camera3D", "CameraController3D", "Context", "Dataset", "DeviceBuffer", "RayMarcher", "Sequence", "VisualizationSettings", "gauss_kernel", "bgeo_probe", "bgeo_read", "bgeo_write", "dataset_count",
           "FR_COUNT_CELL_EXACT", "FR_COUNT_CENTRE_BOX", "FR_PASS_ALL", "FR_PASS_DEPTH", "FR_PASS_MARCH", "FR_PASS_SHADE", "FluidMarchError", "LIB_PATH",
           "load", "scenes", "camera"]
