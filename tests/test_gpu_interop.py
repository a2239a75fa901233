"""CUDA-Vulkan hand-off without a Vulkan driver: the exporting side is played by CUDA's own virtual
memory API -- cuMemCreate with a POSIX-fd shareable handle exports device memory as a file descriptor, the way
vkGetMemoryFdKHR does for VkDeviceMemory allocated with VkExportMemoryAllocateInfo{OPAQUE_FD}.  fr_import_vk_memory_fd
imports that fd (cudaImportExternalMemory, opaque fd) and renders the colour image straight into it; the exporter reads
its own mapping back.  Skipped where the driver refuses the import of a VMM-exported fd as an opaque-fd external memory
(the real Vulkan path then remains unverified on this box, as DESIGN.md states)."""
import os

import numpy as np
import pytest

from conftest import golden_camera

pytestmark = pytest.mark.gpu


def _check(res):
    err, *vals = res
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return vals[0] if len(vals) == 1 else vals


def test_color_image_lands_in_imported_external_memory(fm, gpu_ctx_factory):
    try:
        try:
            from cuda.bindings import driver as cu
        except Exception:
            from cuda import cuda as cu
    except Exception:
        pytest.skip("cuda-python is not importable")
    W, H = 320, 180
    ctx = gpu_ctx_factory(W, H)                       # creates the primary context on device 0
    cam = golden_camera("camera_close_16x9")
    ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
    ctx.set_settings(fm.VisualizationSettings())
    ctx.upload_frame(0, fm.scenes.dam_break(8000), 0.1, 2.0)
    ctx.render(fm.FR_PASS_ALL)
    want = ctx.download(False, False, False, True)[3].copy()

    _check(cu.cuInit(0))
    prop = cu.CUmemAllocationProp()
    prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = 0
    prop.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    gran = _check(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
    size = ((W * H * 4 + gran - 1) // gran) * gran
    handle = _check(cu.cuMemCreate(size, prop, 0))
    fd = int(_check(cu.cuMemExportToShareableHandle(handle, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0)))
    # the exporter's own view of the memory
    va = _check(cu.cuMemAddressReserve(size, 0, 0, 0))
    _check(cu.cuMemMap(va, size, 0, handle, 0))
    acc = cu.CUmemAccessDesc()
    acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    acc.location.id = 0
    acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    _check(cu.cuMemSetAccess(va, size, [acc], 1))
    _check(cu.cuMemsetD8(va, 0x5a, size))
    try:
        ctx_import_ok = True
        try:
            fm._cabi.check(ctx.lib.fr_import_vk_memory_fd(ctx.h, os.dup(fd), size, 0), "fr_import_vk_memory_fd")
        except fm.FluidMarchError as e:
            ctx_import_ok = False
            reason = str(e)
        if not ctx_import_ok:
            pytest.skip("the driver does not import a VMM-exported fd as opaque-fd external memory: " + reason[:200])
        ctx.render(fm.FR_PASS_ALL)                     # colour target = the imported memory
        ctx.wait()
        got = np.zeros((H, W, 4), np.uint8)
        _check(cu.cuMemcpyDtoH(got.ctypes.data, va, W * H * 4))
        assert np.array_equal(got, want)
        ctx.set_color_target(None)
    finally:
        os.close(fd)
        cu.cuMemUnmap(va, size)
        cu.cuMemAddressFree(va, size)
        cu.cuMemRelease(handle)


def test_positions_and_normals_land_in_imported_external_memory(fm, gpu_ctx_factory):
    """fr_import_vk_images_fd: the two RGBA32F images of the reference's composition pass as exported buffers"""
    try:
        try:
            from cuda.bindings import driver as cu
        except Exception:
            from cuda import cuda as cu
    except Exception:
        pytest.skip("cuda-python is not importable")
    W, H = 320, 180
    ctx = gpu_ctx_factory(W, H)
    cam = golden_camera("camera_close_16x9")
    ctx.set_camera(cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
    ctx.set_settings(fm.VisualizationSettings())
    ctx.upload_frame(0, fm.scenes.dam_break(8000), 0.1, 2.0)
    ctx.render(fm.FR_PASS_ALL)
    _, want_pos, want_nrm, _ = ctx.download()
    want_pos, want_nrm = want_pos.copy(), want_nrm.copy()

    _check(cu.cuInit(0))
    prop = cu.CUmemAllocationProp()
    prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = 0
    prop.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    gran = _check(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
    size = ((W * H * 16 + gran - 1) // gran) * gran
    made = []
    try:
        for _ in range(2):
            handle = _check(cu.cuMemCreate(size, prop, 0))
            fd = int(_check(cu.cuMemExportToShareableHandle(handle, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0)))
            va = _check(cu.cuMemAddressReserve(size, 0, 0, 0))
            _check(cu.cuMemMap(va, size, 0, handle, 0))
            acc = cu.CUmemAccessDesc()
            acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
            acc.location.id = 0
            acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
            _check(cu.cuMemSetAccess(va, size, [acc], 1))
            _check(cu.cuMemsetD8(va, 0x5a, size))
            made.append((handle, fd, va))
        try:
            fm._cabi.check(ctx.lib.fr_import_vk_images_fd(ctx.h, os.dup(made[0][1]), os.dup(made[1][1]), size), "fr_import_vk_images_fd")
        except fm.FluidMarchError as e:
            pytest.skip("the driver does not import a VMM-exported fd as opaque-fd external memory: " + str(e)[:200])
        ctx.render(fm.FR_PASS_ALL)
        ctx.wait()
        got = [np.zeros((H, W, 4), np.float32) for _ in range(2)]
        for g, (_, _, va) in zip(got, made):
            _check(cu.cuMemcpyDtoH(g.ctypes.data, va, W * H * 16))
        assert np.array_equal(got[0].view(np.uint32), want_pos.view(np.uint32))
        assert np.array_equal(got[1].view(np.uint32), want_nrm.view(np.uint32))
        # fr_download reads the same (imported) images
        _, p2, n2, _ = ctx.download()
        assert np.array_equal(p2.view(np.uint32), want_pos.view(np.uint32)) and np.array_equal(n2.view(np.uint32), want_nrm.view(np.uint32))
    finally:
        ctx.resize(W, H)                                # drops the imported images before their memory goes away
        for handle, fd, va in made:
            os.close(fd)
            cu.cuMemUnmap(va, size)
            cu.cuMemAddressFree(va, size)
            cu.cuMemRelease(handle)
