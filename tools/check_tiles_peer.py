"""Tile-parallel over peer memory, parity on N GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_tiles_peer.py

Every rank renders its interleaved 64x64 tiles of one frame straight into rank 0's colour image (fr_ipc_*); rank 0
compares that image with its own single-GPU render of the whole frame: bit-identical, isotropic and anisotropic."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
fm = importlib.import_module("bachelor-thesis_b200")
from conftest import golden_camera  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = 1280, 720
    cam = golden_camera("camera_default_16x9")
    cam_args = (cam["view"], cam["proj"], cam["inv_proj_view"], cam["position"], cam["system"].reshape(3, 3)[2])
    xyz = fm.scenes.dam_break(64_000)
    ok = True
    for aniso in (False, True):
        ctx = fm.Context(W, H, device=local)
        ctx.set_camera(*cam_args)
        ctx.set_settings(fm.VisualizationSettings(EnableAnisotropy=aniso))
        ctx.upload_frame(0, xyz, 0.1, 2.0)
        want = None
        if rank == 0:
            ctx.render(fm.FR_PASS_ALL)
            want = ctx.download(False, False, False, True)[3].copy()
            # poison the image so stale pixels cannot pass
            poison = torch.full((H, W, 4), 0x5a, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            import ctypes
            ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(ctx.device_images()["rgba"]), ctypes.c_void_p(poison.data_ptr()),
                                                   ctypes.c_size_t(H * W * 4), ctypes.c_int(3))
        handle = [ctx.ipc_export_color() if rank == 0 else None]
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            ctx.ipc_open_color_target(handle[0])
        ctx.set_tile_partition(rank, world, 64, 64)
        dist.barrier()
        ctx.render(fm.FR_PASS_ALL)
        ctx.wait()
        dist.barrier()
        if rank == 0:
            got = ctx.download(False, False, False, True)[3]
            same = bool(np.array_equal(got, want))
            print(f"tile-parallel over peer memory, {world} GPUs, {'anisotropic' if aniso else 'isotropic'}: "
                  f"{'bit-identical' if same else 'MISMATCH'} ({int((got != want).any(-1).sum())} pixels differ)", flush=True)
            ok = ok and same
        dist.barrier()
        if rank != 0:
            ctx.ipc_close_color_target()
        ctx.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
