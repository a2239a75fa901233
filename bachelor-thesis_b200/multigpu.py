"""Multi-GPU partitioning of the ray-march path: one process per GPU, one fr_context each.

The path shards with no reduction -- every pixel / frame has exactly one owner:

  * frame-parallel   frame f of an animation sequence is rendered by rank f mod G; no data-path collective is
                     needed to render; finished RGBA frames go to the presenting rank with send/recv.
  * tile-parallel    one large frame: particles + grid replicated, screen tiles dealt round-robin
                     (fr_set_tile_partition: tile t belongs to rank t mod G), RGBA tiles gathered on the presenting
                     rank (NCCL over NVLink on the box; gloo in the CPU tests) and merged by ownership mask.

Everything here is host-side plumbing over torch.distributed; the rendering itself is the C ABI
(bachelor-thesis_b200/raymarcher.py).  The reference has no counterpart (single process, src/app/ThreadPool.cpp).
"""
from __future__ import annotations

import numpy as np


def frames_of_rank(n_frames: int, rank: int, world: int) -> list[int]:
    """frame-parallel assignment: rank r renders frames r, r + G, r + 2G, ..."""
    if not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    return list(range(rank, n_frames, world))


def owner_of_frame(frame: int, world: int) -> int:
    return frame % world


def tile_owner_map(width: int, height: int, world: int, tile_w: int = 64, tile_h: int = 64) -> np.ndarray:
    """(H, W) int32 image of the rank that owns each pixel -- the rule fr_set_tile_partition applies on the device
    (csrc/fm_march.cu: pixel_active): row-major tile index t = (y // tile_h) * tiles_x + (x // tile_w), owner t mod G"""
    if tile_w < 32 or tile_h < 8 or tile_w % 32 or tile_h % 8:
        raise ValueError("tile_w must be a multiple of 32 and tile_h a multiple of 8")
    tiles_x = (width + tile_w - 1) // tile_w
    ty, tx = np.meshgrid(np.arange(height) // tile_h, np.arange(width) // tile_w, indexing="ij")
    return ((ty * tiles_x + tx) % world).astype(np.int32)


def merge_tiles(parts, owner):
    """parts[r]: the (H, W, C) image rank r rendered (only its own tiles are meaningful); returns the full image.
    Works on numpy arrays and torch tensors (owner must be of the matching kind)."""
    if hasattr(parts[0], "clone"):          # torch: torch.where keeps the merge on the device without a host sync
        import torch
        out = parts[0]
        for r in range(1, len(parts)):
            m = owner == r
            out = torch.where(m.reshape(m.shape + (1,) * (out.dim() - m.dim())), parts[r], out)
        return out
    out = parts[0].copy()
    for r in range(1, len(parts)):
        m = owner == r
        out[m] = parts[r][m]
    return out


def gather_tiles(local, rank: int, world: int, owner, dst: int = 0, group=None):
    """tile-parallel exchange step: every rank contributes its (H, W, C) image, rank `dst` returns the merged
    frame (others return None).  `local` is a torch tensor (CUDA with NCCL, CPU with gloo)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    bufs = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return merge_tiles(bufs, owner)


def collect_frames(render_frame, n_frames: int, rank: int, world: int, like, dst: int = 0, group=None):
    """frame-parallel sequence: `render_frame(f)` renders frame f into a tensor shaped like `like` (called only for
    this rank's frames).  Rank `dst` returns the list of all frames in order; the transfer of frame f overlaps the
    rendering of the next one (isend).  Other ranks return None."""
    import torch
    import torch.distributed as dist
    mine = frames_of_rank(n_frames, rank, world)
    pending = []
    out = [None] * n_frames if rank == dst else None
    recvs = []
    if rank == dst:
        for f in range(n_frames):
            o = owner_of_frame(f, world)
            if o != dst:
                buf = torch.empty_like(like)
                recvs.append((f, buf, dist.irecv(buf, src=o, group=group, tag=f)))
    for f in mine:
        img = render_frame(f)
        if rank == dst:
            out[f] = img.clone()
        else:
            keep = img.clone()               # the renderer reuses its target
            pending.append((keep, dist.isend(keep, dst=dst, group=group, tag=f)))
    for _, w in pending:
        w.wait()
    for f, buf, w in recvs:
        w.wait()
        out[f] = buf
    return out
