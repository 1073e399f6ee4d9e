"""Host-side logic of the multi-GPU path: which 8x4-pixel ray tiles a rank owns, and the one
exchange step that combines the per-rank radiance frames on rank 0.

The device side mirrors `tile_owner` exactly (dt_rank_tile in csrc/dt_kernels.cuh: strips of TILE_GROUP tiles are dealt out
round-robin, so a rank writes long contiguous row segments into rank 0's frame).  The
combine step is a single reduce(SUM): every pixel is non-zero on exactly one rank, so the sum is a gather.
`combine_frames` works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np

TILE_W, TILE_H = 8, 4
TILE_GROUP = 8


def tile_grid(width, height):
    return (width + TILE_W - 1) // TILE_W, (height + TILE_H - 1) // TILE_H


def tile_owner(width, height, world):
    """[tiles_y, tiles_x] array of owning ranks: each tile row is cut into strips of TILE_GROUP tiles, the strips are dealt
    out round-robin in row-major strip order (dt_rank_tile, csrc/dt_kernels.cuh)."""
    tx, ty = tile_grid(width, height)
    strips_x = (tx + TILE_GROUP - 1) // TILE_GROUP
    strip = np.arange(ty, dtype=np.int64)[:, None] * strips_x + (np.arange(tx, dtype=np.int64) // TILE_GROUP)[None, :]
    return strip % world


def pixel_owner(width, height, world):
    own = tile_owner(width, height, world)
    return np.repeat(np.repeat(own, TILE_H, axis=0), TILE_W, axis=1)[:height, :width]


def owned_pixels(width, height, rank, world):
    return int((pixel_owner(width, height, world) == rank).sum())


def combine_frames(frame, dst=0):
    """frame: torch tensor holding this rank's radiance frame (zeros outside its tiles).  After the call the
    tensor on rank `dst` holds the full frame.  One collective, no data-path communication while rendering."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(frame, dst=dst, op=dist.ReduceOp.SUM)
    return frame
