"""Host-side logic of the multi-GPU path: which 8x4-pixel ray tiles a rank owns, and the one
exchange step that combines the per-rank radiance frames on rank 0.

The device side mirrors `tile_owner` exactly (k_generate: tile = local_index * tile_world + tile_rank).  The
combine step is a single reduce(SUM): every pixel is non-zero on exactly one rank, so the sum is a gather.
`combine_frames` works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np

TILE_W, TILE_H = 8, 4


def tile_grid(width, height):
    return (width + TILE_W - 1) // TILE_W, (height + TILE_H - 1) // TILE_H


def tile_owner(width, height, world):
    """[tiles_y, tiles_x] array of owning ranks (round-robin over the row-major tile index)."""
    tx, ty = tile_grid(width, height)
    return (np.arange(tx * ty, dtype=np.int64) % world).reshape(ty, tx)


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
