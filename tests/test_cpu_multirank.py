"""CPU tests of the N>1 path (world_size 2, gloo): tile ownership and the single exchange step.  The per-rank
"render" here is the oracle restricted to the rank's pixels, so the host logic (partition + combine) is what
is under test; the device side of the same partition is covered by the GPU suite (tile-sharded renders)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dtb200 import tiles
from oracle_util import oracle_render
from scenes_util import golden_scene


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hs, _ = golden_scene("simple")
    cam = hs.camera(0)
    cam.width, cam.height = 100, 52                       # not a multiple of the 8x4 tile
    _, hdr, _ = oracle_render(hs, cam, threads=1)
    mine = tiles.pixel_owner(cam.width, cam.height, world) == rank
    part = np.where(mine[..., None], hdr, 0.0).astype(np.float32)
    t = torch.from_numpy(part.copy())
    tiles.combine_frames(t, dst=0)
    if rank == 0:
        np.save(out_path, np.stack([t.numpy(), hdr]))
    dist.barrier()
    dist.destroy_process_group()


def test_tile_partition_is_a_partition():
    for (w, h, world) in ((100, 52, 2), (1920, 1080, 8), (7, 3, 4), (8, 4, 3)):
        own = tiles.pixel_owner(w, h, world)
        assert own.shape == (h, w) and own.min() >= 0 and own.max() < world
        assert sum(tiles.owned_pixels(w, h, r, world) for r in range(world)) == w * h
        if w * h > 10000:                                   # round-robin tiles balance the load
            counts = [tiles.owned_pixels(w, h, r, world) for r in range(world)]
            assert max(counts) - min(counts) <= 2 * tiles.TILE_W * tiles.TILE_H * (h // tiles.TILE_H + 1)


def test_two_rank_gloo_combine(tmp_path):
    out = str(tmp_path / "combined.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    combined, full = np.load(out)
    assert np.array_equal(combined, full)                 # SUM of disjoint shards == the unsharded frame, bit for bit


def test_tiles_are_dealt_out_in_row_aligned_strips():
    """dt_rank_tile (csrc/dt_kernels.cuh) / tiles.tile_owner: strips of TILE_GROUP tiles never wrap around a tile row, every
    strip has one owner, and consecutive strips (row-major) go to consecutive ranks."""
    for (w, h, world) in ((1920, 1080, 8), (2712, 1528, 2), (100, 52, 3), (64, 4, 4)):
        own = tiles.tile_owner(w, h, world)
        ty, tx = own.shape
        strips_x = (tx + tiles.TILE_GROUP - 1) // tiles.TILE_GROUP
        for y in range(0, ty, max(1, ty // 7)):
            for sx in range(strips_x):
                seg = own[y, sx * tiles.TILE_GROUP:(sx + 1) * tiles.TILE_GROUP]
                assert (seg == seg[0]).all()
                assert seg[0] == (y * strips_x + sx) % world
