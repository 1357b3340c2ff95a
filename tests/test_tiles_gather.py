"""CPU test (gloo, world_size 2) of the multi-GPU host logic: tile sharding, size all-gather, offset prefix sum and
padded payload all-gather (lerc_b200/tiles.py).  The per-tile streams here come from the oracle (test infrastructure)
standing in for the CUDA encoder, which needs a GPU; the -m gpu tests cover the encoder itself."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _tile_blobs():
    from cases import c2_raster
    from lercapi import oracle_lib
    orc = oracle_lib()
    img = c2_raster(96, 160, seed=3)
    tiles = [img[i:i + 32, j:j + 32].copy() for i in range(0, 96, 32) for j in range(0, 160, 32)]   # 15 tiles
    blobs = []
    for t in tiles:
        st, b, _ = orc.encode(t, 0.01)
        assert st == 0
        blobs.append(b)
    return tiles, blobs


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lerc_b200.tiles import gather_streams, shard_tiles
    tiles, blobs = _tile_blobs()
    lo, hi = shard_tiles(len(blobs), rank, world)
    container, offsets = gather_streams(blobs[lo:hi], len(blobs))
    want = b"".join(blobs)
    ok = bytes(container.numpy().tobytes()) == want
    sizes = [len(b) for b in blobs]
    ok = ok and offsets.tolist() == np.concatenate([[0], np.cumsum(sizes)]).tolist()
    ret[rank] = ok
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    from lerc_b200.tiles import shard_tiles
    for n in (0, 1, 7, 15, 16, 65536):
        for world in (1, 2, 3, 8):
            r = [shard_tiles(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_gather_world2_gloo():
    from lercapi import oracle_lib
    if oracle_lib() is None:
        pytest.skip("oracle not built")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def test_gather_single_process():
    from lercapi import oracle_lib
    if oracle_lib() is None:
        pytest.skip("oracle not built")
    from lerc_b200.tiles import gather_streams
    tiles, blobs = _tile_blobs()
    container, offsets = gather_streams(blobs, len(blobs))
    assert bytes(container.numpy().tobytes()) == b"".join(blobs)
    # every tile stream decodes on its own from its offset
    orc = oracle_lib()
    for k, t in enumerate(tiles):
        b = bytes(container[offsets[k]:offsets[k + 1]].numpy().tobytes())
        st, d, _ = orc.decode(b)
        assert st == 0 and np.abs(d[0, :, :, 0].astype(np.float64) - t).max() <= 0.011
