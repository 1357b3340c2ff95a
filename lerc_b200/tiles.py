"""Tile sharding and stream gathering for tiled rasters coded on several GPUs (BASELINE config 5: a 65536 x 65536
raster as 256 x 256 tiles, every tile its own standard Lerc2 blob).

Tiles are independent objects, so the codec needs no data-path collective: rank r codes the contiguous tile range
shard_tiles(n_tiles, r, world).  Only the finished streams are exchanged: one fixed-size all-gather of the per-tile byte
counts, an exclusive prefix sum for the global offsets, and the payloads straight into their final place in the container
(equal chunks of the compact container through one all_gather_into_tensor, the misfit at the chunk ends point to point: NCCL
has no all-gather-v).  Works with any torch.distributed backend (nccl on GPUs, gloo
in tests/test_tiles_gather.py).
"""
import torch
import torch.distributed as dist


def shard_tiles(n_tiles, rank, world):
    """contiguous, ordered tile range [lo, hi) of `rank`; the first n_tiles % world ranks hold one tile more"""
    base, extra = divmod(n_tiles, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def offsets_from_sizes(sizes):
    """exclusive prefix sum: byte offset of every tile stream in the concatenated container (+ total as last entry)"""
    sizes = torch.as_tensor(sizes, dtype=torch.int64)
    out = torch.zeros(sizes.numel() + 1, dtype=torch.int64, device=sizes.device)
    out[1:] = torch.cumsum(sizes, 0)
    return out


def gather_container(local, local_offsets, n_tiles, group=None, out=None):
    """local: uint8 tensor with this rank's tile streams back to back in tile order (what lerc_b200_encodeTiles writes for the
    rank's tile range); local_offsets: its n_local + 1 byte offsets.  Returns (container with all n_tiles streams back to back
    in global tile order, int64 offsets[n_tiles + 1]) on every rank.

    One fixed-size all-gather of the per-tile byte counts and an exclusive prefix sum give every rank the global offsets and
    the byte range of every rank's streams; the compact container is then cut into `world` EQUAL chunks: the few bytes of a rank's
    streams that fall outside its own chunk travel point to point, and one all_gather_into_tensor of the equal chunks writes
    every rank's streams straight to their final place (NCCL has no all-gather-v).  Nothing is padded, nothing is copied a
    second time.  `out`: optional preallocated uint8 tensor for the container (at least the total size)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_tiles(n_tiles, rank, world)
    device = local.device
    off = torch.as_tensor(local_offsets).to(device=device, dtype=torch.int64)
    assert off.numel() == hi - lo + 1, (off.numel(), lo, hi)
    max_tiles = -(-n_tiles // world)
    my_sizes = torch.zeros(max_tiles, dtype=torch.int64, device=device)
    my_sizes[: hi - lo] = off[1:] - off[:-1]
    if world == 1:
        all_sizes = my_sizes[None, :]
    else:
        parts = [torch.empty_like(my_sizes) for _ in range(world)]           # (a few KB: the list form works with every backend)
        dist.all_gather(parts, my_sizes, group=group)
        all_sizes = torch.stack(parts)
    # per-tile sizes in global tile order
    counts = [shard_tiles(n_tiles, r, world)[1] - shard_tiles(n_tiles, r, world)[0] for r in range(world)]
    sizes = torch.cat([all_sizes[r, : counts[r]] for r in range(world)])
    offsets = offsets_from_sizes(sizes)
    rank_bytes = all_sizes.sum(dim=1).tolist()                    # (one small device-to-host read: the byte range of every rank)
    assert rank_bytes[rank] <= local.numel()
    total = int(sum(rank_bytes))
    starts = [0]
    for r in range(world):
        starts.append(starts[-1] + rank_bytes[r])
    if world == 1:
        container = out[:total] if out is not None else torch.empty(total, dtype=torch.uint8, device=device)
        container.copy_(local[:total])
        return container, offsets
    # The compact container is cut into `world` equal chunks of C bytes.  Rank r's own streams [starts[r], starts[r + 1]) are chunk r
    # up to the small misfit at its two ends (the ranks' byte counts differ by a percent): those end pieces go point to point to the
    # rank whose chunk they belong to, then ONE all_gather_into_tensor of the equal chunks lands every byte at its final place --
    # the collective runs at all-gather bandwidth (all NVLink channels; a group of per-peer send/recv pairs measured 240 GB/s, a third of it).
    C_ = -(-total // world)
    C_ = (C_ + 15) // 16 * 16
    full = out if (out is not None and out.numel() >= world * C_) else torch.empty(world * C_, dtype=torch.uint8, device=device)
    chunk = full[rank * C_:(rank + 1) * C_]
    ops = []
    lo_b, hi_b = starts[rank], starts[rank + 1]
    for q in range(world):
        c0, c1 = q * C_, min((q + 1) * C_, total)
        # my bytes that belong to chunk q
        a0, a1 = max(lo_b, c0), min(hi_b, c1)
        if a1 > a0:
            piece = local[a0 - lo_b:a1 - lo_b]
            if q == rank:
                chunk[a0 - c0:a1 - c0].copy_(piece)
            else:
                ops.append(dist.P2POp(dist.isend, piece.contiguous(), dist.get_global_rank(group, q) if group is not None else q, group))
        # rank q's bytes that belong to my chunk
        if q != rank:
            m0, m1 = rank * C_, min((rank + 1) * C_, total)
            b0, b1 = max(starts[q], m0), min(starts[q + 1], m1)
            if b1 > b0:
                ops.append(dist.P2POp(dist.irecv, chunk[b0 - m0:b1 - m0], dist.get_global_rank(group, q) if group is not None else q, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    in_place_ok = dist.get_backend(group) == "nccl"                    # (NCCL's in-place convention: the input is the rank's slot of the output)
    dist.all_gather_into_tensor(full[:world * C_], chunk if in_place_ok else chunk.clone(), group=group)
    container = full[:total]
    return container, offsets


def gather_streams(local_blobs, n_tiles, group=None, device=None):
    """local_blobs: the tile streams (bytes / uint8 tensors) of this rank's shard, in tile order.
    Returns (container uint8 tensor with all n_tiles streams back to back in tile order, int64 offsets[n_tiles + 1])."""
    if device is None:
        device = local_blobs[0].device if (local_blobs and torch.is_tensor(local_blobs[0])) else torch.device("cpu")
    blobs = [b if torch.is_tensor(b) else torch.frombuffer(bytearray(b), dtype=torch.uint8) for b in local_blobs]
    blobs = [b.to(device) for b in blobs]
    local = torch.cat(blobs) if blobs else torch.zeros(0, dtype=torch.uint8, device=device)
    off = offsets_from_sizes(torch.tensor([b.numel() for b in blobs], dtype=torch.int64, device=device))
    return gather_container(local, off, n_tiles, group=group)
