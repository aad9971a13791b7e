"""Block-range sharding of one bzip2 stream over the ranks of a torch.distributed group.

Each rank holds `buf` = its slice of the input followed by a halo (the bytes after the slice that its last block may
need).  Cross-rank traffic is three scalars per rank -- no collective on the data path:
  1. chain   : global offset of the first block of rank r, sent by rank r-1 after its cut walk
  2. exscan  : bit lengths of the segments (all_gather of one int64), giving each segment's bit offset
  3. fold    : (n_blocks, crc_fold) of each segment, for the combined CRC (rotations compose)
The result is byte-identical to compressing the whole input on one GPU.
"""
import torch
import torch.distributed as dist

from . import _native


def compress_shard(engine, buf, base, own_len, level, is_last, rank=None, world=None, device=None, group=None,
                   device_ptr=None, nbytes=None, to_host=True):
    """Compress the blocks that start inside [base, base+own_len) of the global input.

    buf: bytes-like slice+halo (host) -- or pass device_ptr/nbytes for data already in HBM.
    Returns (segment_bytes, ShardInfo, bit_offset_of_segment_in_stream)."""
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    dev = device or torch.device("cpu")
    engine.shard_begin(buf, level, device_ptr=device_ptr, nbytes=nbytes)   # summaries: no dependency on other ranks
    start = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank > 0:
        dist.recv(start, src=rank - 1, group=group)                       # (1) first-block offset, global coordinates
    s_local = max(int(start.item()) - base, 0)
    info = engine.shard_cut(s_local, own_len, is_last)
    if not info.complete:
        raise RuntimeError("halo too short: the last owned block needs input beyond the buffer")
    if rank + 1 < world:
        nxt = torch.tensor([max(base + int(info.next_start), int(start.item()))], dtype=torch.int64, device=dev)
        dist.send(nxt, dst=rank + 1, group=group)
    engine.shard_compress(info)
    mine = torch.tensor([int(info.bits)], dtype=torch.int64, device=dev)
    allbits = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    if world > 1:
        dist.all_gather(allbits, mine, group=group)                       # (2) exclusive scan of bit lengths
    else:
        allbits = [mine]
    bit_off = 32 + sum(int(b.item()) for b in allbits[:rank])
    seg = engine.shard_emit(info, bit_off & 7, to_host=to_host)
    return seg, info, bit_off


def gather_and_stitch(engine, seg, info, level, group=None):
    """Rank 0 assembles the stream (bytes copies + one OR per boundary); other ranks return None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    payload = (bytes(seg), (int(info.next_start), int(info.bits), int(info.n_blocks), int(info.crc_fold), int(info.complete), int(info.bit_phase)))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0, group=group)
    if rank != 0:
        return None
    segs = [g[0] for g in gathered]
    infos = [_native.ShardInfo(*g[1]) for g in gathered]
    return engine.stitch_shards(level, segs, infos)
