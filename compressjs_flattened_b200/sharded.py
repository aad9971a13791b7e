"""Block-range sharding of one bzip2 stream over the ranks of a torch.distributed group.

Each rank holds `buf` = its slice of the input followed by a halo (the bytes after the slice that its last block may
need).  Cross-rank traffic is three scalars per rank -- no collective on the data path (they travel through the
process group, or through a shared-memory HostMailbox when all ranks are on one node):
  1. chain   : global offset of the first block of rank r, sent by rank r-1 after its cut walk
  2. exscan  : bit lengths of the segments (all_gather of one int64), giving each segment's bit offset
  3. fold    : (n_blocks, crc_fold) of each segment, for the combined CRC (rotations compose)
The result is byte-identical to compressing the whole input on one GPU.
"""
import time

import torch
import torch.distributed as dist

from . import _native


class HostMailbox:
    """Scalars between the ranks of ONE node through a page of shared memory (/dev/shm): a slot per rank holding
    (sequence number, value) pairs; a reader spins on the sequence number.  A few microseconds per hop, against
    a millisecond for a TCP-backed process group -- the chain of first-block offsets is N-1 hops long."""
    FIELDS = 6  # (chain, bits, G) x 2 alternating slots

    def __init__(self, rank, world, key, timeout_s=120.0):
        import numpy as np
        self.rank, self.world, self.seq, self.timeout_s = rank, world, 0, timeout_s
        self.path = f"/dev/shm/bz2b200_mailbox_{key}"
        shape = (world, self.FIELDS * 2)
        if rank == 0:
            m = np.lib.format.open_memmap(self.path, mode="w+", dtype=np.int64, shape=shape)
            m[:] = 0
            m.flush()
        if world > 1:
            dist.barrier()
        self.m = np.lib.format.open_memmap(self.path, mode="r+") if rank else m

    def next_round(self):
        self.seq += 1
        return self.seq

    # Two slots per field, used alternately: a rank cannot finish round k+1 before every rank has published its
    # round-k+1 bit length, which each does only after reading everything of round k -- so a slot is never
    # overwritten while a reader of the round before last still needs it.
    def put(self, field, value):
        f = 2 * field + (self.seq & 1)
        self.m[self.rank, 2 * f + 1] = value
        self.m[self.rank, 2 * f] = self.seq

    def get(self, src, field):
        f = 2 * field + (self.seq & 1)
        row = self.m[src]
        t0, spins = time.time(), 0
        while int(row[2 * f]) != self.seq:
            time.sleep(0)   # let other threads of this process have the interpreter
            spins += 1
            if spins & 0xfff == 0 and time.time() - t0 > self.timeout_s:   # a peer that died or never joined: an error, not a hang
                raise TimeoutError(f"shard mailbox: rank {src} did not publish field {field} of round {self.seq} within {self.timeout_s} s")
        return int(row[2 * f + 1])

    def close(self):
        import os
        if self.world > 1:
            dist.barrier()
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def compress_shard(engine, buf, base, own_len, level, is_last, rank=None, world=None, device=None, group=None,
                   device_ptr=None, nbytes=None, to_host=True, mailbox=None):
    """Compress the blocks that start inside [base, base+own_len) of the global input.

    buf: bytes-like slice+halo (host) -- or pass device_ptr/nbytes for data already in HBM.
    mailbox: a HostMailbox (single node) carries the scalars instead of the process group.
    to_host: True -> bytes, False -> the segment stays in HBM (length returned), "raw" -> (pointer, length) owned
    by the caller (engine.free_raw).
    Returns (segment, ShardInfo, bit_offset_of_segment_in_stream)."""
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    dev = device or torch.device("cpu")
    engine.shard_begin(buf, level, device_ptr=device_ptr, nbytes=nbytes)   # summaries: no dependency on other ranks
    if mailbox is not None:
        # All ranks cut AT ONCE from speculated starts (blocks begin where G reaches a multiple of B, give or take a few
        # units for every cut that fell inside a run upstream); the chain of true first-block offsets then only has to
        # be compared, one scalar hop per rank.
        mailbox.next_round()
        mailbox.put(2, engine.shard_gtotal(own_len))                        # (0) what this shard adds to G
        g_before = sum(mailbox.get(r, 2) for r in range(rank))
        engine.shard_cut_g(g_before, own_len)                               # 64 phases around the expected one, at once
        start_v = mailbox.get(rank - 1, 0) if rank > 0 else 0               # (1) first-block offset, global coordinates
        s_local = max(start_v - base, 0)
        info = engine.shard_cut_pick(s_local, own_len, is_last)
        compress_shard.guesses[info is not None] += 1
        if info is None:
            info = engine.shard_cut(s_local, own_len, is_last)              # no speculated walk started there: cut now
    else:
        start = torch.zeros(1, dtype=torch.int64, device=dev)
        if rank > 0:
            dist.recv(start, src=rank - 1, group=group)
        start_v = int(start.item())
        s_local = max(start_v - base, 0)
        info = engine.shard_cut(s_local, own_len, is_last)
    if not info.complete:
        raise RuntimeError("halo too short: the last owned block needs input beyond the buffer")
    nxt_v = max(base + int(info.next_start), start_v) if info.n_blocks else start_v
    if mailbox is not None:
        mailbox.put(0, nxt_v)
    elif rank + 1 < world:
        dist.send(torch.tensor([nxt_v], dtype=torch.int64, device=dev), dst=rank + 1, group=group)
    engine.shard_compress(info)
    if mailbox is not None:
        mailbox.put(1, int(info.bits))                                       # (2) exclusive scan of bit lengths
        allb = [mailbox.get(r, 1) for r in range(world)]                     # everybody reads everybody: keeps the ranks within a round
        bit_off = 32 + sum(allb[:rank])
    else:
        mine = torch.tensor([int(info.bits)], dtype=torch.int64, device=dev)
        allbits = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        if world > 1:
            dist.all_gather(allbits, mine, group=group)
        else:
            allbits = [mine]
        bit_off = 32 + sum(int(b.item()) for b in allbits[:rank])
    seg = engine.shard_emit(info, bit_off & 7, to_host=to_host)
    return seg, info, bit_off


compress_shard.guesses = {True: 0, False: 0}   # speculative starts that held / had to be recut (this process)


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
