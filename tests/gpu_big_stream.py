"""BASELINE.json configs 4 + 5a on N ranks (torchrun): an N x MB corpus (seed 64) is compressed as ONE stream by all ranks
(interleaved block-range shards, scalars through a shared-memory group), stitched, compared with the oracle's golden
SHA-256 (tests/golden/corpus_goldens.json, `text:<bytes>:64:L9`), and decoded back as ONE stream by all ranks (byte slices,
walk state chained through the group; every block CRC and the stream CRC are verified by the decoder, the total length and
a sample of every rank's part are compared with the corpus).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tests/gpu_big_stream.py --mb 1000
Prints one JSON line on rank 0."""
import argparse
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from compressjs_flattened_b200 import _native  # noqa: E402
from compressjs_flattened_b200.corpus import CHUNK, gen_text  # noqa: E402
from compressjs_flattened_b200.pool import Bzip2Pool, ShardGroup, shard_plan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=1000, help="MB of corpus per rank")
ap.add_argument("--seed", type=int, default=64)
ap.add_argument("--level", type=int, default=9)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctl = dist.new_group(backend="gloo")
GOLD = json.load(open(os.path.join(HERE, "golden", "corpus_goldens.json")))
nbytes, level = args.mb * 1_000_000, args.level
total = world * nbytes
tag = f"{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
shm = f"/dev/shm/bz2b200_big_{tag}"


def gen_range(lo, hi):
    c0, c1 = lo // CHUNK, (hi + CHUNK - 1) // CHUNK
    buf = gen_text((c1 - c0) * CHUNK, args.seed, first_chunk=c0, workers=4)
    return buf[lo - c0 * CHUNK: hi - c0 * CHUNK]


def barrier():
    torch.cuda.synchronize()
    dist.barrier()


t_gen = time.time()
plan = shard_plan(nbytes, level, 1)
halo = 2_000_000
jobs, keep, at = [], [], 0
for k, sz in enumerate(plan):
    base = world * at + rank * sz
    buf = torch.from_numpy(np.ascontiguousarray(gen_range(base, min(base + sz + halo, total)))).pin_memory()
    keep.append(buf)
    jobs.append(dict(src=buf.numpy(), own_len=sz, base=base, index=k * world + rank))
    at += sz
nshards = len(plan) * world
t_gen = time.time() - t_gen
pool = Bzip2Pool([local], 1)
grp = ShardGroup(f"big_{tag}", rank, world, timeout_ms=600_000)
for seg, info, off, nb in pool.compress_shards(grp, jobs, nshards, level, to_bytes=False):   # warm-up: buffers, page-locked result memory
    if seg:
        pool.free_raw(seg)
barrier()
t0 = time.time()
for _ in range(args.reps):
    res = pool.compress_shards(grp, jobs, nshards, level, to_bytes=False)
    if _ + 1 < args.reps:
        for seg, info, off, nb in res:
            if seg:
                pool.free_raw(seg)
barrier()
comp_ms = (time.time() - t0) * 1e3 / args.reps
# ---- one stream: segments to shared memory, rank 0 stitches ----
import ctypes as C  # noqa: E402
meta = []
for j, (seg, info, off, nb) in zip(jobs, res):
    path = f"{shm}_seg{j['index']}"
    with open(path, "wb") as f:
        if nb:
            f.write(_native.take_bytes(seg, nb))
    if seg:
        pool.free_raw(seg)
    meta.append((j["index"], (int(info.next_start), int(info.bits), int(info.n_blocks), int(info.crc_fold), int(info.complete), int(info.bit_phase)), off))
allmeta = [None] * world if rank == 0 else None
dist.gather_object(meta, allmeta, dst=0, group=ctl)
out = {}
nw = torch.zeros(1, dtype=torch.int64, device=dev)
if rank == 0:
    parts = sorted(x for m in allmeta for x in m)
    from compressjs_flattened_b200.bzip2 import Bzip2Engine
    eng = Bzip2Engine(local)
    segs = [open(f"{shm}_seg{p[0]}", "rb").read() for p in parts]
    whole = eng.stitch_shards(level, segs, [_native.ShardInfo(*p[1]) for p in parts])
    del segs
    sha = hashlib.sha256(whole).hexdigest()
    key = f"text:{total}:{args.seed}:L{level}"
    g = GOLD.get(key)
    out.update({"stream_bytes": len(whole), "stream_bits": len(whole) * 8, "bit_offsets_above_2^32": len(whole) * 8 > 1 << 32, "shards": len(parts),
                "blocks": sum(p[1][2] for p in parts), "golden": key if g else None,
                "sha256_equals_oracle_golden": (sha == g["out_sha256"] and len(whole) == g["out_bytes"] and sum(p[1][2] for p in parts) == g["n_blocks"]) if g else None,
                "sha256": sha})
    with open(f"{shm}_stream", "wb") as f:
        f.write(whole)
    nw[0] = len(whole)
    del whole
dist.broadcast(nw, 0)
nw = int(nw.item())
for j in jobs:
    try:
        os.unlink(f"{shm}_seg{j['index']}")
    except OSError:
        pass
# ---- decode the ONE stream on all ranks: byte slices of the stream, a few per rank ----
per = min(((nw + world - 1) // world + 4095) & ~4095, 352 << 20)   # one slice per rank while it fits one decode batch (~1200 blocks)
nsl = (nw + per - 1) // per
mine = [s for s in range(nsl) if s % world == rank]
mm = np.memmap(f"{shm}_stream", dtype=np.uint8, mode="r")
dhalo = 4 << 20
djobs, dkeep = [], []
for s in mine:
    lo, hi = s * per, min((s + 1) * per, nw)
    buf = torch.from_numpy(np.ascontiguousarray(mm[lo:min(hi + dhalo, nw)])).pin_memory()
    dkeep.append(buf)
    djobs.append(dict(src=buf.numpy(), own_len=hi - lo, base=lo, index=s))
r0 = pool.decompress_shards(grp, djobs, nsl, nw, level, to_bytes=False)   # warm-up and verification pass
ok, nbytes_dec, nblk = 1, 0, 0
for part, off, nb, rc, blocks in r0:
    ok &= int(rc == 0)
    nbytes_dec += nb
    nblk += blocks
    if nb:
        k = min(nb, 1_000_000)
        got = np.frombuffer(C.string_at(part, k), dtype=np.uint8)
        ok &= int(np.array_equal(got, gen_range(off, off + k)))
        tail = np.frombuffer(C.string_at(C.addressof(part.contents) + nb - k, k), dtype=np.uint8)
        ok &= int(np.array_equal(tail, gen_range(off + nb - k, off + nb)))
    if part:
        pool.free_raw(part)
barrier()
t1 = time.time()
for _ in range(args.reps):
    r1 = pool.decompress_shards(grp, djobs, nsl, nw, level, to_bytes=False)
    for part, off, nb, rc, blocks in r1:
        if part:
            pool.free_raw(part)
barrier()
dec_ms = (time.time() - t1) * 1e3 / args.reps
t = torch.tensor([ok, nbytes_dec, nblk], dtype=torch.int64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.SUM)
tm = torch.tensor([comp_ms, dec_ms, t_gen], dtype=torch.float64, device=dev)
dist.all_reduce(tm, op=dist.ReduceOp.MAX)
if rank == 0:
    os.unlink(f"{shm}_stream")
    out.update({"config": f"BASELINE configs 4+5a: {total} B synthetic enwik8-like text (seed {args.seed}), level {level}, {world} x B200, one .bz2",
                "n_gpus": world, "plan_mb_per_rank": [round(x / 1e6, 1) for x in plan],
                "compress": {"MBps": round(total / 1e6 / (tm[0].item() / 1e3), 1), "ms": round(tm[0].item(), 1),
                             "api": "bz2b200_pool_compress_shards over a shared-memory group, pinned host shards in, page-locked host segments out"},
                "decompress": {"MBps": round(total / 1e6 / (tm[1].item() / 1e3), 1), "ms": round(tm[1].item(), 1), "slices": int(nsl),
                               "api": "bz2b200_pool_decompress_shards: ONE stream, byte slices round-robin over the ranks, host slices in, page-locked host parts out",
                               "decoded_bytes": int(t[1].item()), "blocks": int(t[2].item()),
                               "all_crcs_verified_and_samples_equal": bool(t[0].item() == world and t[1].item() == total)},
                "corpus_generation_s": round(tm[2].item(), 1)})
    print(json.dumps(out), flush=True)
pool.close()
grp.close()
dist.destroy_process_group()
