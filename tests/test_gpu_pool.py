"""GPU tests of the shard scheduler (csrc/pool.inl) through the C-ABI: lanes of one device, several devices in one
process (skipped below 2 visible GPUs), adversarial inputs at full size against the oracle's golden
SHA-256, block-range decompression of one stream, and the device copy of the code-length allocator against the
reference's own vectors."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import HERE

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(HERE, "golden", "corpus_goldens.json")))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def pool2():
    from compressjs_flattened_b200.pool import Bzip2Pool
    p = Bzip2Pool([0], 2)
    yield p
    p.close()


@pytest.mark.parametrize("name", ["zeros1e8", "ab5e7", "rand1e8", "line97x1e6"])
@pytest.mark.parametrize("level", [9, 1])
def test_adversarial_full_size_equals_oracle_golden(gpu_engine, name, level):
    """SURVEY 8d C5b at full size: zeros(1e8) (one block eats ~46 MB of input: the cut walk, i64 offsets and -- through the
    context's pool -- halos that must grow 40-fold), ('ab') x 5e7 (periodic: ceil(log2 n) doubling rounds, tie rule),
    100 MB of random bytes, 1e6 copies of a 97-byte line.  SHA-256 of the stream == the oracle's golden; round trip."""
    from compressjs_flattened_b200.corpus import gen_adversarial
    data = gen_adversarial(name)
    g = GOLD[f"adv:{name}:L{level}"]
    assert hashlib.sha256(memoryview(data)).hexdigest() == g["input_sha256"]
    comp = gpu_engine.compressFile(data, None, level)          # >= 32 MB: through the context's two-lane pool
    assert (len(comp), hashlib.sha256(comp).hexdigest()) == (g["out_bytes"], g["out_sha256"])
    assert gpu_engine.stats().n_blocks == g["n_blocks"]
    gpu_engine.debug_set_pool(1 << 62)                         # and through the single-launch path
    try:
        comp1 = gpu_engine.compressFile(data, None, level)
    finally:
        gpu_engine.debug_set_pool()
    assert comp1 == comp
    back = gpu_engine.decompressFile(comp)
    assert hashlib.sha256(back).hexdigest() == g["input_sha256"]


def test_pool_plans_and_staging_same_stream(pool2):
    from compressjs_flattened_b200.corpus import gen_text
    data = gen_text(40_000_000, 8)
    g = None
    for first, growth, shard, staging in ((0, 0, 0, False), (3_000_000, 2.0, 0, True), (0, 0, 7_000_000, False)):
        pool2.set_plan(first, growth)
        pool2.debug(force_staging=staging)
        comp = pool2.compressFile(data, None, 9, shard_bytes=shard)
        h = hashlib.sha256(comp).hexdigest()
        g = g or h
        assert h == g
    pool2.set_plan()
    pool2.debug()
    single = __import__("compressjs_flattened_b200").Bzip2Engine(0)
    single.debug_set_pool(1 << 62)
    assert hashlib.sha256(single.compressFile(data, None, 9)).hexdigest() == g
    assert pool2.decompressFile(comp) == data.tobytes()
    single.close()


def test_pool_decompress_one_stream_in_slices(pool2, gpu_engine):
    """block-range decompression: slices of the stream over the lanes, several slice sizes incl. slices smaller
    than a block; multistream; errors are the first in stream order"""
    from compressjs_flattened_b200.bzip2 import Bzip2Error
    from compressjs_flattened_b200.corpus import gen_text
    data = gen_text(30_000_000, 8).tobytes()
    comp = gpu_engine.compressFile(data, None, 9)
    for sl in (0, 3_000_000, 500_000):
        assert pool2.decompressFile(comp, len(data), False, slice_bytes=sl) == data
    ms = comp + gpu_engine.compressFile(data[:1_000_000], None, 1)
    assert pool2.decompressFile(ms, None, True, slice_bytes=2_000_000) == data + data[:1_000_000]
    assert pool2.decompressFile(ms, None, False, slice_bytes=2_000_000) == data
    bad = bytearray(comp)
    bad[len(comp) // 3] ^= 0x40
    bad[2 * len(comp) // 3] ^= 0x40
    with pytest.raises(Bzip2Error) as e1:
        pool2.decompressFile(bytes(bad), None, False, slice_bytes=1_000_000)
    with pytest.raises(Bzip2Error) as e2:
        gpu_engine.decompressFile(bytes(bad))
    assert e1.value.errorCode == e2.value.errorCode


def test_single_context_decode_batches(gpu_engine):
    """candidates decode in batches (1280 at most); a forced batch of 7 gives the same bytes"""
    from compressjs_flattened_b200.corpus import gen_text
    data = gen_text(20_000_000, 8).tobytes()
    comp = gpu_engine.compressFile(data, None, 1)
    gpu_engine._L.bz2b200_debug_set_decode_batch(gpu_engine._ctx, None, 7)
    try:
        assert gpu_engine.decompressFile(comp) == data
    finally:
        gpu_engine._L.bz2b200_debug_set_decode_batch(gpu_engine._ctx, None, 0)
    rows = []
    gpu_engine.table(comp, lambda p, s: rows.append(s))
    assert sum(rows) == len(data) and len(rows) >= 200


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 visible GPUs")
def test_two_devices_in_one_process(gpu_engine):
    """the multi-GPU form of the reference-facing call: one process, one pool over every visible device"""
    from compressjs_flattened_b200.corpus import gen_text
    from compressjs_flattened_b200.pool import Bzip2Pool
    n = _ngpu()
    data = gen_text(200_000_000, 8)
    g = GOLD["text:200000000:8:L9"]
    pool = Bzip2Pool(list(range(n)), 1)
    try:
        comp = pool.compressFile(data, None, 9)
        assert (len(comp), hashlib.sha256(comp).hexdigest()) == (g["out_bytes"], g["out_sha256"])
        back = pool.decompressFile(comp, len(data))
        assert hashlib.sha256(back).hexdigest() == g["input_sha256"]
    finally:
        pool.close()


def test_device_allocator_matches_reference_vectors(gpu_engine, oracle):
    """VERDICT r1 parity gap: ha_allocate on the DEVICE against the reference's own allocator vectors (the oracle's copy is
    pinned by the same vectors in test_oracle_golden.py)"""
    from test_oracle_golden import HUFF_VECTORS
    for freqs, limit, expect in HUFF_VECTORS:
        got = gpu_engine.debug_huffman_lengths(freqs, limit)
        assert got == list(expect), (freqs, limit)
        assert got == oracle.huff_alloc(freqs, limit)


def test_stream_objects_on_the_gpu(gpu_engine):
    """zstream / dstream (csrc/stream_abi.inl) at real block sizes: 30 MB of text in 4 MiB chunks == compressFile; the
    decoder fed in 1 MiB pieces gives the input back; level 1; a damaged stream raises after delivering the blocks before"""
    import io
    from compressjs_flattened_b200.bzip2 import Bzip2Error
    from compressjs_flattened_b200.corpus import gen_text
    data = gen_text(30_000_000, 8).tobytes()
    whole = gpu_engine.compressFile(data, None, 9)
    assert gpu_engine.compressStream(io.BytesIO(data), None, 9, chunk_bytes=4 << 20) == whole
    assert gpu_engine.decompressStream(io.BytesIO(whole), None, False, chunk_bytes=2 << 20, piece_bytes=1 << 20) == data
    z1 = gpu_engine.compressStream(io.BytesIO(data[:5_000_000]), None, 1, chunk_bytes=1 << 20, piece_bytes=333_333)
    assert z1 == gpu_engine.compressFile(data[:5_000_000], None, 1)
    assert gpu_engine.decompressStream(io.BytesIO(whole + z1), None, True, chunk_bytes=3 << 20) == data + data[:5_000_000]
    bad = bytearray(whole)
    bad[len(whole) // 2] ^= 0x08
    sink = io.BytesIO()
    with pytest.raises(Bzip2Error):
        gpu_engine.decompressStream(io.BytesIO(bytes(bad)), sink, False, chunk_bytes=2 << 20)
    assert 0 < len(sink.getvalue()) < len(data) and data.startswith(sink.getvalue())


def _gpu_rank_main(rank, world, name, q):
    try:
        import hashlib as H
        import numpy as np
        from compressjs_flattened_b200.corpus import gen_text
        from compressjs_flattened_b200.pool import Bzip2Pool, ShardGroup, shard_plan
        n_rank, level, halo = 60_000_000, 9, 2_000_000
        total = world * n_rank
        plan = shard_plan(n_rank, level, 1)
        grp = ShardGroup(name, rank, world, timeout_ms=120_000)
        pool = Bzip2Pool([rank], 1)
        jobs, at = [], 0
        data = gen_text(total, 8)   # every rank makes the corpus; it only hands in its own shards
        for k, sz in enumerate(plan):
            base = world * at + rank * sz
            jobs.append(dict(src=np.ascontiguousarray(data[base:min(base + sz + halo, total)]), own_len=sz, base=base, index=k * world + rank))
            at += sz
        res = pool.compress_shards(grp, jobs, len(plan) * world, level)
        q.put(("segs", rank, [(j["index"], seg, (int(i.next_start), int(i.bits), int(i.n_blocks), int(i.crc_fold), int(i.complete), int(i.bit_phase)))
                              for j, (seg, i, _, _) in zip(jobs, res)]))
        stream = q_in[rank].get(timeout=300)   # the stitched stream comes back from the parent
        per = ((len(stream) + world - 1) // world + 15) & ~15
        lo, hi = min(rank * per, len(stream)), min((rank + 1) * per, len(stream))
        sl = [dict(src=np.frombuffer(stream[lo:], dtype=np.uint8), own_len=hi - lo, base=lo, index=rank)]
        part, off, nb, rc, nblk = pool.decompress_shards(grp, sl, world, len(stream), level)[0]
        q.put(("part", rank, (off, nb, rc, H.sha256(part).hexdigest(), H.sha256(data[off:off + nb].tobytes()).hexdigest())))
        pool.close()
        grp.close()
    except Exception as e:  # pragma: no cover
        q.put(("error", rank, repr(e)))


q_in = None


def _gpu_rank_entry(rank, world, name, q, qs):
    global q_in
    q_in = qs
    _gpu_rank_main(rank, world, name, q)


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 visible GPUs")
def test_two_ranks_two_gpus_one_stream(gpu_engine):
    """one process per GPU (the torchrun shape of bench.py at N > 1): the ranks exchange the per-shard scalars through the
    shared-memory group, their segments stitch into the oracle's stream (SHA golden), and the two ranks decode that ONE
    stream back, each its byte slice"""
    import multiprocessing as mp
    from compressjs_flattened_b200 import _native
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    qs = [ctx.Queue() for _ in range(world)]
    name = f"gputest_{os.getpid()}"
    procs = [ctx.Process(target=_gpu_rank_entry, args=(r, world, name, q, qs)) for r in range(world)]
    for p in procs:
        p.start()
    segs = []
    for _ in range(world):
        kind, rank, payload = q.get(timeout=600)
        assert kind == "segs", (kind, rank, payload)
        segs += payload
    segs.sort()
    stream = gpu_engine.stitch_shards(9, [s[1] for s in segs], [_native.ShardInfo(*s[2]) for s in segs])
    g = GOLD["text:120000000:8:L9"] if "text:120000000:8:L9" in GOLD else None
    if g:
        assert (len(stream), hashlib.sha256(stream).hexdigest()) == (g["out_bytes"], g["out_sha256"])
    assert gpu_engine.decompressFile(stream)[:1000] == __import__("compressjs_flattened_b200").corpus.gen_text(1000, 8).tobytes()
    for r in range(world):
        qs[r].put(stream)
    total = 0
    for _ in range(world):
        kind, rank, payload = q.get(timeout=600)
        assert kind == "part", (kind, rank, payload)
        off, nb, rc, got, exp = payload
        assert rc == 0 and got == exp
        total += nb
    assert total == world * 60_000_000
    for p in procs:
        p.join(timeout=120)
