"""The shard scheduler (csrc/pool.inl) on the CPU simulator: lanes of one process, forced staging, halo growth, shards
that own no block, and two PROCESSES exchanging the per-shard scalars through a shared-memory group -- every stream
byte-identical to the oracle (and to the single-call path)."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from conftest import HERE

SIM_SO = os.path.join(HERE, "sim", "libbz2b200_sim.so")


def _runny(rng, n, nsym, plong):
    out = bytearray()
    while len(out) < n:
        b = int(rng.integers(0, nsym))
        ln = int(rng.integers(4, 700)) if rng.random() < plong else int(rng.integers(1, 4))
        out += bytes([b]) * ln
    return bytes(out[:n])


def _inputs():
    rng = np.random.default_rng(21)
    from compressjs_flattened_b200.corpus import gen_text
    return {
        "text": gen_text(60_000, 4).tobytes(),
        "runny": _runny(rng, 50_000, 3, 0.4),        # blocks that read far more input than they hold: halos must grow
        "zeros": bytes(40_000),
        "rand": rng.integers(0, 256, 30_000, dtype=np.uint8).tobytes(),
        "empty": b"",
        "tiny": b"abc",
    }


@pytest.fixture(scope="module")
def sim_lib(sim_engine):
    from compressjs_flattened_b200 import _native
    return _native.Library(SIM_SO)


@pytest.mark.parametrize("name", sorted(_inputs()))
def test_context_pool_equals_oracle(sim_lib, oracle, name):
    """bz2b200_compress routed through the context's own two-lane pool (threshold lowered for the test)."""
    from compressjs_flattened_b200.bzip2 import Bzip2Engine
    data = _inputs()[name]
    eng = Bzip2Engine(0, sim_lib)
    eng.debug_set_block_cap(997)
    oracle.set_block_cap(997)
    try:
        exp, st = oracle.compress(data, 9, return_stats=True)
        for shard, halo, staging in ((4096, 0, False), (8192, 64, True), (1000, 16, False)):
            eng.debug_set_pool(0, shard, halo, staging)
            assert eng.compressFile(data, None, 9) == exp, (shard, halo, staging)
            assert eng.stats().n_blocks == st.n_blocks
    finally:
        oracle.set_block_cap(0)
        eng.close()


def test_pool_lanes_and_devices(sim_lib, oracle):
    """an explicit pool: 2 'devices' x 2 lanes, several shard sizes, levels 1 and 9 at the real block sizes"""
    from compressjs_flattened_b200.pool import Bzip2Pool
    from compressjs_flattened_b200.corpus import gen_text
    data = gen_text(450_000, 7).tobytes()
    pool = Bzip2Pool([0, 0], 2, library=sim_lib)
    try:
        for level, shard in ((1, 0), (1, 120_000), (9, 200_000)):
            assert pool.compressFile(data, None, level, shard_bytes=shard) == oracle.compress(data, level)
        with pytest.raises(ValueError):
            pool.compressFile(data, None, 0)
    finally:
        pool.close()


def _rank_main(rank, world, name, data, shard, cap, q):
    try:
        from compressjs_flattened_b200 import _native
        from compressjs_flattened_b200.pool import Bzip2Pool, ShardGroup
        lib = _native.Library(SIM_SO)
        grp = ShardGroup(name, rank, world, timeout_ms=60_000, library=lib)
        pool = Bzip2Pool([0], 2, library=lib)
        pool.debug(block_cap=cap, first_halo=32)
        total = (len(data) + shard - 1) // shard
        outs = []
        for rep in range(2):      # two collective calls on one group: epochs must not mix
            jobs = [dict(src=np.frombuffer(data[j * shard:], dtype=np.uint8), own_len=min(shard, len(data) - j * shard), base=j * shard, index=j)
                    for j in range(total) if j % world == rank]
            res = pool.compress_shards(grp, jobs, total, 9)
            outs.append([(j["index"], seg, (int(i.next_start), int(i.bits), int(i.n_blocks), int(i.crc_fold), int(i.complete), int(i.bit_phase)))
                         for j, (seg, i, _, _) in zip(jobs, res)])
        q.put((rank, outs))
        pool.close()
        grp.close()
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))


@pytest.mark.parametrize("world", [2, 3])
def test_group_of_processes_equals_oracle(sim_lib, oracle, world):
    from compressjs_flattened_b200 import _native
    from compressjs_flattened_b200.bzip2 import Bzip2Engine
    rng = np.random.default_rng(5)
    data = _runny(rng, 30_000, 4, 0.2) + bytes(rng.integers(0, 256, 20_000, dtype=np.uint8))
    shard, cap = 3000, 701
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = f"test_{os.getpid()}_{world}"
    procs = [ctx.Process(target=_rank_main, args=(r, world, name, data, shard, cap, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r, v in got.items():
        assert not isinstance(v, str), f"rank {r}: {v}"
    oracle.set_block_cap(cap)
    try:
        exp = oracle.compress(data, 9)
    finally:
        oracle.set_block_cap(0)
    eng = Bzip2Engine(0, sim_lib)
    for rep in range(2):
        parts = sorted(x for r in range(world) for x in got[r][rep])
        segs = [p[1] for p in parts]
        infos = [_native.ShardInfo(*p[2]) for p in parts]
        assert eng.stitch_shards(9, segs, infos) == exp


def _dead_peer_main(name, q):
    from compressjs_flattened_b200 import _native
    from compressjs_flattened_b200.pool import Bzip2Pool, ShardGroup
    lib = _native.Library(SIM_SO)
    grp = ShardGroup(name, 1, 2, timeout_ms=8_000, library=lib)
    pool = Bzip2Pool([0], 1, library=lib)
    data = np.frombuffer(b"x" * 5000, dtype=np.uint8)
    try:   # shard 0 never arrives: rank 0 joined the group and left
        pool.compress_shards(grp, [dict(src=data[2500:], own_len=2500, base=2500, index=1)], 2, 9)
        q.put("no error")
    except RuntimeError as e:
        q.put(str(e))


def test_dead_peer_is_an_error_not_a_hang(sim_lib):
    """Robustness of the shard protocol: a peer that never publishes makes the wait time out."""
    from compressjs_flattened_b200.pool import ShardGroup
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = f"dead_{os.getpid()}"
    p = ctx.Process(target=_dead_peer_main, args=(name, q))
    p.start()
    grp = ShardGroup(name, 0, 2, timeout_ms=8_000, library=sim_lib)   # attaches, then does nothing
    msg = q.get(timeout=120)
    p.join(timeout=30)
    grp.close()
    assert "peer" in msg
