"""World-size-2/3 gloo tests of the block-range shard protocol (compressjs_flattened_b200/sharded.py) on CPU.
The kernels run on the CPU simulator (tests/sim); with a tiny block capacity every rank owns dozens of blocks and
cuts fall inside runs that straddle the slice boundary.  The stitched stream must equal the oracle's."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import HERE, ROOT

WORKER = textwrap.dedent("""
    import os, sys, numpy as np
    sys.path.insert(0, {root!r}); sys.path.insert(0, {here!r})
    import torch.distributed as dist
    import oracle_binding as O
    from compressjs_flattened_b200 import _native
    from compressjs_flattened_b200.bzip2 import Bzip2Engine
    from compressjs_flattened_b200.sharded import compress_shard, gather_and_stitch
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = Bzip2Engine(0, _native.Library(os.path.join({here!r}, "sim", "libbz2b200_sim.so")))
    cap, level = {cap}, 9
    rng = np.random.default_rng(123)
    parts = []
    for _ in range(1500):
        parts.append(bytes([int(rng.integers(0, 4))]) * int(rng.choice([1, 1, 1, 2, 3, 4, 5, 9, 255, 256, 300, 700])))
    data = b"".join(parts)[:{nbytes}]
    eng.debug_set_block_cap(cap)
    n = len(data)
    slice_len = (n + world - 1) // world
    base = rank * slice_len
    own = max(0, min(slice_len, n - base))
    halo = {halo}
    buf = data[base:min(n, base + own + halo)]
    seg, info, off = compress_shard(eng, buf, base, own, level, is_last=(rank == world - 1))
    out = gather_and_stitch(eng, seg, info, level)
    if rank == 0:
        O.set_block_cap(cap)
        exp = O.compress(data, level)
        assert out == exp, (len(out), len(exp))
        assert O.decompress(out) == data
        print("STITCH_OK", len(out))
    dist.destroy_process_group()
""")


@pytest.mark.parametrize("world,cap,halo", [(2, 64, 20000), (3, 301, 20000)])
def test_sharded_stream_equals_oracle(tmp_path, sim_engine, oracle, world, cap, halo):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, here=HERE, cap=cap, nbytes=30000, halo=halo))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29500 + world), str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "STITCH_OK" in r.stdout


def test_single_rank_shards_equal_whole_stream(sim_engine, oracle):
    """The same protocol driven in-process: 4 shards of one stream, compressed one after the other."""
    import numpy as np
    from compressjs_flattened_b200 import _native
    rng = np.random.default_rng(9)
    data = bytes(np.repeat(rng.integers(0, 3, 3000, dtype=np.uint8), rng.choice([1, 1, 2, 4, 7, 260], 3000)))
    cap = 200
    try:
        oracle.set_block_cap(cap)
        sim_engine.debug_set_block_cap(cap)
        n, world = len(data), 4
        slice_len = (n + world - 1) // world
        start, bitpos, segs, infos = 0, 32, [], []
        for r in range(world):
            base = r * slice_len
            own = max(0, min(slice_len, n - base))
            sim_engine.shard_begin(data[base:min(n, base + own + 5000)], 9)
            info = sim_engine.shard_cut(max(start - base, 0), own, r == world - 1)
            assert info.complete
            start = max(base + info.next_start, start)
            sim_engine.shard_compress(info)
            segs.append(sim_engine.shard_emit(info, bitpos & 7))
            infos.append(info)
            bitpos += info.bits
        out = sim_engine.stitch_shards(9, segs, infos)
        assert out == oracle.compress(data, 9) == sim_engine.compressFile(data, None, 9)
    finally:
        oracle.set_block_cap(0)
        sim_engine.debug_set_block_cap(0)


MAILBOX_WORKER = textwrap.dedent("""
    import os, sys, numpy as np
    sys.path.insert(0, {root!r}); sys.path.insert(0, {here!r})
    import torch.distributed as dist
    import oracle_binding as O
    from compressjs_flattened_b200 import _native
    from compressjs_flattened_b200.bzip2 import Bzip2Engine
    from compressjs_flattened_b200.sharded import HostMailbox, compress_shard, gather_and_stitch
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = Bzip2Engine(0, _native.Library(os.path.join({here!r}, "sim", "libbz2b200_sim.so")))
    mb = HostMailbox(rank, world, os.environ["MASTER_PORT"])
    rng = np.random.default_rng(5)
    eng.debug_set_block_cap(300)
    for rnd in range(4):   # several rounds through the same mailbox: the alternating slots must not mix rounds up
        if rnd < 3:        # run-heavy: cuts fall inside runs, speculated starts are often wrong and must be recut
            data = bytes(np.repeat(rng.integers(0, 4, 4000, dtype=np.uint8), rng.choice([1, 1, 2, 5, 300], 4000)))[:30000]
        else:              # text-like: no run of 4, every speculated start must hold
            data = bytes(rng.integers(97, 123, 30000, dtype=np.uint8))
        n = len(data); sl = (n + world - 1) // world; base = rank * sl; own = max(0, min(sl, n - base))
        seg, info, off = compress_shard(eng, data[base:min(n, base + own + 20000)], base, own, 9, rank == world - 1, mailbox=mb)
        out = gather_and_stitch(eng, seg, info, 9)
        if rank == 0:
            O.set_block_cap(300)
            assert out == O.compress(data, 9), rnd
            print("MAILBOX_OK", rnd)
        if rnd == 2:
            before = dict(compress_shard.guesses)
    assert compress_shard.guesses[True] == before[True] + 1 and compress_shard.guesses[False] == before[False], (rank, compress_shard.guesses, before)
    mb.close()
    dist.destroy_process_group()
""")


def test_sharded_stream_through_host_mailbox(tmp_path, sim_engine, oracle):
    """Same protocol, scalars through the shared-memory mailbox (what bench.py uses on one node), 3 ranks, 3 rounds."""
    script = tmp_path / "worker_mb.py"
    script.write_text(MAILBOX_WORKER.format(root=ROOT, here=HERE))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("MAILBOX_OK") == 4
