"""GPU parity tests: the CUDA path, called through the C-ABI (compressjs_flattened_b200 -> libbz2b200.so),
against the CPU oracle on the same seeded inputs, against the committed goldens, and -- at BASELINE.json's
full sizes -- through SHA-256 goldens of the oracle's output plus size-independent properties
(round trip, block index == encoder's block table, CRC folding)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import HERE, fixture_bytes

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(HERE, "golden", "corpus_goldens.json")))
RNG = np.random.default_rng(11)
EDGE = {
    "empty": b"", "one": b"Q", "aaaa": b"aaaa", "aaaaa": b"aaaaa", "a256": b"a" * 256, "a255x": b"a" * 255 + b"x",
    "zeros1000": bytes(1000), "abab": b"abab", "abc3": b"abcabcabc",
    "rand64k": RNG.integers(0, 256, 65536, dtype=np.uint8).tobytes(),
    "rand4sym": RNG.integers(0, 4, 20000, dtype=np.uint8).tobytes(),
    "two_sym_d1": bytes(RNG.integers(0, 2, 6000, dtype=np.uint8)),
    "period257": bytes(list(range(256)) + [0]) * 3900,
    "line97": (b"0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ-the quick brown fox jumps over a do\n") * 20000,
}


@pytest.mark.parametrize("name", sorted(EDGE))
def test_edge_cases_byte_identical(gpu_engine, oracle, name):
    data = EDGE[name]
    for level in (9, 1):
        exp, st = oracle.compress(data, level, threads=8, return_stats=True)
        got = gpu_engine.compressFile(data, None, level)
        assert got == exp, f"{name} L{level}: {len(got)} vs {len(exp)} bytes"
        if not st.d1_triggered:
            assert gpu_engine.decompressFile(exp) == data


@pytest.mark.parametrize("n", range(6))
@pytest.mark.parametrize("level", [1, 2, 5, 9])
def test_reference_samples_byte_identical(gpu_engine, oracle, n, level):
    data = fixture_bytes(f"sample{n}.ref")
    exp = oracle.compress(data, level, threads=8)
    got = gpu_engine.compressFile(data, None, level)
    assert got == exp
    assert gpu_engine.decompressFile(got) == data  # NPM/test/file.js round trip


def test_default_level_is_9(gpu_engine, oracle):
    data = fixture_bytes("sample1.ref")
    assert gpu_engine.compressFile(data) == oracle.compress(data, 9) == gpu_engine.compressFile(data, None, None)


def test_adversarial_suffix_sort_inputs(gpu_engine, oracle):
    """BASELINE config 5: all-zero and periodic inputs; the doubling must stop at h >= n and apply the tie rule."""
    cases = [bytes(2_000_000), b"ab" * 500_000, b"abc" * 333_333, b"a" * 899_981, bytes(5) * 179_996]
    for data in cases:
        exp = oracle.compress(data, 9, threads=8)
        assert gpu_engine.compressFile(data, None, 9) == exp
        assert gpu_engine.stats().sort_rounds <= 21
        assert gpu_engine.decompressFile(exp) == data
    assert len(gpu_engine.compressFile(bytes(2_000_000), None, 9)) == 49


def test_stage_dumps_match_oracle(gpu_engine, oracle):
    data = fixture_bytes("sample5.ref")
    gpu_engine.compressFile(data, None, 9)
    recs, metas = gpu_engine.block_table(), gpu_engine.block_meta()
    starts, lens, crcs = oracle.cut_points(data, 9)
    assert [(r.s, r.n, r.crc) for r in recs] == list(zip(starts[:-1], lens, crcs))
    facts = [(899981, 298344, 196, 293714), (899981, 105439, 202, 235309), (330739, 144412, 193, 104124)]
    for k, r in enumerate(recs):
        blk, _, _ = oracle.rle1_block(data[starts[k]:], 899981)
        assert gpu_engine.debug_fetch(1, k, r.n).tobytes() == blk.tobytes()
        st = oracle.block_stages(blk)
        assert gpu_engine.debug_fetch(2, k, r.n).tobytes() == st["U"].tobytes()
        A = np.frombuffer(gpu_engine.debug_fetch(3, k, 2 * metas[k].m).tobytes(), dtype=np.uint16)
        assert np.array_equal(A, st["A"])
        assert (r.n, r.orig_ptr, metas[k].alpha, metas[k].m) == facts[k] == (st["n"], st["orig_ptr"], st["alpha"], st["m"])
        assert (metas[k].n_groups, metas[k].n_sel, metas[k].bits) == (st["n_groups"], st["n_sel"], st["bits"] + 80)


def test_cut_points_with_runs_across_blocks(gpu_engine, oracle):
    rng = np.random.default_rng(3)
    parts = []
    for _ in range(3000):
        parts.append(bytes([int(rng.integers(0, 256))]) * int(rng.choice([1, 1, 1, 2, 3, 4, 5, 7, 255, 256, 257, 600, 4000])))
    data = b"".join(parts) * 3
    for level in (1, 9):
        exp = oracle.compress(data, level, threads=8)
        assert gpu_engine.compressFile(data, None, level) == exp
        starts, lens, crcs = oracle.cut_points(data, level)
        assert [(r.s, r.n, r.crc) for r in gpu_engine.block_table()] == list(zip(starts[:-1], lens, crcs))
        assert gpu_engine.decompressFile(exp) == data


def test_decode_reference_fixtures(gpu_engine):
    for n in range(5):  # NPM/test/bzip2-basic.js, streams made by real bzip2
        assert gpu_engine.decompressFile(fixture_bytes(f"sample{n}.bz2"), len(fixture_bytes(f"sample{n}.ref"))) == fixture_bytes(f"sample{n}.ref")
        rows = []  # NPM/test/bzip2-table.js
        gpu_engine.table(fixture_bytes(f"sample{n}.bz2"), lambda p, s: rows.append(f"{p}\t{s}\n"))
        assert "".join(rows) == fixture_bytes(f"sample{n}.bzt").decode()
    assert gpu_engine.decompressBlock(fixture_bytes("sample0.bz2"), 32) == b"This is a test\n"  # NPM/test/bzip2-block.js
    for f, b in [("sample2", 544888), ("sample4", 32), ("sample4", 1596228), ("sample4", 2342106)]:
        assert gpu_engine.decompressBlock(fixture_bytes(f + ".bz2"), b, len(fixture_bytes(f"{f}.{b}"))) == fixture_bytes(f"{f}.{b}")


def test_decode_errors_match_oracle(gpu_engine, oracle):
    from compressjs_flattened_b200.bzip2 import Bzip2Error
    good = oracle.compress(fixture_bytes("sample1.ref")[:30000], 9)
    blobs = [b"", b"BZ", b"BZh0xxxx", b"XXXXXXXX", good[:-3], good[:200]]
    for pos in (10, 15, 40, 200, len(good) - 2):
        x = bytearray(good)
        x[pos] ^= 0x40
        blobs.append(bytes(x))
    for blob in blobs:
        try:
            exp = ("ok", oracle.decompress(blob))
        except oracle.OracleError as e:
            exp = ("err", e.errorCode)
        try:
            got = ("ok", gpu_engine.decompressFile(blob))
        except Bzip2Error as e:
            got = ("err", e.errorCode)
        assert got == exp
    ms = oracle.compress(b"first stream ") + oracle.compress(b"second stream", 1)
    assert gpu_engine.decompressFile(ms, None, True) == b"first stream second stream"
    assert gpu_engine.decompressFile(ms) == b"first stream "
    with pytest.raises(ValueError, match="Invalid block size multiplier"):
        gpu_engine.compressFile(b"x", None, 0)


@pytest.mark.parametrize("key", ["html:2130640:5:L9", "html:2130640:5:L1", "text:10000000:8:L9", "text:100000000:8:L9", "text:100000000:8:L1"])
def test_full_size_configs(gpu_engine, key):
    """BASELINE.json configs 1-3 at full size: SHA-256 of the GPU output == SHA-256 of the oracle's output
    (tests/golden/corpus_goldens.json, generated by tests/golden/make_corpus_goldens.py), then the
    size-independent properties: round trip, block index from the magic search == encoder's block table."""
    from compressjs_flattened_b200.corpus import gen_html, gen_text
    kind, n, seed, level = key.split(":")
    n, seed, level = int(n), int(seed), int(level[1:])
    data = (gen_text if kind == "text" else gen_html)(n, seed)
    g = GOLD[key]
    assert hashlib.sha256(data.tobytes()).hexdigest() == g["input_sha256"]
    comp = gpu_engine.compressFile(data, None, level)   # >= 32 MB: shards over the context's two lanes (pool.inl)
    st = gpu_engine.stats()
    assert (len(comp), hashlib.sha256(comp).hexdigest()) == (g["out_bytes"], g["out_sha256"])
    assert (st.n_blocks, st.rle1_bytes, st.mtf_syms) == (g["n_blocks"], g["rle1_bytes"], g["mtf_syms"])
    if n >= 32_000_000:   # the same through the single-launch path, whose block table the checks below read
        gpu_engine.debug_set_pool(1 << 62)
        try:
            comp1 = gpu_engine.compressFile(data, None, level)
        finally:
            gpu_engine.debug_set_pool()
        assert comp1 == comp
    recs = gpu_engine.block_table()
    sizes_enc = [r.p - r.s for r in recs]
    fold = 0
    for r in recs:  # combined CRC = fold of the block CRCs (BJ:2237); it is the last 32 bits before the zero padding
        fold = (((fold << 1) | (fold >> 31)) ^ r.crc) & 0xFFFFFFFF
    end_bit = 32 + sum(m.bits for m in gpu_engine.block_meta()) + 80
    assert (end_bit + 7) // 8 == len(comp)
    tail = int.from_bytes(comp[-16:], "big") >> (len(comp) * 8 - end_bit)
    assert tail & 0xFFFFFFFF == fold and (tail >> 32) & 0xFFFFFFFFFFFF == 0x177245385090
    back = gpu_engine.decompressFile(comp)
    assert hashlib.sha256(back).hexdigest() == g["input_sha256"]
    rows = []
    gpu_engine.table(comp, lambda p, s: rows.append((p, s)))
    assert [s for _, s in rows] == sizes_enc and rows[0][0] == 32


def test_batched_blocks_same_stream(gpu_engine):
    """A call with more blocks than one batch holds (forced here: 7 blocks per batch at level 1): same bytes,
    same block table, as the single-batch run and the oracle's golden."""
    from compressjs_flattened_b200.corpus import gen_text
    data = gen_text(10_000_000, 8)
    whole = gpu_engine.compressFile(data, None, 1)
    recs_whole = [(r.s, r.p, r.n, r.crc, r.orig_ptr) for r in gpu_engine.block_table()]
    try:
        for per in (7, 64):
            gpu_engine.debug_set_batch_blocks(per)
            assert gpu_engine.compressFile(data, None, 1) == whole
            assert [(r.s, r.p, r.n, r.crc, r.orig_ptr) for r in gpu_engine.block_table()] == recs_whole
    finally:
        gpu_engine.debug_set_batch_blocks(0)
    assert gpu_engine.decompressFile(whole) == data.tobytes()
    g = GOLD["text:10000000:8:L9"]
    assert hashlib.sha256(gpu_engine.compressFile(data, None, 9)).hexdigest() == g["out_sha256"]


def test_shards_in_process_equal_whole_stream(gpu_engine):
    """The multi-GPU protocol (begin / cut / compress / emit / stitch) driven on one device: 4 shards of a
    20 MB stream, every segment emitted at its bit phase, stitched on the host == the single-call stream."""
    from compressjs_flattened_b200.corpus import gen_text
    data = gen_text(20_000_000, 8).tobytes()
    whole = gpu_engine.compressFile(data, None, 9)
    n, world = len(data), 4
    slice_len = (n + world - 1) // world
    start, bitpos, segs, infos = 0, 32, [], []
    for r in range(world):
        base = r * slice_len
        own = max(0, min(slice_len, n - base))
        gpu_engine.shard_begin(data[base:min(n, base + own + 1_200_000)], 9)
        info = gpu_engine.shard_cut(max(start - base, 0), own, r == world - 1)
        assert info.complete
        start = max(base + info.next_start, start)
        gpu_engine.shard_compress(info)
        segs.append(gpu_engine.shard_emit(info, bitpos & 7))
        infos.append(info)
        bitpos += info.bits
    assert gpu_engine.stitch_shards(9, segs, infos) == whole


def test_compress_stream_equals_whole(gpu_engine):
    """Stream flavour on the GPU: 10 MB of text in 3 MB chunks == compressFile of the whole, and decodes back."""
    import io
    from compressjs_flattened_b200.corpus import gen_text
    data = gen_text(10_000_000, 8).tobytes()
    g = GOLD["text:10000000:8:L9"]
    out = gpu_engine.compressStream(io.BytesIO(data), None, 9, chunk_bytes=3_000_000)
    assert hashlib.sha256(out).hexdigest() == g["out_sha256"]
    sink = io.BytesIO()
    gpu_engine.compressStream(io.BytesIO(data), sink, 1, chunk_bytes=1_000_000)
    assert sink.getvalue() == gpu_engine.compressFile(data, None, 1)


@pytest.mark.parametrize("parse_mode", ["1", "2"])
def test_decode_damaged_streams_match_oracle(oracle, parse_mode, monkeypatch):
    """Both parse kernels (BZ2B200_PARSE: 1 = one CTA per block and group, 2 = speculative window) on a full level-9
    block (~11 000 selectors) whose selector list, code lengths or symbol data are damaged: same error code as the
    oracle, and -- with the block CRC comparison switched off on both sides -- the same decoded bytes."""
    from compressjs_flattened_b200.bzip2 import Bzip2Engine, Bzip2Error
    from compressjs_flattened_b200.corpus import gen_text
    from test_sim_kernels import _decode_outcome, _selector_region
    monkeypatch.setenv("BZ2B200_PARSE", parse_mode)
    eng = Bzip2Engine(0)
    rng = np.random.default_rng(13)
    data = gen_text(880_000, 4).tobytes()
    good = oracle.compress(data, 9)
    assert eng.decompressFile(good) == data
    s0 = _selector_region(good)
    cases = []
    for _ in range(10):
        b = bytearray(good)
        for pos in rng.integers(s0, s0 + 40_000, rng.integers(1, 3)):
            b[pos >> 3] ^= 0x80 >> (pos & 7)
        cases.append(bytes(b))
    for _ in range(14):
        b = bytearray(good)
        pos = int(rng.integers(8 * len(good) // 4, 8 * len(good) - 200))
        b[pos >> 3] ^= 0x80 >> (pos & 7)
        cases.append(bytes(b))
    for cut in (s0 // 8 + 1000, len(good) // 2, len(good) - 9):
        cases.append(good[:cut])
    for blob in cases:
        assert _decode_outcome(eng, blob, Bzip2Error) == _decode_outcome(oracle, blob, oracle.OracleError)
    eng.debug_set_ignore_block_crc(True)
    oracle.set_ignore_block_crc(True)
    try:
        outcomes = [_decode_outcome(oracle, blob, oracle.OracleError) for blob in cases]
        for blob, exp in zip(cases, outcomes):
            assert _decode_outcome(eng, blob, Bzip2Error) == exp
        assert sum(1 for o in outcomes if o[0] == "ok") >= 3 and sum(1 for o in outcomes if o[0] == "err") >= 3
    finally:
        oracle.set_ignore_block_crc(False)


@pytest.mark.gpu
def test_decode_run_length_sum_cannot_wrap(gpu_engine, oracle):
    """run lengths that sum past 2^32 are a Data error (BJ:1647), never an out-of-bounds write."""
    from compressjs_flattened_b200.bzip2 import Bzip2Error
    from test_sim_kernels import _decode_outcome, crafted_run_overflow_stream
    for deep, tail in ((1023, 1000), (1024, 5000), (2047, 10)):
        blob, _ = crafted_run_overflow_stream(deep, tail)
        assert _decode_outcome(gpu_engine, blob, Bzip2Error) == ("err", -5) == _decode_outcome(oracle, blob, oracle.OracleError)
    good = oracle.compress(b"still alive after the crafted streams " * 100)
    assert gpu_engine.decompressFile(good) == b"still alive after the crafted streams " * 100
