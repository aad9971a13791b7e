"""Development aid (CPU simulator): randomized stress of the stream objects and of the scheduler's decoder -- random feed
sizes into bz2b200_zstream / bz2b200_dstream, concatenated streams with multistream on and off, pool decompression with
random slice sizes and batch limits, and one damaged copy per trial whose outcome (bytes or error code) must equal the
oracle's.  Meant for the sanitizer builds too (tests/sim/build_sim.sh).
    python tests/sim_stress_streams.py [trials] [seed]"""
import io
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402

import oracle_binding as O  # noqa: E402
from compressjs_flattened_b200 import _native  # noqa: E402
from compressjs_flattened_b200.bzip2 import Bzip2Engine, Bzip2Error  # noqa: E402
from compressjs_flattened_b200.corpus import gen_text  # noqa: E402
from compressjs_flattened_b200.pool import Bzip2Pool  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 9)
lib = _native.Library(os.environ.get("BZ2B200_SIM_LIB") or os.path.join(HERE, "sim", "libbz2b200_sim.so"))
eng = Bzip2Engine(0, lib)
pool = Bzip2Pool([0], 2, library=lib)


def outcome(fn):
    try:
        return ("ok", bytes(fn()))
    except (Bzip2Error, O.OracleError) as e:
        return ("err", e.errorCode)


def piece(n):
    a = int(rng.choice([0, 1, 3, 40, 200, 256]))
    if a:
        w = 1.0 / np.arange(1, a + 1)
        return rng.permutation(256)[:a].astype(np.uint8)[rng.choice(a, n, p=w / w.sum())]
    return gen_text(n, int(rng.integers(1, 999)))


bad = 0
t0 = time.time()
for t in range(trials):
    cap = int(rng.choice([300, 997, 2500]))
    level = int(rng.integers(1, 10))
    n = int(rng.choice([0, 1, 50, 900, 6000, 20_000]) * rng.uniform(0.6, 1.3))
    d = np.ascontiguousarray(piece(n) if n else np.zeros(0, np.uint8), dtype=np.uint8)
    raw = d.tobytes()
    O.set_block_cap(cap)
    eng.debug_set_block_cap(cap)
    pool.debug(cap, int(rng.choice([0, 3])), 0, False)
    exp = O.compress(d, level, O.SORT_STABLE)
    d2 = piece(int(rng.integers(1, 3000)))
    exp2 = O.compress(d2, int(rng.integers(1, 10)), O.SORT_STABLE)
    O.set_block_cap(0)
    cb = int(rng.choice([1 << 10, 1 << 12, 1 << 16]))
    pb = int(rng.choice([1, 7, 333, 4096]))
    fails = []
    if eng.compressStream(io.BytesIO(raw), None, level, chunk_bytes=cb, piece_bytes=pb) != exp:
        fails.append("zstream")
    if outcome(lambda: eng.decompressStream(io.BytesIO(exp), None, False, chunk_bytes=cb, piece_bytes=pb)) != ("ok", raw):
        fails.append("dstream")
    both = exp + exp2 + bytes(int(rng.integers(0, 5)))   # a second stream, then a few stray bytes
    for ms in (False, True):
        want = outcome(lambda: O.decompress(both, ms))
        if outcome(lambda: eng.decompressStream(io.BytesIO(both), None, ms, chunk_bytes=cb, piece_bytes=pb)) != want:
            fails.append(f"dstream multistream={ms}")
        if outcome(lambda: eng.decompressFile(both, None, ms)) != want:
            fails.append(f"decompressFile multistream={ms}")
        pool.debug_decode_batch(int(rng.choice([0, 2, 5])))
        if outcome(lambda: pool.decompressFile(both, None, ms, slice_bytes=int(rng.choice([0, 64, 500, 4000])))) != want:
            fails.append(f"pool multistream={ms}")
    if len(exp) > 20:
        dam = bytearray(both)
        pos = int(rng.integers(32, len(exp) * 8 - 1))
        dam[pos >> 3] ^= 0x80 >> (pos & 7)
        dam = bytes(dam)
        want = outcome(lambda: O.decompress(dam, True))
        got = [outcome(lambda: eng.decompressFile(dam, None, True)),
               outcome(lambda: pool.decompressFile(dam, None, True, slice_bytes=int(rng.choice([0, 100, 3000]))))]
        if any(g != want for g in got):
            fails.append(f"damaged whole/pool {[g[0] if g[0] == 'ok' else g for g in got]} want {want[0] if want[0] == 'ok' else want}")
        # the stream decoder has delivered the blocks before the error when it raises: only the error code is compared
        gs = outcome(lambda: eng.decompressStream(io.BytesIO(dam), None, True, chunk_bytes=cb, piece_bytes=pb))
        if gs[0] != want[0] or (gs[0] == "err" and gs != want) or (gs[0] == "ok" and gs != want):
            fails.append(f"damaged dstream {gs[0] if gs[0] == 'ok' else gs} want {want[0] if want[0] == 'ok' else want}")
    if fails:
        bad += 1
        print(f"MISMATCH trial {t}: n={n} cap={cap} level={level} chunk={cb} piece={pb}: {fails}", flush=True)
print(f"{trials} trials, {bad} mismatches, {time.time() - t0:.0f} s", flush=True)
sys.exit(1 if bad else 0)
