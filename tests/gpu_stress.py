"""Development aid (GPU): randomized parity stress -- inputs of many shapes (alphabets of 1..256 symbols, Zipf skews, runs,
periodic pieces, text) at levels 1..9 through the single-launch path and the two-lane scheduler, every stream compared
with the oracle's; the decoder (whole, pool slices, stream object) must give the input back.
    python tests/gpu_stress.py [trials] [seed]"""
import io
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402

import oracle_binding as O  # noqa: E402
from compressjs_flattened_b200 import Bzip2Engine  # noqa: E402
from compressjs_flattened_b200.corpus import gen_html, gen_text  # noqa: E402
from compressjs_flattened_b200.pool import Bzip2Pool  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
eng = Bzip2Engine(0)
pool = Bzip2Pool([0], 2)
bad = 0
t0 = time.time()
for t in range(trials):
    kind = rng.choice(["iid", "runs", "periodic", "text", "html", "mixed"])
    n = int(rng.choice([1, 7, 100, 5000, 99_981, 100_000, 300_000, 899_981, 900_000, 1_500_000, 4_000_000]) * rng.uniform(0.6, 1.3)) + 1
    A = int(rng.choice([1, 2, 3, 4, 6, 16, 17, 40, 64, 65, 128, 200, 256]))
    skew = float(rng.choice([0.0, 0.7, 1.2, 2.5]))
    w = 1.0 / np.power(np.arange(1, A + 1), skew)
    w /= w.sum()
    syms = rng.permutation(256)[:A].astype(np.uint8)
    if kind == "iid":
        d = syms[rng.choice(A, n, p=w)]
    elif kind == "runs":
        m = max(1, n // 6)
        d = np.repeat(syms[rng.choice(A, m, p=w)], rng.choice([1, 1, 2, 3, 4, 5, 9, 255, 256, 300, 1000], m))[:n]
    elif kind == "periodic":
        p = int(rng.integers(1, 400))
        d = np.tile(syms[rng.choice(A, p, p=w)], n // p + 1)[:n]
    elif kind == "text":
        d = gen_text(n, int(rng.integers(1, 1000)))
    elif kind == "html":
        d = gen_html(n, int(rng.integers(1, 1000)))
    else:
        a = syms[rng.choice(A, n // 2 + 1, p=w)]
        b = gen_text(n // 2 + 1, 3)
        d = np.concatenate([a, b, a[: n // 5]])[:n]
    d = np.ascontiguousarray(d, dtype=np.uint8)
    level = int(rng.integers(1, 10))
    exp = O.compress(d, level, threads=8)
    eng.debug_set_pool(1 << 62)
    got1 = eng.compressFile(d, None, level)
    shard = int(rng.choice([0, 150_000, 400_000, 1_000_000]))
    pool.set_plan(int(rng.choice([0, 200_000])), 0.0)
    got2 = pool.compressFile(d, None, level, shard_bytes=shard)
    ok = got1 == exp and got2 == exp
    if ok:
        raw = d.tobytes()
        ok = eng.decompressFile(exp) == raw and pool.decompressFile(exp, len(raw), False, slice_bytes=int(rng.choice([0, 70_000, 500_000]))) == raw
        if ok and t % 4 == 0:
            ok = eng.decompressStream(io.BytesIO(exp), None, False, chunk_bytes=1 << 18) == raw and \
                eng.compressStream(io.BytesIO(raw), None, level, chunk_bytes=1 << 20) == exp
    if not ok:
        bad += 1
        np.save(f"gpurun_out/stress_fail_{t}.npy", d)
        print(f"MISMATCH trial {t}: kind={kind} n={n} A={A} skew={skew} level={level} shard={shard} single_ok={got1 == exp} pool_ok={got2 == exp}", flush=True)
print(f"{trials} trials, {bad} mismatches, {time.time() - t0:.0f} s", flush=True)
sys.exit(1 if bad else 0)
