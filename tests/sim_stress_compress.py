"""Development aid (CPU simulator): randomized compress stress -- inputs of many shapes at every level and block caps of 300 to
5000 bytes through the single-launch path, the context's two-lane scheduler and a pool with random shard sizes; every stream
must equal the oracle's.  Meant to be run on a sanitizer build as well (tests/sim/build_sim.sh).
    python tests/sim_stress_compress.py [trials] [seed]"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402

import oracle_binding as O  # noqa: E402
from compressjs_flattened_b200 import _native  # noqa: E402
from compressjs_flattened_b200.bzip2 import Bzip2Engine  # noqa: E402
from compressjs_flattened_b200.corpus import gen_html, gen_text  # noqa: E402
from compressjs_flattened_b200.pool import Bzip2Pool  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
lib = _native.Library(os.environ.get("BZ2B200_SIM_LIB") or os.path.join(HERE, "sim", "libbz2b200_sim.so"))
eng = Bzip2Engine(0, lib)
pool = Bzip2Pool([0], 2, library=lib)
bad = 0
t0 = time.time()
for t in range(trials):
    kind = rng.choice(["iid", "runs", "periodic", "text", "html", "mixed"])
    n = int(rng.choice([1, 7, 300, 2000, 9000, 25_000]) * rng.uniform(0.6, 1.3)) + 1
    A = int(rng.choice([1, 2, 3, 4, 16, 17, 64, 65, 200, 256]))
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
        d = np.concatenate([a, gen_text(n // 2 + 1, 3), a[: n // 5]])[:n]
    d = np.ascontiguousarray(d, dtype=np.uint8)
    level = int(rng.integers(1, 10))
    cap = int(rng.choice([300, 997, 2500, 5000]))
    O.set_block_cap(cap)
    exp = O.compress(d, level, O.SORT_STABLE)
    O.set_block_cap(0)
    eng.debug_set_block_cap(cap)
    eng.debug_set_pool(1 << 62)
    got1 = eng.compressFile(d, None, level)             # single launch
    eng.debug_set_pool(1, int(rng.choice([0, 700, 4000])))
    got2 = eng.compressFile(d, None, level)             # the context's own two lanes
    pool.debug(cap, 0, int(rng.choice([0, 64])), int(rng.integers(0, 2)))
    got3 = pool.compressFile(d, None, level, shard_bytes=int(rng.choice([0, 500, 3000, 20_000])))
    if not (got1 == exp and got2 == exp and got3 == exp):
        bad += 1
        print(f"MISMATCH trial {t}: kind={kind} n={n} A={A} skew={skew} level={level} cap={cap} single={got1 == exp} ctx_pool={got2 == exp} pool={got3 == exp}", flush=True)
print(f"{trials} trials, {bad} mismatches, {time.time() - t0:.0f} s", flush=True)
sys.exit(1 if bad else 0)
