"""Development aid (CPU simulator): randomized decode stress of BOTH parse kernels -- streams of many shapes (alphabets of
1..256 symbols, skews, runs, text; 2..6 tables; block caps from 300 to 5000 bytes so that a stream has many blocks and
selector groups) made by the oracle (plus one libbz2-made stream of the same input per trial), decoded by the simulated kernels with BZ2B200_PARSE=1 and =2, compared with the input.
Damaged copies (one flipped bit) must give the oracle's outcome.
    python tests/sim_stress_decode.py [trials] [seed]"""
import bz2
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

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
lib = _native.Library(os.environ.get("BZ2B200_SIM_LIB") or os.path.join(HERE, "sim", "libbz2b200_sim.so"))  # e.g. an -fsanitize=address build


def outcome(fn):
    try:
        return ("ok", fn())
    except (Bzip2Error, O.OracleError) as e:
        return ("err", e.errorCode)


bad = 0
t0 = time.time()
for t in range(trials):
    kind = rng.choice(["iid", "runs", "text", "mixed"])
    n = int(rng.choice([1, 60, 700, 4000, 12_000, 30_000]) * rng.uniform(0.6, 1.3)) + 1
    A = int(rng.choice([1, 2, 3, 6, 17, 64, 128, 256]))
    skew = float(rng.choice([0.0, 0.7, 1.2, 2.5]))
    w = 1.0 / np.power(np.arange(1, A + 1), skew)
    w /= w.sum()
    syms = rng.permutation(256)[:A].astype(np.uint8)
    if kind == "iid":
        d = syms[rng.choice(A, n, p=w)]
    elif kind == "runs":
        m = max(1, n // 6)
        d = np.repeat(syms[rng.choice(A, m, p=w)], rng.choice([1, 1, 2, 3, 4, 5, 9, 255, 256, 300], m))[:n]
    elif kind == "text":
        d = gen_text(n, int(rng.integers(1, 1000)))
    else:
        a = syms[rng.choice(A, n // 2 + 1, p=w)]
        d = np.concatenate([a, gen_text(n // 2 + 1, 3), a[: n // 5]])[:n]
    d = np.ascontiguousarray(d, dtype=np.uint8)
    cap = int(rng.choice([300, 997, 2500, 5000]))
    O.set_block_cap(cap)
    comp = O.compress(d, int(rng.integers(1, 10)), O.SORT_STABLE)
    O.set_block_cap(0)
    flipped = bytearray(comp)
    if len(flipped) > 20:
        pos = int(rng.integers(32, len(flipped) * 8 - 1))
        flipped[pos >> 3] ^= 0x80 >> (pos & 7)
    want_bad = outcome(lambda: O.decompress(bytes(flipped)))
    foreign = bz2.compress(d.tobytes(), int(rng.integers(1, 10)))   # libbz2's table choices and code lengths
    for mode in ("1", "2"):
        os.environ["BZ2B200_PARSE"] = mode
        eng = Bzip2Engine(0, lib)
        got = outcome(lambda: eng.decompressFile(comp))
        got_bad = outcome(lambda: eng.decompressFile(bytes(flipped)))
        ok = got == ("ok", d.tobytes()) and got_bad == want_bad and outcome(lambda: eng.decompressFile(foreign)) == ("ok", d.tobytes())
        if not ok:
            bad += 1
            print(f"MISMATCH trial {t} mode {mode}: kind={kind} n={n} A={A} skew={skew} cap={cap} good={got[0]} bad={got_bad[0]}/{want_bad[0]}", flush=True)
        del eng
print(f"{trials} trials x 2 parse kernels, {bad} mismatches, {time.time() - t0:.0f} s", flush=True)
sys.exit(1 if bad else 0)
