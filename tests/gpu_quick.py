"""Quick GPU parity + timing run (development aid; the real tests are tests/test_gpu_*.py).

  python tests/gpu_quick.py [--big] [--decode]
Prints one line per case; exits non-zero on the first parity failure.
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402

import oracle_binding as O  # noqa: E402
from compressjs_flattened_b200 import Bzip2Engine  # noqa: E402
from compressjs_flattened_b200.corpus import gen_html, gen_text  # noqa: E402

G = os.path.join(HERE, "golden", "ref_fixtures")
big = "--big" in sys.argv
decode = "--decode" in sys.argv
E = Bzip2Engine(0)
fails = 0


def stage_diff(data, level):
    cap = level * 100000 - 19
    starts, lens, crcs = O.cut_points(data, level)
    recs = E.block_table()
    metas = E.block_meta()
    for k, r in enumerate(recs):
        if k >= len(lens):
            print("   extra block", k)
            return
        if (r.s, r.n, r.crc) != (starts[k], lens[k], crcs[k]):
            print(f"   blk{k}: cut/crc differ: got s={r.s} n={r.n} crc={r.crc:08x}; oracle s={starts[k]} n={lens[k]} crc={crcs[k]:08x}")
            return
        blk, _, _ = O.rle1_block(data[starts[k]:], cap)
        if E.debug_fetch(1, k, r.n).tobytes() != blk.tobytes():
            print(f"   blk{k}: RLE1 bytes differ")
            return
        st = O.block_stages(blk)
        gl = E.debug_fetch(2, k, r.n)
        if gl.tobytes() != st["U"].tobytes() or r.orig_ptr != st["orig_ptr"]:
            d = np.flatnonzero(gl != st["U"])
            print(f"   blk{k}: BWT differs: origPtr {r.orig_ptr} vs {st['orig_ptr']}, {d.size} bytes differ, first {d[:5]}")
            return
        mm = metas[k]
        ga = np.frombuffer(E.debug_fetch(3, k, 2 * mm.m).tobytes(), dtype=np.uint16)
        if mm.m != st["m"] or not np.array_equal(ga, st["A"]):
            print(f"   blk{k}: MTF/RLE2 differs: m {mm.m} vs {st['m']}, alpha {mm.alpha} vs {st['alpha']}")
            return
        if (mm.n_groups, mm.bits) != (st["n_groups"], st["bits"] + 80):
            print(f"   blk{k}: Huffman/emit differs: groups {mm.n_groups} vs {st['n_groups']}, bits {mm.bits} vs {st['bits'] + 80}")
            return
    print("   all stage dumps equal; the stitch differs")


def check(name, data, level, golden=None):
    global fails
    data = bytes(data)
    t0 = time.time()
    got = E.compressFile(data, None, level)
    t1 = time.time()
    st = E.stats()
    if golden is not None:
        ok = hashlib.sha256(got).hexdigest() == golden["out_sha256"]
        exp_len = golden["out_bytes"]
    else:
        exp = O.compress(data, level, threads=os.cpu_count())
        ok = got == exp
        exp_len = len(exp)
    stages = " ".join(f"{x:.2f}" for x in st.ms_stage[:5])
    print(f"{name:28s} L{level} n={len(data):>10d} out={len(got):>9d} exp={exp_len:>9d} {'OK ' if ok else 'BAD'} blocks={st.n_blocks} rounds={st.sort_rounds} "
          f"launches={st.kernel_launches} dev_ms={st.ms_total:.2f} [{stages}] wall={t1 - t0:.3f}s  {len(data) / 1e6 / max(st.ms_total, 1e-9) * 1e3:.1f} MB/s(dev)", flush=True)
    if not ok:
        fails += 1
        if golden is None or len(data) <= 12_000_000:
            stage_diff(data, level)
    if decode and ok:
        t0 = time.time()
        back = E.decompressFile(got)
        t1 = time.time()
        st = E.stats()
        okd = back == data
        stages = " ".join(f"{x:.2f}" for x in st.ms_stage[:5])
        print(f"{'  decode':28s}    n={len(got):>10d} out={len(back):>9d} {'OK ' if okd else 'BAD'} launches={st.kernel_launches} dev_ms={st.ms_total:.2f} [{stages}] "
              f"wall={t1 - t0:.3f}s {len(back) / 1e6 / max(st.ms_total, 1e-9) * 1e3:.1f} MB/s(dev)", flush=True)
        if not okd:
            fails += 1
    return ok


rng = np.random.default_rng(1)
small = [("empty", b""), ("Q", b"Q"), ("aaaa", b"aaaa"), ("a256", b"a" * 256), ("zeros1000", bytes(1000)),
         ("abcabcabc", b"abcabcabc"), ("rand4sym", rng.integers(0, 4, 20000, dtype=np.uint8).tobytes()),
         ("rand64k", rng.integers(0, 256, 65536, dtype=np.uint8).tobytes())]
for name, d in small:
    check(name, d, 9)
for n in range(6):
    d = open(os.path.join(G, f"sample{n}.ref"), "rb").read()
    check(f"sample{n}", d, 9)
    check(f"sample{n}", d, 1)
check("zeros2M", bytes(2_000_000), 9)
check("ab x 500k", b"ab" * 500_000, 9)
check("html2M", gen_html(2_130_640, 5), 9)
gold = {}
gp = os.path.join(HERE, "golden", "corpus_goldens.json")
if os.path.exists(gp):
    gold = json.load(open(gp))
check("text10M", gen_text(10_000_000, 8), 9, gold.get("text:10000000:8:L9"))
if big:
    t = gen_text(100_000_000, 8)
    for rep in range(2):
        check("text100M", t, 9, gold.get("text:100000000:8:L9"))
    for rep in range(2):
        check("text100M", t, 1, gold.get("text:100000000:8:L1"))
print("FAILS", fails)
sys.exit(1 if fails else 0)
