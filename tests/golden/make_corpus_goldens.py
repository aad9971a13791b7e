"""Generates tests/golden/corpus_goldens.json: SHA-256 of the ORACLE's output (stable sort) for the
synthetic corpora at the BASELINE.json sizes, so GPU parity at full size needs no oracle run of that
size on the GPU box.  Run here (CPU):  python tests/golden/make_corpus_goldens.py [--only-missing] [--big]

  text:<n>:<seed>:L<level>   gen_text      configs 1-3, the bench streams at N = 2/4/8 (200/400/800 MB, seed 8),
                             and with --big BASELINE config 4: the 8 GB corpus (seed 64)
  html:...                   gen_html
  adv:<name>:L<level>        corpus.gen_adversarial (SURVEY 8d C5b at full size)
Every entry also records the bit position and decoded size of a few blocks (first, middle, last), which pins the
block table of streams too large to keep."""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_binding as O  # noqa: E402
from compressjs_flattened_b200.corpus import ADVERSARIAL, gen_adversarial, gen_html, gen_text  # noqa: E402

PATH = os.path.join(HERE, "corpus_goldens.json")
only_missing = "--only-missing" in sys.argv
big = "--big" in sys.argv
out = json.load(open(PATH)) if os.path.exists(PATH) else {}
cases = [("text", 100_000_000, 8, 9), ("text", 100_000_000, 8, 1), ("text", 10_000_000, 8, 9),
         ("html", 2_130_640, 5, 9), ("html", 2_130_640, 5, 1),
         ("text", 200_000_000, 8, 9), ("text", 400_000_000, 8, 9), ("text", 800_000_000, 8, 9), ("text", 120_000_000, 8, 9)]
cases += [("adv", name, None, lv) for name in ADVERSARIAL for lv in (9, 1)]
if big:
    cases += [("text", 1_000_000_000, 64, 9), ("text", 8_000_000_000, 64, 9)]
for kind, n, seed, level in cases:
    key = f"adv:{n}:L{level}" if kind == "adv" else f"{kind}:{n}:{seed}:L{level}"
    if only_missing and key in out:
        continue
    t = time.time()
    data = gen_adversarial(n) if kind == "adv" else (gen_text if kind == "text" else gen_html)(n, seed)
    comp, st = O.compress(data, level, O.SORT_STABLE, threads=os.cpu_count(), return_stats=True)
    h = hashlib.sha256()
    h.update(memoryview(data))
    out[key] = dict(input_sha256=h.hexdigest(), out_bytes=len(comp),
                    out_sha256=hashlib.sha256(comp).hexdigest(), n_blocks=st.n_blocks, rle1_bytes=st.rle1_bytes, mtf_syms=st.mtf_syms)
    print(key, out[key], f"{time.time() - t:.1f}s", flush=True)
    json.dump(out, open(PATH, "w"), indent=1, sort_keys=True)
