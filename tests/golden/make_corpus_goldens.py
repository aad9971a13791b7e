"""Generates tests/golden/corpus_goldens.json: SHA-256 of the ORACLE's output (stable sort) for the
synthetic corpora at the BASELINE.json sizes, so GPU parity at full size needs no 100 MB oracle run
on the GPU box.  Run here (CPU):  python tests/golden/make_corpus_goldens.py"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_binding as O  # noqa: E402
from compressjs_flattened_b200.corpus import gen_html, gen_text  # noqa: E402

out = {}
cases = [("text", gen_text, 100_000_000, 8, 9), ("text", gen_text, 100_000_000, 8, 1), ("text", gen_text, 10_000_000, 8, 9),
         ("html", gen_html, 2_130_640, 5, 9), ("html", gen_html, 2_130_640, 5, 1)]
for kind, fn, n, seed, level in cases:
    t = time.time()
    data = fn(n, seed)
    comp, st = O.compress(data, level, O.SORT_STABLE, threads=os.cpu_count(), return_stats=True)
    key = f"{kind}:{n}:{seed}:L{level}"
    out[key] = dict(input_sha256=hashlib.sha256(data.tobytes()).hexdigest(), out_bytes=len(comp),
                    out_sha256=hashlib.sha256(comp).hexdigest(), n_blocks=st.n_blocks, rle1_bytes=st.rle1_bytes, mtf_syms=st.mtf_syms)
    print(key, out[key], f"{time.time() - t:.1f}s", flush=True)
    json.dump(out, open(os.path.join(HERE, "corpus_goldens.json"), "w"), indent=1, sort_keys=True)
