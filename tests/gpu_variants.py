"""Development aid: time alternative builds of the library (tests/variant_*.so) on the 100 MB bench text."""
import glob
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from compressjs_flattened_b200 import _native
from compressjs_flattened_b200.bzip2 import Bzip2Engine
from compressjs_flattened_b200.corpus import gen_text

gold = json.load(open(os.path.join(HERE, "golden", "corpus_goldens.json")))["text:100000000:8:L9"]
data = gen_text(100_000_000, 8)
for so in sorted(glob.glob(os.path.join(HERE, "variant_*.so"))):
    E = Bzip2Engine(0, _native.Library(so))
    best = None
    for rep in range(3):
        out = E.compressFile(data, None, 9)
        st = E.stats()
        if best is None or st.ms_total < best[0]:
            best = (st.ms_total, [round(x, 2) for x in st.ms_stage[:5]])
    ok = hashlib.sha256(out).hexdigest() == gold["out_sha256"]
    print(f"{os.path.basename(so):24s} {'OK ' if ok else 'BAD'} total {best[0]:.2f} ms stages {best[1]}", flush=True)
    del E
