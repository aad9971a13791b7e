"""Development aid: isolate a decode failure on the GPU."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import oracle_binding as O
from compressjs_flattened_b200 import Bzip2Engine, Bzip2Error
E = Bzip2Engine(0)
G = os.path.join(HERE, "golden", "ref_fixtures")
d = open(os.path.join(G, "sample2.ref"), "rb").read()
if len(sys.argv) > 1: d = d[:int(sys.argv[1])]
c = O.compress(d, 1)
tab = O.table(c)
print("oracle table", tab)
for rep in range(2):
    try:
        r = E.decompressFile(c); print("decompressFile ok", r == d)
    except Bzip2Error as e:
        print("decompressFile ERR", e.errorCode, e)
    off = 0
    for pos, size in tab:
        try:
            b = E.decompressBlock(c, pos); print(" block", pos, "ok", b == d[off:off + size], len(b), size)
        except Bzip2Error as e:
            print(" block", pos, "ERR", e.errorCode)
        off += size
    t = []
    try:
        E.table(c, lambda p, s: t.append((p, s))); print("table", t)
    except Bzip2Error as e:
        print("table ERR", e.errorCode, e)
