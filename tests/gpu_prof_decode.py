"""Development aid: one compress + two decompress calls of a 100 MB level-9 stream (for ncu launch lists)."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE))
from compressjs_flattened_b200 import Bzip2Engine
from compressjs_flattened_b200.corpus import gen_text
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
level = int(sys.argv[2]) if len(sys.argv) > 2 else 9
E = Bzip2Engine(0)
d = gen_text(mb * 1_000_000, 8)
c = E.compressFile(d, None, level)
print("compress launches", E.stats().kernel_launches)
for i in range(2):
    b = E.decompressFile(c)
    st = E.stats()
    print("decode", i, len(b), "launches", st.kernel_launches, "ms", round(st.ms_total, 2), [round(x, 2) for x in st.ms_stage[:5]])
assert b == d.tobytes()
