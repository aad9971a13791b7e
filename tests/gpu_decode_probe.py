"""Development aid (GPU): host-call decompression, single context vs the pool (lanes x slices), 100 MB and 400 MB of text."""
import ctypes as C
import os
import sys
import time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import torch  # noqa: E402
from compressjs_flattened_b200 import Bzip2Engine  # noqa: E402
from compressjs_flattened_b200.corpus import gen_text  # noqa: E402
from compressjs_flattened_b200.pool import Bzip2Pool  # noqa: E402
eng = Bzip2Engine(0)
L = eng._L
for mb in (100, 400):
    data = gen_text(mb * 1_000_000, 8)
    comp = eng.compressFile(data, None, 9)
    cp = torch.frombuffer(bytearray(comp), dtype=torch.uint8).pin_memory()

    def timed(fn, reps=4):
        fn(); fn()
        t0 = time.time()
        for _ in range(reps):
            fn()
        return (time.time() - t0) / reps * 1e3

    def single():
        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        rc = L.bz2b200_decompress(eng._ctx, cp.data_ptr(), len(comp), 0, C.byref(out), C.byref(n))
        assert rc == 0 and n.value == data.size
        L.bz2b200_free(out)
    print(f"{mb} MB: single context {timed(single):.2f} ms", flush=True)
    for lanes in (1, 2, 3):
        pool = Bzip2Pool([0], lanes)
        for sl in (0, 8, 16, 33, 66):
            def f():
                p, n = pool.decompress_raw(cp.data_ptr(), len(comp), False, data.size, int(sl * 1e6))
                assert n == data.size
                pool.free_raw(p)
            print(f"{mb} MB: pool lanes={lanes} slice {sl} MB: {timed(f):.2f} ms", flush=True)
        pool.close()
