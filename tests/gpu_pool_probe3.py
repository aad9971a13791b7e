"""Development aid (GPU): two-shard plans [F, n-F] on two lanes for the host-buffer call on 100 MB."""
import os
import sys
import time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import torch  # noqa: E402
from compressjs_flattened_b200.corpus import gen_text  # noqa: E402
from compressjs_flattened_b200.pool import Bzip2Pool  # noqa: E402
n = 100_000_000
data = [gen_text(n, 8), gen_text(n, 8, first_chunk=100)]
pinned = [torch.from_numpy(d).pin_memory() for d in data]


def timed(fn, reps=10):
    fn(0); fn(1)
    torch.cuda.synchronize()
    t0 = time.time()
    for i in range(reps):
        fn(i)
    return (time.time() - t0) / reps * 1e3


for lanes in (2, 1):
    pool = Bzip2Pool([0], lanes)
    for f_mb in (10, 15, 20, 25, 30, 40, 50):
        pool.set_plan(int(f_mb * 1e6), 1000.0)
        def f(i, ptrs=None):
            src = pinned[i % 2].data_ptr() if ptrs is None else ptrs[i % 2]
            p, ln = pool.compress_raw(src, n, 9, 0)
            pool.free_raw(p)
        t1 = timed(f)
        t2 = timed(lambda i: f(i, [d.ctypes.data for d in data]))
        print(f"lanes={lanes} plan=[{f_mb}, {100 - f_mb}]: pinned {t1:.2f} ms, pageable {t2:.2f} ms", flush=True)
    pool.close()
