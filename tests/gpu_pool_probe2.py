"""Development aid (GPU): shard plans (first shard, growth, lanes) for the host-buffer call on 100 MB."""
import ctypes as C
import os
import sys
import time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import torch  # noqa: E402
from compressjs_flattened_b200.corpus import gen_text  # noqa: E402
from compressjs_flattened_b200.pool import Bzip2Pool, shard_plan  # noqa: E402
n = 100_000_000
level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
data = [gen_text(n, 8), gen_text(n, 8, first_chunk=100)]
pinned = [torch.from_numpy(d).pin_memory() for d in data]


def timed(fn, reps=8):
    fn(0); fn(1)
    torch.cuda.synchronize()
    t0 = time.time()
    for i in range(reps):
        fn(i)
    return (time.time() - t0) / reps * 1e3


for lanes in (1, 2, 3):
    pool = Bzip2Pool([0], lanes)
    for first_mb, growth in ((0, 0), (4, 3), (6, 4), (8, 12), (10, 2), (3, 3), (12, 8), (50, 1)):
        pool.set_plan(int(first_mb * 1e6), growth)
        plan = shard_plan(n, level, lanes, int(first_mb * 1e6), growth)
        def f(i, ptrs=None):
            src = pinned[i % 2].data_ptr() if ptrs is None else ptrs[i % 2]
            p, ln = pool.compress_raw(src, n, level, 0)
            pool.free_raw(p)
        t1 = timed(f)
        t2 = timed(lambda i: f(i, [d.ctypes.data for d in data]))
        print(f"lanes={lanes} first={first_mb} growth={growth} plan={[round(x / 1e6, 1) for x in plan]}: pinned {t1:.2f} ms, pageable {t2:.2f} ms", flush=True)
    pool.close()
