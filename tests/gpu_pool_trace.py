"""Development aid (GPU): BZ2B200_TRACE=1 python tests/gpu_pool_trace.py [lanes] [shard_mb] -- per-kernel event trace of every shard."""
import os
import sys
import time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import torch  # noqa: E402
from compressjs_flattened_b200.corpus import gen_text  # noqa: E402
from compressjs_flattened_b200.pool import Bzip2Pool  # noqa: E402
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
shard = float(sys.argv[2]) if len(sys.argv) > 2 else 12.5
n = 100_000_000
d = torch.from_numpy(gen_text(n, 8)).pin_memory()
pool = Bzip2Pool([0], lanes)
for i in range(3):
    if i == 2:
        print("==== traced call ====", file=sys.stderr, flush=True)
    t0 = time.time()
    p, ln = pool.compress_raw(d.data_ptr(), n, 9, int(shard * 1e6))
    pool.free_raw(p)
    print(f"call {i}: {(time.time() - t0) * 1e3:.2f} ms", file=sys.stderr, flush=True)
