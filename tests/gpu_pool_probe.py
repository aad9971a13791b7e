"""Development aid (GPU): the shard scheduler on 100 MB of bench text -- parity with the golden SHA and timings of the
host-buffer call for several lane / shard settings, pinned and pageable input.
    python tests/gpu_pool_probe.py [mb] [level]"""
import ctypes as C
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from compressjs_flattened_b200 import Bzip2Engine  # noqa: E402
from compressjs_flattened_b200.corpus import gen_text  # noqa: E402
from compressjs_flattened_b200.pool import Bzip2Pool  # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 100
level = int(sys.argv[2]) if len(sys.argv) > 2 else 9
n = mb * 1_000_000
GOLD = json.load(open(os.path.join(HERE, "golden", "corpus_goldens.json")))
data = [gen_text(n, 8), gen_text(n, 8, first_chunk=mb)]
pinned = [torch.from_numpy(d).pin_memory() for d in data]
gold = GOLD.get(f"text:{n}:8:L{level}")
eng = Bzip2Engine(0)
L = eng._L


def timed(fn, reps=6):
    fn(0)
    fn(1)
    torch.cuda.synchronize()
    t0 = time.time()
    for i in range(reps):
        fn(i)
    torch.cuda.synchronize()
    return (time.time() - t0) / reps * 1e3


def host_call(ptrs):
    def f(i):
        out, ln = C.POINTER(C.c_uint8)(), C.c_size_t()
        rc = L.bz2b200_compress(eng._ctx, ptrs[i % 2], n, level, C.byref(out), C.byref(ln))
        assert rc == 0, (rc, L.bz2b200_last_error(eng._ctx))
        L.bz2b200_free(out)
    return f


pin_ptrs = [p.data_ptr() for p in pinned]
pag_ptrs = [d.ctypes.data for d in data]
# parity first
got = eng.compressFile(data[0], None, level)
print("pool path sha ok:", gold is None or hashlib.sha256(got).hexdigest() == gold["out_sha256"], len(got), "blocks", eng.stats().n_blocks, flush=True)
d_in = [p.cuda() for p in pinned]
d_out = torch.empty(eng.compress_bound(n, level) + 64, dtype=torch.uint8, device="cuda")
t = timed(lambda i: eng.compress_device(d_in[i % 2].data_ptr(), n, level, d_out.data_ptr(), d_out.numel()))
print(f"device-resident: {t:.2f} ms  ({mb / t:.2f} GB/s); stats ms_total {eng.stats().ms_total:.2f}", flush=True)
eng.debug_set_pool(1 << 62)
print(f"single path, pinned:   {timed(host_call(pin_ptrs)):.2f} ms", flush=True)
print(f"single path, pageable: {timed(host_call(pag_ptrs)):.2f} ms", flush=True)
for shard_mb in (0, 6, 10, 12.5, 17, 25, 34, 50):
    eng.debug_set_pool(0, int(shard_mb * 1e6))
    t1 = timed(host_call(pin_ptrs))
    t2 = timed(host_call(pag_ptrs))
    print(f"ctx pool (2 lanes) shard {shard_mb} MB: pinned {t1:.2f} ms, pageable {t2:.2f} ms", flush=True)
for lanes in (1, 2, 3, 4):
    pool = Bzip2Pool([0], lanes)
    for shard_mb in (0, 12.5, 25):
        def f(i, ptrs=pin_ptrs):
            p, ln = pool.compress_raw(ptrs[i % 2], n, level, int(shard_mb * 1e6))
            pool.free_raw(p)
        t1 = timed(f)
        t2 = timed(lambda i: f(i, pag_ptrs))
        print(f"pool lanes={lanes} shard {shard_mb} MB: pinned {t1:.2f} ms ({mb / t1:.2f} GB/s), pageable {t2:.2f} ms", flush=True)
    got = pool.compressFile(data[0], None, level)
    print("   sha ok:", gold is None or hashlib.sha256(got).hexdigest() == gold["out_sha256"], flush=True)
    pool.close()
ng = torch.cuda.device_count()
if ng > 1:
    for lanes in (1, 2):
        pool = Bzip2Pool(list(range(ng)), lanes)
        big = gen_text(n * ng, 8)
        bp = torch.from_numpy(big).pin_memory()
        def f(i):
            p, ln = pool.compress_raw(bp.data_ptr(), n * ng, level, 0)
            pool.free_raw(p)
        t1 = timed(f, 4)
        print(f"{ng} GPUs in one process, lanes={lanes}: {t1:.2f} ms ({mb * ng / t1:.2f} GB/s)", flush=True)
        got = pool.compressFile(big, None, level)
        g2 = GOLD.get(f"text:{n * ng}:8:L{level}")
        print("   sha ok:", g2 is None or hashlib.sha256(got).hexdigest() == g2["out_sha256"], len(got), flush=True)
        pool.close()
