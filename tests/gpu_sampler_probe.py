"""Development aid: how much does a clock sampler beside the run cost?  torchrun, N GPUs of one node."""
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import torch
import torch.distributed as dist

from compressjs_flattened_b200 import Bzip2Engine
from compressjs_flattened_b200.corpus import gen_text
from compressjs_flattened_b200.sharded import HostMailbox, compress_shard

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Bzip2Engine(local)
mb = HostMailbox(rank, world, os.environ.get("MASTER_PORT", "0")) if world > 1 else None
nbytes = 100_000_000
halo = 2_000_000 if rank < world - 1 else 0
d = torch.from_numpy(gen_text(nbytes + halo, 8, first_chunk=rank * 100)).to(local)
out = torch.empty(eng.compress_bound(nbytes, 9) + 8, dtype=torch.uint8, device=local)
FULL = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
LITE = "index,clocks.sm,clocks.max.sm,clocks_event_reasons.active"
MODES = [("none", None, 0), ("smi100_full", FULL, 100), ("smi100_lite", LITE, 100), ("smi500_full", FULL, 500), ("smi1000_full", FULL, 1000), ("none2", None, 0)]


def step():
    if world == 1:
        eng.compress_device(d.data_ptr(), nbytes, 9, out.data_ptr(), out.numel())
    else:
        compress_shard(eng, None, rank * nbytes, nbytes, 9, rank == world - 1, rank=rank, world=world, device_ptr=d.data_ptr(), nbytes=nbytes + halo,
                       to_host=False, mailbox=mb)


for _ in range(3):
    step()
for name, q, ms in MODES:
    proc = None
    if q and rank == 0:
        proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", str(ms), "-i", str(local)],
                                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 40
    for _ in range(K):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) * 1e3 / K
    if proc:
        proc.terminate()
    if rank == 0:
        print(f"{name:14s} {dt:7.2f} ms/step", flush=True)
if mb:
    mb.close()
    dist.destroy_process_group()
