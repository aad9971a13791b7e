"""Development aid: wall-clock trace of the shard protocol per rank (torchrun, N GPUs of one node)."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import torch
import torch.distributed as dist

from compressjs_flattened_b200 import Bzip2Engine
from compressjs_flattened_b200.corpus import gen_text
from compressjs_flattened_b200.sharded import HostMailbox

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Bzip2Engine(local)
mb = HostMailbox(rank, world, os.environ.get("MASTER_PORT", "0"))
nbytes = 100_000_000
halo = 2_000_000 if rank < world - 1 else 0
h = gen_text(nbytes + halo, 8, first_chunk=rank * 100)
d = torch.from_numpy(h).to(local)
torch.cuda.synchronize()
base = rank * nbytes
for step in range(6):
    dist.barrier()
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    eng.shard_begin(None, 9, device_ptr=d.data_ptr(), nbytes=nbytes + halo); t.append(time.perf_counter())
    mb.next_round()
    start_v = mb.get(rank - 1, 0) if rank > 0 else 0; t.append(time.perf_counter())
    info = eng.shard_cut(max(start_v - base, 0), nbytes, rank == world - 1); t.append(time.perf_counter())
    mb.put(0, max(base + int(info.next_start), start_v))
    eng.shard_compress(info); t.append(time.perf_counter())
    mb.put(1, int(info.bits))
    allb = [mb.get(r, 1) for r in range(world)]; t.append(time.perf_counter())
    n = eng.shard_emit(info, (32 + sum(allb[:rank])) & 7, to_host=False); t.append(time.perf_counter())
    st = eng.stats()
    names = ["begin", "wait_chain", "cut", "compress", "wait_bits", "emit"]
    print(f"rank{rank} step{step} " + " ".join(f"{nm}={1e3 * (b - a):.2f}" for nm, a, b in zip(names, t, t[1:])) +
          f" | dev stages {[round(x, 2) for x in st.ms_stage[:5]]} total {st.ms_total:.2f}", flush=True)
mb.close()
dist.destroy_process_group()
