#!/usr/bin/env python
"""bench.py -- bzip2 level-9 compress throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--level 9] [--mb 100]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path (Bzip2.compressFile at level 9) over one 100 MB batch of
synthetic enwik8-like text per GPU (BASELINE.json configs[1]; at N > 1 every rank takes its own
100 MB block-range shard of the N x 100 MB corpus: weak scaling, no collective on the data path).

  value     whole-job MB/s (MB = 1e6 uncompressed input bytes) with the input resident in HBM,
            timed on the device with CUDA events on the launching stream, max over ranks
  e2e       same metric through the reference-facing host API (Bzip2.compressFile over host
            buffers -> bz2b200_compress): pinned host input, H2D + kernels + D2H all inside the
            timed region
  roofline  the dominant kernel (k_rs_scatter, the radix-sort scatter of the BWT stage): algorithmic
            bytes / live CUDA-event time vs the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle (a C port of the reference's algorithm, oracle/) on the host cores,
            on a bounded sample of the same workload (rank 0, N = 1 only)

--impl reference times that CPU port with all host threads on the same config/metric.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "bzip2-9 compress MB/s, byte-identical output"
UNIT = "MB/s"


def workload_desc(level, mb, n_gpus):
    return {"workload": f"bzip2 level {level} ({level * 100} KB blocks) on {mb} MB synthetic enwik8-like text per GPU "
                        f"(compressjs_flattened_b200.corpus.gen_text, seed 8; BASELINE.json configs[1])",
            "level": level, "bytes_per_gpu": mb * 1_000_000, "parallelism": f"block-range shards x{n_gpus} of one {mb * n_gpus} MB stream; cross-rank: first-block offset chain + bit-length exscan (scalars), no data-path collective",
            "l2": "two distinct 100 MB input buffers alternate between steps (200 MB > 126 MB L2); per-call device state is 5.3 GB"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe): clocks.sm,
    clocks.max.sm and the clocks_event_reasons.active bitmask (decoded below), every 100 ms.  Measured beside this
    bench (tests/gpu_sampler_probe.py) the loop costs nothing as long as the timed path makes no driver-lock calls:
    a per-call cudaMemGetInfo in the library used to collide with it (+3..24 ms per step) and was removed."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake"}

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = "index,clocks.sm,clocks.max.sm,clocks_event_reasons.active"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for t, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 4 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
                mask = int(f[3], 16)
            except ValueError:
                continue
            for bit, name in self.REASONS.items():
                if mask & bit:
                    reasons.add(name)
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}
        if not sm and self.lines:  # a window shorter than the sampling period: report the sample nearest to it
            t, line = min(self.lines, key=lambda x: abs(x[0] - (t0 + t1) / 2))
            f = [x.strip() for x in line.split(",")]
            try:
                out.update({"sm_mhz": float(f[1]), "sm_max_mhz": float(f[2]), "reasons": sorted(n for b, n in self.REASONS.items() if int(f[3], 16) & b),
                            "samples": 1, "note": f"nearest sample, {abs(t - (t0 + t1) / 2) * 1e3:.0f} ms from the middle of the timed region"})
            except (ValueError, IndexError):
                pass
        return out


# dram__bytes_read.sum + dram__bytes_write.sum of one k_rs_scatter<9> launch over 100 M keys (first pass of a bench step),
# `ncu --set full`, profiles/r01_ncu_notes.md; the algorithmic figure for that launch is 1.6e9 bytes
NCU_TRAFFIC_PER_LAUNCH = 1602111488
NCU_TRAFFIC_NOTE = "ncu --set full, first scatter pass of a step (100 M keys, 1.6e9 algorithmic bytes): 850.6 MB read + 751.5 MB written"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def cpu_baseline(level, sample_mb, threads):
    import oracle_binding as O
    from compressjs_flattened_b200.corpus import gen_text
    data = gen_text(sample_mb * 1_000_000, 8)
    t0 = time.time()
    out = O.compress(data, level, O.SORT_STABLE, threads=threads)
    dt = time.time() - t0
    return {"value": round(sample_mb / dt, 3), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {sample_mb} MB of the workload, oracle/liboracle.so (C port of Bzip2_joined_.js: SA-IS BWT, same table optimiser), "
                      f"block-parallel over {threads} pthreads, {dt:.1f} s; the reference itself is single-threaded JavaScript and Node is not installed here",
            "out_bytes": len(out)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_binding as O
    from compressjs_flattened_b200.corpus import gen_text
    threads = os.cpu_count() or 1
    sample_mb = args.ref_mb
    data = gen_text(sample_mb * 1_000_000, 8)
    for _ in range(min(args.warmup, 1)):
        O.compress(data[:5_000_000], args.level, threads=threads)
    times = []
    for _ in range(args.steps):
        t0 = time.time()
        O.compress(data, args.level, threads=threads)
        times.append(time.time() - t0)
    ms = 1e3 * sum(times) / len(times)
    val = sample_mb / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_desc(args.level, args.mb, args.gpus),
            "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"each step = first {sample_mb} MB of the workload through oracle/liboracle.so on {threads} pthreads "
                                       "(Node is absent, so the reference's own JS cannot run; the oracle is its C restatement)"},
            "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--level", type=int, default=9)
    ap.add_argument("--mb", type=int, default=100)
    ap.add_argument("--ref-mb", type=int, default=40)
    ap.add_argument("--cpu-sample-mb", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="device-resident compress steps only (for ncu runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from compressjs_flattened_b200 import Bzip2Engine
    from compressjs_flattened_b200.corpus import gen_text

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctl = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ctl = dist.new_group(backend="gloo")   # the per-rank scalars (first-block offset, bit length) travel host-side
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()   # started early (corpus generation takes seconds): it is streaming well before the timed region
    eng = Bzip2Engine(local)
    from compressjs_flattened_b200.sharded import HostMailbox, compress_shard, gather_and_stitch
    # one node: the three scalars per rank go through shared memory (microseconds) instead of the TCP-backed group
    mbox = HostMailbox(rank, world, os.environ.get("MASTER_PORT", "0")) if world > 1 else None
    nbytes = args.mb * 1_000_000
    chunks = args.mb
    halo_mb = 2 if (world > 1 and rank < world - 1) else 0   # bytes after the slice that the last owned block may need
    # two distinct corpora (j) so consecutive steps never re-read a cached input (200 MB > L2); rank r owns
    # bytes [r*mb MB, (r+1)*mb MB) of corpus j = one block-range shard of a world*mb MB stream
    host = [gen_text(nbytes + halo_mb * 1_000_000, 8, first_chunk=j * world * chunks + rank * chunks) for j in range(2)]
    pinned = [torch.from_numpy(h).pin_memory() for h in host]
    d_in = [p.to(dev) for p in pinned]
    navail = nbytes + halo_mb * 1_000_000
    bound = eng.compress_bound(nbytes, args.level)
    d_out = torch.empty((bound + 3) // 4 * 4, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def dev_step(i):
        if world == 1:
            return eng.compress_device(d_in[i % 2].data_ptr(), nbytes, args.level, d_out.data_ptr(), d_out.numel())
        n, _, _ = compress_shard(eng, None, rank * nbytes, nbytes, args.level, rank == world - 1, rank=rank, world=world, group=ctl,
                                 device_ptr=d_in[i % 2].data_ptr(), nbytes=navail, to_host=False, mailbox=mbox)
        return n

    # ---------------- device-resident: `value` ----------------
    for i in range(args.warmup):
        out_len = dev_step(i)
    barrier()
    t0 = time.time()
    dev_ms, dom_ms, dom_bytes, dom_launches, launches, stage = 0.0, 0.0, 0, 0, 0, [0.0] * 5
    for i in range(args.steps):
        out_len = dev_step(i)
        st = eng.stats()
        dev_ms += st.ms_total
        dom_ms += st.dom_ms
        dom_bytes += st.dom_bytes
        dom_launches += st.dom_launches
        launches += st.kernel_launches
        for k in range(5):
            stage[k] += st.ms_stage[k]
    barrier()
    t1 = time.time()
    if world > 1:
        print(f"[bench] rank {rank}: speculated shard starts held / recut = {compress_shard.guesses[True]} / {compress_shard.guesses[False]}", file=sys.stderr, flush=True)
    wall_ms = (t1 - t0) * 1e3
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    st_last = eng.stats()

    if args.profile_only:
        print(json.dumps({"profile_only": True, "ms_per_step": round(dev_ms / args.steps, 3), "launches_per_step": launches // args.steps}), flush=True)
        return
    # ---------------- end to end through the host API: `e2e` ----------------
    L = eng._L
    out_p, out_n = ctypes.POINTER(ctypes.c_uint8)(), ctypes.c_size_t()

    def host_call(j):
        if world > 1:  # host slice+halo in, host segment out, through the shard API
            (ptr, n), _, _ = compress_shard(eng, pinned[j].numpy(), rank * nbytes, nbytes, args.level, rank == world - 1, rank=rank, world=world,
                                            group=ctl, to_host="raw", mailbox=mbox)
            eng.free_raw(ptr)
            return n
        rc = L.bz2b200_compress(eng._ctx, pinned[j].data_ptr(), nbytes, args.level, ctypes.byref(out_p), ctypes.byref(out_n))
        if rc:
            eng._raise(rc)
        n = out_n.value
        L.bz2b200_free(out_p)
        return n

    host_call(0)
    barrier()
    e0 = time.time()
    e2e_out = 0
    for i in range(args.steps):
        e2e_out = host_call(i % 2)
    barrier()
    e2e_ms = (time.time() - e0) * 1e3

    # ---------------- N > 1: stitch the ranks' segments into ONE stream and check it ----------------
    stitched = None
    if world > 1:
        seg, info, _ = compress_shard(eng, pinned[0].numpy(), rank * nbytes, nbytes, args.level, rank == world - 1, rank=rank, world=world, group=ctl)
        whole = gather_and_stitch(eng, seg, info, args.level, group=ctl)
        if rank == 0:
            back = eng.decompressFile(whole)   # verifies every block CRC and the combined CRC
            stitched = {"bytes": len(whole), "decoded_bytes": len(back), "blocks": int(eng.stats().n_blocks),
                        "crc_checked_roundtrip": len(back) == world * nbytes and back[:nbytes] == host[0][:nbytes].tobytes()}
            del back, whole
        # per-rank decompress below works on an ordinary single-rank stream of this rank's slice
        out_len = eng.compress_device(d_in[0].data_ptr(), nbytes, args.level, d_out.data_ptr(), d_out.numel())
        last_in = 0
    else:
        last_in = (args.steps - 1) % 2
    # ---------------- decompress (reported alongside) ----------------
    comp = torch.empty(out_len, dtype=torch.uint8, device=dev)
    comp.copy_(d_out[:out_len])
    d_back = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
    dec_steps = max(1, min(args.steps, 3))
    eng.decompress_device(comp.data_ptr(), out_len, False, d_back.data_ptr(), d_back.numel())
    dec_ms = 0.0
    for i in range(dec_steps):
        got = eng.decompress_device(comp.data_ptr(), out_len, False, d_back.data_ptr(), d_back.numel())
        dec_ms += eng.stats().ms_total
    roundtrip_ok = bool(got == nbytes and torch.equal(d_back[:nbytes], d_in[last_in][:nbytes]))

    # ---------------- reduce over ranks ----------------
    t = torch.tensor([dev_ms, wall_ms, e2e_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max, e2e_ms_max, dec_ms_max = [float(x) for x in t.tolist()]
    if rank == 0:
        peak, peak_src = peaks()
        total_mb = world * args.mb * args.steps
        achieved = dom_bytes / 1e9 / (dom_ms / 1e3) if dom_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": round(total_mb / (dev_ms_max / 1e3), 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dev_ms_max / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_desc(args.level, args.mb, world),
            "e2e": {"value": round(total_mb / (e2e_ms_max / 1e3), 2), "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(e2e_out),
                    "ms_per_step": round(e2e_ms_max / args.steps, 3), "api": "bz2b200_compress (N=1) / bz2b200_shard_begin..emit (N>1): pinned host input, page-locked host output from the library's result pool"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_rs_scatter<9>/<8> (BWT radix-sort scatter passes)", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": NCU_TRAFFIC_PER_LAUNCH, "traffic_note": NCU_TRAFFIC_NOTE, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(dom_bytes / max(dom_launches, 1)), "launches_per_step": dom_launches // args.steps,
                         "avg_launch_ms": round(dom_ms / max(dom_launches, 1), 4), "share_of_step": round(dom_ms / dev_ms, 3) if dev_ms else None,
                         "pipeline_algorithmic_GBps": round((1 + 12 * st_last.rle1_bytes / nbytes + 22 * st_last.mtf_syms / nbytes + 3 * out_len / nbytes)
                                                            * nbytes / 1e9 / (dev_ms / args.steps / 1e3), 1)},
            "stage_ms_per_step": {k: round(v / args.steps, 3) for k, v in zip(("rle1_cut_crc", "bwt", "mtf_rle2", "huffman_emit", "stitch"), stage)},
            "wall_ms_per_step": round(wall_ms_max / args.steps, 3),
            "clocks": clocks,
            "decompress": {"value": round(world * args.mb * dec_steps / (dec_ms_max / 1e3), 2), "unit": UNIT, "ms_per_step": round(dec_ms_max / dec_steps, 3),
                           "roundtrip_bit_exact": roundtrip_ok},
            "stitched_stream": stitched,
            "out_bytes_per_step": int(out_len), "blocks_per_step": int(st_last.n_blocks), "sort_rounds": int(st_last.sort_rounds),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.level, args.cpu_sample_mb, os.cpu_count() or 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
