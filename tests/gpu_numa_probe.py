"""Development aid (not a test): host topology of the GPU box and what thread placement does to pinned-copy bandwidth.

python tests/gpu_numa_probe.py           prints NUMA nodes, the GPU's node, and H2D / D2H GB/s of 256 MB pinned copies with the
                                         calling thread (and therefore the first-touch placement of the pinned buffer) on each node.
"""
import glob
import os
import subprocess
import sys

import torch


def cpulist(s):
    out = set()
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out


def main():
    mask0 = os.sched_getaffinity(0)
    print("cpus allowed:", len(mask0), "of", os.cpu_count())
    nodes = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        k = int(d.rsplit("node", 1)[1])
        nodes[k] = cpulist(open(d + "/cpulist").read())
        print("node", k, "cpus", len(nodes[k]), "allowed", len(nodes[k] & mask0))
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout)
    except Exception as e:  # noqa: BLE001
        print("nvidia-smi topo failed:", e)
    for i in range(torch.cuda.device_count()):
        p = torch.cuda.get_device_properties(i)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        try:
            node = open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip()
        except OSError as e:
            node = "? (%s)" % e
        print("gpu", i, bus, "numa_node", node)
    n = 256 << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    placements = [("unbound", mask0)] + [("node%d" % k, c & mask0) for k, c in nodes.items() if c & mask0]
    for name, cpus in placements:
        os.sched_setaffinity(0, cpus)
        host = torch.empty(n, dtype=torch.uint8).pin_memory()
        host.fill_(1)
        for label, fn in (("h2d", lambda: dev.copy_(host, non_blocking=True)), ("d2h", lambda: host.copy_(dev, non_blocking=True))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                fn()
            b.record()
            torch.cuda.synchronize()
            print("%-8s %s %.1f GB/s" % (name, label, 10 * n / a.elapsed_time(b) / 1e6), flush=True)
        del host
    os.sched_setaffinity(0, mask0)


if __name__ == "__main__":
    sys.exit(main())
