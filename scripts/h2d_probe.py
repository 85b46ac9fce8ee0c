"""Probe: pinned host -> device copy bandwidth as a function of the CPU cores the process (and its pinned allocation)
sits on.  Prints nvidia-smi topo and the GB/s of 128 MB copies for a few affinity masks."""
import os
import subprocess
import sys

import torch

print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
try:
    print(open("/sys/devices/system/node/online").read().strip(), "numa nodes online")
    for nd in sorted(os.listdir("/sys/devices/system/node")):
        if nd.startswith("node"):
            print(nd, open(f"/sys/devices/system/node/{nd}/cpulist").read().strip())
except OSError as e:
    print("no numa info", e)
allowed = sorted(os.sched_getaffinity(0))
print("allowed cpus:", len(allowed), allowed[:4], "...", allowed[-4:])


def bw(nbytes=128 << 20, iters=10):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return nbytes * iters / e0.elapsed_time(e1) / 1e6


torch.cuda.init()
print("default affinity: %.1f GB/s" % bw())
n = len(allowed)
for name, cpus in (("first quarter", allowed[: n // 4]), ("second quarter", allowed[n // 4: n // 2]),
                   ("third quarter", allowed[n // 2: 3 * n // 4]), ("last quarter", allowed[3 * n // 4:])):
    if not cpus:
        continue
    os.sched_setaffinity(0, cpus)
    print(f"{name} ({cpus[0]}..{cpus[-1]}): {bw():.1f} GB/s")
os.sched_setaffinity(0, allowed)
for sz in (16, 64, 256):
    print(f"{sz} MB copies: {bw(sz << 20):.1f} GB/s")
