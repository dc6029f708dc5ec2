#!/usr/bin/env python
"""Developer tool (torchrun, one rank per GPU): pinned H2D / D2H / bidirectional copy bandwidth with EVERY rank copying at
the same time -- what the host side (PCIe switches, root complex, host DRAM) gives N GPUs at once. Names the limiter of the
multi-GPU e2e figure. Prints per-rank and aggregate GB/s per direction."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tracer_b200 import dist as D  # noqa: E402

rank, local_rank, world = D.init()
torch.cuda.set_device(local_rank)
n = 197_166_528
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=10):
    D.barrier(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return D.max_over_ranks((time.perf_counter() - t) / reps)


try:
    aff = sorted(os.sched_getaffinity(0))
    numa = open(f"/sys/bus/pci/devices/{torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), 'pci_bus_id') else ''}/numa_node").read().strip()
except Exception:
    aff, numa = [], "?"
for name, a, b in (("h2d", 1, 0), ("d2h", 0, 1), ("both", 1, 1)):
    run(a, b, 2); dt = run(a, b)
    if rank == 0:
        print(f"{world} GPUs at once, {name}: {dt * 1e3:.2f} ms per 197 MB per GPU -> {n / dt / 1e9:.1f} GB/s per GPU per direction, "
              f"{world * n / dt / 1e9:.1f} GB/s aggregate per direction", flush=True)
if rank == 0:
    print(f"host: {os.cpu_count()} cpus visible, rank 0 affinity {len(aff)} cpus")
D.barrier()
