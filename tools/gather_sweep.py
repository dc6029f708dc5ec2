#!/usr/bin/env python
"""Developer tool (torchrun, one rank per GPU): C3 step with every rank's hits gathered on every rank, through
trq_trace_gather (peer stores issued by the traversal kernel as each ray retires; 32-byte and 16-byte records) and
through NCCL (dist.trace_and_gather)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tracer_b200 import Scene, dist as D, harness as H, layout as L, rays_to_torch  # noqa: E402

rank, local_rank, world = D.init()
dev = f"cuda:{local_rank}"
prim = D.replicate_primitive(H.scene_c3(2) if rank == 0 else None)
scene = Scene(prim, local_rank)
d = rays_to_torch(H.cornell_camera_rays(3840, 2160), dev)
recs = scene.expand(d, scene.hit(d)).cpu().numpy().view(L.record_dtype).reshape(-1)
bounce, _ = H.bounce_rays(recs, rank << 32)
n = int(-D.max_over_ranks(-float(bounce.size)))
rays = rays_to_torch(bounce[:n], dev)
steps = 30


def nvlink_kib():
    """Cumulative NVLink payload counters of this rank's GPU (NVML field values, KiB): (tx, rx), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        vals = pynvml.nvmlDeviceGetFieldValues(h, [pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX])
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                return None
            out.append(int(v.value.ullVal))
        return tuple(out)
    except Exception:
        return None


def timed(fn):
    for _ in range(3):
        fn()
    D.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return D.max_over_ranks(e0.elapsed_time(e1)) / steps


out = torch.empty((n, 8), dtype=torch.float32, device=dev)
ms = timed(lambda: scene.hit(rays, out=out))
if rank == 0:
    print(f"world {world}  n/rank {n}  no gather: {ms:.3f} ms  {n * world / ms / 1e3:.0f} Mrays/s", flush=True)
hg = D.HitGather(scene, n)
for h16 in (False, True):
    c0 = nvlink_kib()
    ms = timed(lambda: (hg.trace(rays, hit16=h16), hg.wait()))
    c1 = nvlink_kib()
    hg.status()
    if rank == 0:
        rec = 16 if h16 else 32
        line = (f"trq_trace_gather, {rec}-byte records: {ms:.3f} ms  {n * world / ms / 1e3:.0f} Mrays/s  "
                f"(NVLink egress per GPU {n * (world - 1) * rec / ms / 1e6:.0f} GB/s)")
        if c0 and c1:                                         # 3 warm-up + `steps` timed steps between the two readings
            per_step_tx = (c1[0] - c0[0]) * 1024 / (steps + 3); per_step_rx = (c1[1] - c0[1]) * 1024 / (steps + 3)
            line += (f"  [NVML NVLink counters, rank 0: tx {per_step_tx / 1e6:.1f} MB, rx {per_step_rx / 1e6:.1f} MB per step; "
                     f"payload expected {n * (world - 1) * rec / 1e6:.1f} MB each way]")
        print(line, flush=True)
hg.close()
g_all = torch.empty((world, n, 8), dtype=torch.float32, device=dev)
ms = timed(lambda: D.trace_and_gather(scene, rays, out, g_all))
if rank == 0:
    print(f"NCCL all-gather, 4 chunks overlapped: {ms:.3f} ms  {n * world / ms / 1e3:.0f} Mrays/s", flush=True)
D.barrier()
