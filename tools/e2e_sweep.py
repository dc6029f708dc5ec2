#!/usr/bin/env python
"""Developer tool: end-to-end rate of the host-pointer path (pinned host rays -> trq_trace(TRQ_HOST_PTRS) -> host hits)
on C3 for a list of TRQ_CHUNK_RAYS values. Usage: python tools/e2e_sweep.py [chunk ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tracer_b200 import Scene, harness as H, layout as L, rays_to_torch  # noqa: E402

chunks = [int(x) for x in sys.argv[1:]] or [32768, 65536, 131072, 262144, 524288, 1048576]
prim = H.scene_c3(2)
scene = Scene(prim, 0)
d = rays_to_torch(H.cornell_camera_rays(3840, 2160), "cuda:0")
recs = scene.expand(d, scene.hit(d)).cpu().numpy().view(L.record_dtype).reshape(-1)
bounce, _ = H.bounce_rays(recs)
n = bounce.size
h_rays = torch.from_numpy(bounce.view(np.float32).reshape(-1, 8)).pin_memory()
h_hits = torch.empty((n, 8), dtype=torch.float32).pin_memory()
want = None
for c in chunks:
    os.environ["TRQ_CHUNK_RAYS"] = str(c)
    for _ in range(3):
        scene.hit_host(h_rays.data_ptr(), n, h_hits.data_ptr())
    ts = []
    for _ in range(15):
        t = time.perf_counter(); scene.hit_host(h_rays.data_ptr(), n, h_hits.data_ptr()); ts.append(time.perf_counter() - t)
    if want is None:
        want = h_hits.clone()
    assert torch.equal(want, h_hits)
    ms = float(np.median(ts)) * 1e3
    ti = []
    for _ in range(5):                                           # how long the host takes just to issue the chunks
        t = time.perf_counter(); scene.hit_host(h_rays.data_ptr(), n, h_hits.data_ptr(), asynchronous=True)
        ti.append(time.perf_counter() - t); scene.host_sync()
    print(f"issue {float(np.median(ti)) * 1e3:.3f} ms ", end="")
    print(f"{c:8d}  {n / ms / 1e3:8.1f} Mrays/s  {ms:.3f} ms  (min {min(ts) * 1e3:.3f})", flush=True)
