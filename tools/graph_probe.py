#!/usr/bin/env python
"""Developer tool: does capturing the launch-bound workloads (C1: one 72 us trace; C2: cast -> trace -> spawn -> trace) in a CUDA
graph pay? Runs K steps eagerly and as K replays of one captured step; results must be identical."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tracer_b200 import Scene, harness as H, rays_to_torch  # noqa: E402

dev = "cuda:0"
K = 200


def timed(fn, k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def run(name, step, n, outs):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.synchronize()
    ref = [o.clone() for o in outs]
    eager = timed(step, K)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step()
    for o in outs:
        o.zero_()
    g.replay(); torch.cuda.synchronize()
    same = all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(ref, outs))
    graph = timed(g.replay, K)
    print(f"{name}: eager {eager * 1e3:.1f} us/step = {n / eager / 1e3:.0f} Mrays/s; graph replay {graph * 1e3:.1f} us/step = {n / graph / 1e3:.0f} Mrays/s; same results {same}", flush=True)


prim = H.scene_c1(); scene = Scene(prim, 0)
rays = rays_to_torch(H.camera_rays((13, 2, 3), (0, 0, 0), np.float32(20 * np.pi / 180), 1280, 720), dev)
hits = torch.empty((rays.shape[0], 8), dtype=torch.float32, device=dev)
run("c1", lambda: scene.hit(rays, out=hits), rays.shape[0], [hits])

prim2 = H.scene_c2(); s2 = Scene(prim2, 0)
W, Hh = 1920, 1080
cam = ((278, 278, -800), (278, 278, 278), (0, 1, 0), np.float32(45 * (np.pi / 180)))
r0 = torch.empty((W * Hh, 8), dtype=torch.float32, device=dev); h0 = torch.empty_like(r0); r1 = torch.empty_like(r0); h1 = torch.empty_like(r0)
s1 = torch.empty(W * Hh, dtype=torch.int32, device=dev); c1 = torch.zeros(1, dtype=torch.int64, device=dev)


def wave():
    s2.cast_rays(*cam, W, Hh, out=r0)
    s2.hit(r0, out=h0)
    s2.spawn_bounce(r0, h0, seed_base=0, out=r1, src=s1, count=c1)
    s2.hit_indirect(r1, c1, out=h1)


wave(); torch.cuda.synchronize()
n = W * Hh + int(c1.item())
run("c2 wavefront", wave, n, [h0, c1])
