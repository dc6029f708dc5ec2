#!/usr/bin/env python
"""Developer tool: time host vs GPU BVH build. Usage: python tools/build_perf.py [c3|soup1m|soup10m] [reps]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tracer_b200 import harness as H, layout as L  # noqa: E402
from tracer_b200._lib import check, lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prim = {"c3": lambda: H.scene_c3(2), "soup1m": lambda: H.scene_soup(1_000_000, 1, 0.01),
        "soup10m": lambda: H.scene_soup(10_000_000, 1, 0.004), "c2": H.scene_c2}[name]()
n = int((prim.bvhList["pType"] != L.BVH).sum())
leaves = prim.bvhList[1:n + 1].copy()
for gpu in (False, True, True, True)[: 1 + reps]:
    nodes = np.zeros(2 * n - 1, dtype=L.bvh_dtype); nodes[:n] = leaves
    nn, d = C.c_uint32(0), C.c_uint32(0)
    t = time.perf_counter()
    if gpu:
        check(lib.trq_bvh_build_tree_gpu(nodes.ctypes.data, n, 0, C.byref(nn), C.byref(d)), "gpu")
    else:
        check(lib.trq_bvh_build_tree(nodes.ctypes.data, n, C.byref(nn), C.byref(d)), "host")
    print(f"{name} n={n} {'gpu ' if gpu else 'host'} {time.perf_counter() - t:.4f} s depth {d.value}", flush=True)
