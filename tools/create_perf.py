#!/usr/bin/env python
"""Developer tool: wall time of scene creation from host arrays and from device arrays, and of the device-resident build.
Usage: python tools/create_perf.py [c3|soup1m|c2]     (TRQ_LIB=tools/variants/libtracer_rq_timing.so prints the stages)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tracer_b200 import BVHBuilder, DevicePrimitive, Scene, harness as H, layout as L  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
prim = {"c3": lambda: H.scene_c3(2), "soup1m": lambda: H.scene_soup(1_000_000, 1, 0.01), "c2": H.scene_c2}[name]()
n = int((prim.bvhList["pType"] != L.BVH).sum())
tri = torch.from_numpy(prim.triList.view(np.uint8).reshape(-1).copy()).cuda()
idx = torch.from_numpy(prim.idxList.view(np.uint8).reshape(-1).copy()).cuda()
b = BVHBuilder(); b._chunks.append(prim.bvhList[1:n + 1].copy())
leaves_dev = torch.from_numpy(b.leaves().view(np.uint8).reshape(-1).copy()).cuda()
from tracer_b200._lib import check, lib  # noqa: E402
import ctypes as C  # noqa: E402
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = Scene(prim, 0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    s.close()
    nodes = torch.empty((2 * n - 1) * 64, dtype=torch.uint8, device="cuda")
    nodes[: n * 64] = leaves_dev
    torch.cuda.synchronize(); t2 = time.perf_counter()
    nn, dd = C.c_uint32(0), C.c_uint32(0)
    check(lib.trq_bvh_build_tree_device(nodes.data_ptr(), n, 0, C.byref(nn), C.byref(dd)), "build")
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print("-- device create", file=sys.stderr, flush=True)
    s = Scene(DevicePrimitive(triList=tri, idxList=idx, bvhList=nodes), 0)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    s.close()
    print(f"{name} ({n} leaves) rep {rep}: create from host arrays {1e3 * (t1 - t0):.2f} ms | device build {1e3 * (t3 - t2):.2f} ms + "
          f"device create {1e3 * (t4 - t3):.2f} ms = {1e3 * (t4 - t2):.2f} ms", flush=True)
