#!/usr/bin/env python
"""Generates tests/golden/bvh_golden.npz: leaf arrays and the trees the REFERENCE's own builder (BVH::buildTree,
RT_Metal/Metal/BVH.hh:35-269, compiled by oracle/Makefile into oracle/_ref/libtracer_ref_builder.so) makes of them,
plus BVH::buildNode outputs for random boxes and model matrices. Run in the container that mounts /root/reference:

    python tests/golden/make_bvh_golden.py

The two padding words of a node are uninitialised in the reference and are stored as zero."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.pyoracle import ReferenceBuilder  # noqa: E402
from tracer_b200 import harness as H, layout as L  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bvh_golden.npz")


def leaves_of(prim):
    n = int((prim.bvhList["pType"] != L.BVH).sum())
    lv = prim.bvhList[1:n + 1].copy() if n > 1 else prim.bvhList[:1].copy()
    lv["parent"] = 0
    return lv


def cases():
    rng = np.random.default_rng(77)
    out = {}
    for n in (1, 2, 3, 7, 100, 2000):
        out[f"soup{n}"] = leaves_of(H.scene_soup(n, seed=n, extent=0.1))
    # the reference's own leaf mix: spheres, cubes (model matrices baked by buildNode), squares, a small mesh
    ms = H.MeshSoup()
    p, t = H.load_mesh("teapot")
    ms.add(H.place_mesh(p), t[:400])
    tri, idx = ms.arrays()
    out["cornell_mixed"] = leaves_of(H.build_primitive(tri, idx, spheres=H.cornell_spheres(), squares=H.cornell_squares(), cubes=H.cornell_cubes()))
    # integer lattice: many equal centroids and exact cost ties, zero-area and duplicated triangles
    pts = rng.integers(-3, 4, (500, 3, 3)).astype(np.float32)
    pts[::9, 2] = pts[::9, 1]
    pts[1::13] = pts[0::13][: len(pts[1::13])]
    ms = H.MeshSoup()
    for t3 in pts:
        ms.add(t3, [[0, 1, 2]])
    tri, idx = ms.arrays()
    out["lattice"] = leaves_of(H.build_primitive(tri, idx))
    # all centroids identical: SAH costs are NaN, the sort + median fallback decides (BVH.hh:187-195)
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)
    ms = H.MeshSoup()
    for _ in range(33):
        ms.add(pos, [[0, 1, 2]])
    tri, idx = ms.arrays()
    out["same_centroid"] = leaves_of(H.build_primitive(tri, idx))
    return out


def main():
    ref = ReferenceBuilder()
    blob = {}
    for name, lv in cases().items():
        tree = ref.build_tree(lv) if lv.size > 1 else None
        blob[f"{name}/leaves"] = lv
        if tree is not None:
            blob[f"{name}/tree"] = tree
        print(name, lv.size, "leaves")
    rng = np.random.default_rng(3)
    n = 64
    lo = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
    hi = (lo + rng.uniform(0, 4, (n, 3))).astype(np.float32)
    model = rng.normal(size=(n, 16)).astype(np.float32)
    model[::4] = np.eye(4, dtype=np.float32).reshape(16)
    nodes = np.zeros(n, dtype=L.bvh_dtype)
    for i in range(n):
        nodes[i] = ref.build_node(lo[i], hi[i], model[i], int(L.CUBE), i, L.bvh_dtype)
    blob.update({"node/lo": lo, "node/hi": hi, "node/model": model, "node/out": nodes})
    np.savez_compressed(OUT, **blob)
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
