#!/usr/bin/env python
"""Generate the golden vectors of the ray query by EXECUTING THE REFERENCE ITSELF.

The reference (iaomw/Tracer) ships no tests and no golden vectors, so the pin is:
its own RT_Metal/Metal/Render.hh compiled verbatim as host C++ (oracle/ref_scene_hit.cpp ->
oracle/_ref/libtracer_ref.so) run on the inputs below. Each case stores INPUTS (the six reference-layout
scene arrays + the rays) and the reference's OUTPUTS (Scene::hit return value and the HitRecord fields)
so that the C restatement (oracle/oracle_rq.c) and the CUDA kernels can be checked on any box, without
/root/reference. Run here, once:   python tests/golden/make_golden.py
Output: tests/golden/rq_golden.npz  (+ leaf-level vectors for the five intersectors and offset_ray).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.pyoracle import Reference  # noqa: E402
from tracer_b200 import harness as H, layout as L  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rq_golden.npz")
FIELDS = ("sphereList", "squareList", "cubeList", "triList", "idxList", "bvhList")


def icosphere(levels, radius, center):
    t = (1.0 + 5 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v = (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
    v, f = H.subdivide(v, f, levels)
    v = v / np.linalg.norm(v, axis=1, keepdims=True)
    return (v * np.float32(radius) + np.asarray(center, dtype=np.float32)).astype(np.float32), f


def cases():
    """name -> (Primitive, rays, any)"""
    out = {}
    # (a) the reference's own leaf mix: Cube + Square + Sphere leaves, no mesh
    prim = H.build_primitive(spheres=H.cornell_spheres(), squares=H.cornell_squares(), cubes=H.cornell_cubes())
    ref = Reference()
    primary = H.cornell_camera_rays(64, 36)
    recs = ref.trace(prim, primary)
    out["cornell_primary"] = (prim, primary[:2048], False)
    out["cornell_bounce"] = (prim, H.bounce_rays(recs)[0][:2048], False)
    out["cornell_shadow_any"] = (prim, H.shadow_rays(recs, prim.squareList[5:6], prim.squareList[6:7])[0][:2048], True)
    # (b) incoherent soup, closest and any-hit with finite tmax
    soup = H.scene_soup(1500, seed=3, extent=0.12)
    rnd = H.random_rays(2048, seed=4)
    out["soup_closest"] = (soup, rnd, False)
    short = rnd.copy(); short["tmax"] = 0.4
    out["soup_any_tmax"] = (soup, short, True)
    # (c) shared-vertex mesh: rays aimed exactly at vertices and edge midpoints -> exact-t ties (order-dependent ids)
    pos, tris = icosphere(3, 1.0, (0, 0, 0))
    ms = H.MeshSoup(); ms.add(pos, tris)
    tri, idx = ms.arrays()
    ico = H.build_primitive(tri, idx)
    targets = np.concatenate([pos[:700], (pos[tris[:700, 0]] + pos[tris[:700, 1]]) * np.float32(0.5)])
    rays = np.zeros(len(targets) + 600, dtype=L.ray_dtype)
    origin = np.array([0.3, 0.2, 4.0], dtype=np.float32)
    rays["o"][: len(targets)] = origin
    rays["d"][: len(targets)] = targets - origin
    inside = H.random_rays(600, seed=9, lo=(-0.2, -0.2, -0.2), hi=(0.2, 0.2, 0.2))
    rays[len(targets):] = inside
    rays["tmax"] = L.FLT_MAX
    from tracer_b200._lib import lib
    lib.trqh_normalize_rays(rays.ctypes.data, rays.size)
    out["icosphere_ties"] = (ico, rays, False)
    # (d) C1 random spheres (RT_Nextweek randomScene as Sphere leaves)
    c1 = H.scene_c1()
    cam = H.camera_rays((13, 2, 3), (0, 0, 0), np.float32(20 * np.pi / 180), 64, 32)
    out["c1_spheres"] = (c1, cam, False)
    # (e) degenerate trees: a single leaf (root is a leaf) and two leaves
    quad = H.make_vertices(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float32), [[0, 1, 2], [0, 1, 3]])
    r = H.random_rays(512, seed=12, lo=(-0.5, -0.5, -0.5), hi=(1, 1, 1))
    out["one_leaf"] = (H.build_primitive(quad, np.array([0, 1, 2], dtype=np.uint32)), r, False)
    out["two_leaves"] = (H.build_primitive(quad, np.array([0, 1, 2, 0, 1, 3], dtype=np.uint32)), r, False)
    return out


def leaf_vectors(ref, rng):
    """Known-answer vectors for the individual intersectors, straight from the verbatim headers."""
    v = {}
    n = 256
    box = np.zeros((n, 8), dtype=np.float32)
    lo = rng.uniform(-1, 1, (n, 3)).astype(np.float32); ext = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    box[:, 0:3], box[:, 4:7] = lo, lo + ext
    o = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[::7, 0] = 0.0; d[::11, 1] = 0.0                       # axis-parallel
    o[::5] = lo[::5] + ext[::5] * np.float32(0.5)           # origin inside
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rngs = np.stack([np.full(n, L.FLT_MIN, dtype=np.float32), rng.uniform(0.1, 6, n).astype(np.float32)], 1)
    res = [ref.aabb_hit_t(box[i], o[i], d[i], rngs[i]) for i in range(n)]
    v.update(aabb_box=box, aabb_o=o, aabb_d=d, aabb_range=rngs,
             aabb_hit=np.array([r[0] for r in res], dtype=np.uint8), aabb_t=np.array([r[1] for r in res], dtype=np.float32))
    p = rng.uniform(-600, 600, (n, 3)).astype(np.float32); p[::4] *= np.float32(1e-4)
    nn = rng.normal(size=(n, 3)).astype(np.float32); nn = (nn / np.linalg.norm(nn, axis=1, keepdims=True)).astype(np.float32)
    v.update(off_p=p, off_n=nn, off_out=np.stack([ref.offset_ray(p[i], nn[i]) for i in range(n)]))
    return v


def main():
    ref = Reference()
    blob = {}
    for name, (prim, rays, any_hit) in cases().items():
        recs = ref.trace(prim, rays, any=any_hit)
        for k in FIELDS:
            blob[f"{name}/{k}"] = getattr(prim, k)
        blob[f"{name}/rays"] = rays
        blob[f"{name}/any"] = np.array([1 if any_hit else 0], dtype=np.uint8)
        blob[f"{name}/records"] = recs
        print(f"{name}: {prim.bvhList.size} nodes, {rays.size} rays, hit fraction {recs['hit'].mean():.3f}")
    for k, a in leaf_vectors(ref, np.random.default_rng(2026)).items():
        blob[f"leaf/{k}"] = a
    np.savez_compressed(OUT, **blob)
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
