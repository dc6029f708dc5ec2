"""GPU: randomized parity sweep aimed at the places where CUDA and host arithmetic could part ways: rays with zero
direction components (1/0 = inf slabs), origins lying exactly on box planes / vertices / edges (0 * inf = NaN inside
AABB::hit_t, where only fminf/fmaxf's NaN rule decides), degenerate and duplicated triangles, tiny and huge tmax,
mixed leaf types, any-hit and closest-hit, both kernels."""
import os

import numpy as np
import pytest

from tracer_b200 import layout as L

from .test_gpu_parity import assert_hits_equal, gpu_trace

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def adversarial_rays(rng, prim, n):
    from tracer_b200._lib import lib
    bvh = prim.bvhList
    rays = np.zeros(n, dtype=L.ray_dtype)
    lo, hi = bvh["mini"][0], bvh["maxi"][0]
    rays["o"] = rng.uniform(lo - 0.2 * (hi - lo) - 1e-3, hi + 0.2 * (hi - lo) + 1e-3, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    k = np.arange(n)
    d[k % 4 == 1, rng.integers(3)] = 0.0                                   # one zero component
    two = k % 16 == 2
    d[two] = 0.0; d[two, rng.integers(3)] = rng.choice([-1.0, 1.0])        # axis-aligned
    # origins exactly on a plane of some node's box, on a vertex, or on an edge midpoint
    node = bvh[rng.integers(0, bvh.size, n)]
    plane = k % 3 == 0
    ax = rng.integers(0, 3, n)
    on = np.where(rng.random(n) < 0.5, node["mini"][k, ax], node["maxi"][k, ax])
    rays["o"][plane, ax[plane]] = on[plane]
    if prim.triList.size:
        v = prim.triList["v"]
        tri = prim.idxList.reshape(-1, 3)[rng.integers(0, prim.nTri, n)]
        vert = k % 7 == 3
        rays["o"][vert] = v[tri[vert, 0]]
        edge = k % 7 == 5
        target = (v[tri[edge, 0]] + v[tri[edge, 1]]) * np.float32(0.5)     # aim exactly at an edge midpoint
        d[edge] = target - rays["o"][edge]
    rays["d"] = d
    lib.trqh_normalize_rays(rays.ctypes.data, n)
    bad = ~np.isfinite(rays["d"]).all(axis=1)                              # 0/0 from a zero-length direction
    rays["d"][bad] = np.array([0, 0, 1], dtype=np.float32)
    tm = np.full(n, L.FLT_MAX, dtype=np.float32)
    tm[k % 5 == 0] = rng.uniform(1e-3, 3.0, (k % 5 == 0).sum()).astype(np.float32) * np.float32(np.linalg.norm(hi - lo))
    tm[k % 11 == 0] = np.float32(1e-30)
    rays["tmax"] = tm
    return rays


def scenes(rng):
    from tracer_b200 import harness as H
    yield "cornell mixed", H.scene_reference_cornell()
    yield "soup", H.scene_soup(3000, seed=int(rng.integers(1 << 30)), extent=0.15)
    # degenerate + duplicated + axis-aligned triangles on an integer lattice (many exact ties, flat boxes)
    pts = rng.integers(-3, 4, (600, 3, 3)).astype(np.float32)
    pts[::9, 2] = pts[::9, 1]                                              # zero-area triangles
    pts[1::13] = pts[0::13][: len(pts[1::13])]                             # exact duplicates
    ms = H.MeshSoup()
    for t in pts:
        ms.add(t, [[0, 1, 2]])
    tri, idx = ms.arrays()
    yield "lattice", H.build_primitive(tri, idx)
    sph = np.array([H.make_sphere(float(rng.uniform(0.2, 1.5)), rng.uniform(-3, 3, 3), i) for i in range(60)], dtype=L.sphere_dtype)
    yield "spheres", H.build_primitive(spheres=sph)


def _seeds():
    """Three seeds in the regular suite; TRQ_FUZZ_SEEDS=lo-hi widens the campaign (profiles/r01_fuzz_campaign.txt)."""
    extra = os.environ.get("TRQ_FUZZ_SEEDS", "")
    if "-" in extra:
        lo, hi = (int(x) for x in extra.split("-"))
        return [1, 2, 3] + list(range(lo, hi + 1))
    return [1, 2, 3]


@pytest.mark.parametrize("seed", _seeds())
def test_adversarial_parity(built, port, seed):
    _torch()
    from tracer_b200 import Scene
    rng = np.random.default_rng(seed)
    for name, prim in scenes(rng):
        scene = Scene(prim, 0)
        rays = adversarial_rays(rng, prim, 40000)
        for any_hit in (False, True):
            want = port.trace(prim, rays, any=any_hit, counters=True, nthreads=8)
            for reflayout in (False, True):
                got = gpu_trace(scene, rays, any_hit, reflayout)
                assert_hits_equal(got, want["hits"], f"{name} seed={seed} any={any_hit} reflayout={reflayout}")
        scene.close()
