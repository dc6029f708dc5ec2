"""GPU: trq_bvh_build_tree_gpu (row f-1) must emit the SAME node array as the host restatement of BVH::buildTree
(byte for byte: splits, child order, numbering, boxes), and the trees it builds must trace identically."""
import ctypes as C
import time

import numpy as np
import pytest

from tracer_b200 import layout as L

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def both(leaves):
    from tracer_b200._lib import check, lib
    n = leaves.size
    out = []
    for gpu in (False, True):
        nodes = np.zeros(2 * n - 1, dtype=L.bvh_dtype)
        nodes[:n] = leaves
        nn, d = C.c_uint32(0), C.c_uint32(0)
        t = time.perf_counter()
        if gpu:
            check(lib.trq_bvh_build_tree_gpu(nodes.ctypes.data, n, 0, C.byref(nn), C.byref(d)), "gpu build")
        else:
            check(lib.trq_bvh_build_tree(nodes.ctypes.data, n, C.byref(nn), C.byref(d)), "host build")
        out.append((nodes, d.value, time.perf_counter() - t))
    return out


def leaves_of(prim):
    n = int((prim.bvhList["pType"] != L.BVH).sum())
    lv = prim.bvhList[1:n + 1].copy() if n > 1 else prim.bvhList[:1].copy()
    lv["parent"] = 0
    return lv


def assert_same_tree(a, b):
    for f in ("parent", "left", "right", "axis", "pType", "pIndex"):
        bad = np.nonzero(a[f] != b[f])[0]
        assert bad.size == 0, f"{f} differs at {bad[:8]} ({bad.size} nodes): {a[f][bad[:8]]} vs {b[f][bad[:8]]}"
    assert np.array_equal(a["mini"], b["mini"]) and np.array_equal(a["maxi"], b["maxi"])      # == also accepts -0 vs +0


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 33, 1000, 50_000])
def test_gpu_tree_equals_host_tree_soup(built, n):
    _torch()
    from tracer_b200 import harness as H
    (h, dh, _), (g, dg, _) = both(leaves_of(H.scene_soup(n, seed=n + 3, extent=0.05)))
    assert dh == dg
    assert_same_tree(h, g)


def test_gpu_tree_equals_host_tree_scenes(built):
    _torch()
    from tracer_b200 import harness as H
    for prim in (H.scene_reference_cornell(), H.scene_c2(), H.scene_c1()):
        (h, dh, _), (g, dg, _) = both(leaves_of(prim))
        assert dh == dg
        assert_same_tree(h, g)
        assert_same_tree(h, prim.bvhList)


def test_gpu_build_c3_full_size_and_timing(built, port):
    torch = _torch()
    from tracer_b200 import Scene, harness as H, hits_to_numpy, rays_to_torch
    from tracer_b200.scene import Primitive
    prim = H.scene_c3(2)
    (h, dh, th), (g, dg, tg) = both(leaves_of(prim))
    (_, _, _), (_, _, tg2) = both(leaves_of(prim))                      # second GPU build: context + allocator warm
    print(f"\nC3 {prim.nTri} triangles: host build {th:.2f} s, GPU build {tg:.3f} s (warm {tg2:.3f} s), depth {dg}")
    assert dh == dg
    assert_same_tree(h, g)
    # and it traces like the host-built tree
    gp = Primitive(triList=prim.triList, idxList=prim.idxList, bvhList=g)
    rays = H.random_rays(200000, seed=5, lo=(-245, 0, 0), hi=(800, 555, 555))
    got = hits_to_numpy(Scene(gp, 0).hit(rays_to_torch(rays, "cuda:0")))
    want = port.trace(prim, rays, nthreads=8)["hits"]
    for k in ("flags", "pType", "pIndex", "leafNode"):
        assert np.array_equal(got[k], want[k])
    assert tg2 < 10 * th          # sanity only; tools/build_perf.py records the real numbers (profiles/)


def test_gpu_build_identical_centroids_is_still_a_valid_tree(built, port):
    """>= 3 primitives with one centroid: the reference falls back to std::sort (unspecified order of equal keys), so
    only validity and query results are required here, not identity with the host tree."""
    _torch()
    from tracer_b200 import Scene, harness as H, hits_to_numpy, rays_to_torch
    from tracer_b200.scene import Primitive
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)
    ms = H.MeshSoup()
    for k in range(40):
        ms.add(pos * np.float32(1 + (k % 3)) - np.float32((k % 3) / 3.0) * pos.sum(0), [[0, 1, 2]])
    tri, idx = ms.arrays()
    prim = H.build_primitive(tri, idx)
    (h, _, _), (g, _, _) = both(leaves_of(prim))
    gp = Primitive(triList=tri, idxList=idx, bvhList=g)
    rays = H.random_rays(20000, seed=3, lo=(-1, -1, -1), hi=(2, 2, 1))
    got = hits_to_numpy(Scene(gp, 0).hit(rays_to_torch(rays, "cuda:0")))
    want = port.trace(gp, rays)["hits"]                      # oracle on the GPU-built tree
    ref = port.trace(prim, rays)["hits"]                     # and the host-built tree: same closest t
    for k in ("flags", "pType", "pIndex", "leafNode"):
        assert np.array_equal(got[k], want[k])
    assert np.array_equal(got["t"], ref["t"]) and np.array_equal(got["flags"] & 1, ref["flags"] & 1)


@pytest.mark.parametrize("name", ["soup2", "soup3", "soup7", "soup100", "soup2000", "cornell_mixed"])
def test_gpu_tree_equals_reference_builder_golden(built, name):
    """The GPU builder against trees made by the REFERENCE's own BVH::buildTree (tests/golden/bvh_golden.npz, generated by
    tests/golden/make_bvh_golden.py from oracle/_ref/libtracer_ref_builder.so)."""
    import os
    _torch()
    from tracer_b200._lib import check, lib
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bvh_golden.npz"))
    leaves, want = z[f"{name}/leaves"], z[f"{name}/tree"]
    n = leaves.size
    nodes = np.zeros(2 * n - 1, dtype=L.bvh_dtype)
    nodes[:n] = leaves
    nn, d = C.c_uint32(0), C.c_uint32(0)
    check(lib.trq_bvh_build_tree_gpu(nodes.ctypes.data, n, 0, C.byref(nn), C.byref(d)), "gpu build")
    assert_same_tree(nodes, want)
