"""CPU: the sequential restatement of BVH::buildNode / make / buildTree (BVH.hh:35-314), pinned three ways:
  * against the reference's OWN builder, compiled from its source (oracle/ref_builder.cpp -> oracle/_ref/): node arrays
    byte for byte, live where /root/reference is mounted and through tests/golden/bvh_golden.npz everywhere;
  * by the structural contract the ray query relies on;
  * by brute force: the closest hit found through the tree equals the closest hit over all primitives."""
import os
import numpy as np
import pytest

from tracer_b200 import harness as H, layout as L
from tracer_b200._lib import ERR_DEPTH, ERR_INVALID, TrqError, lib


def check_tree(bvh, n_leaves):
    assert bvh.size == 2 * n_leaves - 1
    if n_leaves == 1:
        assert bvh[0]["pType"] != L.BVH and bvh[0]["parent"] == 0
        return 0
    assert bvh[0]["pType"] == L.BVH and bvh[0]["parent"] == 0                       # root at 0 (BVH.hh:263-268)
    leaves = bvh[1:n_leaves + 1]
    assert (leaves["pType"] != L.BVH).all() and (leaves["left"] == 0).all() and (leaves["right"] == 0).all()
    inter = np.concatenate([[0], np.arange(n_leaves + 1, bvh.size)])
    assert (bvh[inter]["pType"] == L.BVH).all()
    l, r = bvh[inter]["left"], bvh[inter]["right"]
    assert np.array_equal(bvh[l]["parent"], inter) and np.array_equal(bvh[r]["parent"], inter)
    kids = np.concatenate([l, r])
    assert np.array_equal(np.sort(kids), np.arange(1, bvh.size))                    # every non-root node is a child exactly once
    # parent box = union of child boxes, exactly (fminf / fmaxf)
    assert np.array_equal(bvh[inter]["mini"], np.minimum(bvh[l]["mini"], bvh[r]["mini"]))
    assert np.array_equal(bvh[inter]["maxi"], np.maximum(bvh[l]["maxi"], bvh[r]["maxi"]))
    assert (bvh[inter]["axis"] <= 2).all()
    depth = np.zeros(bvh.size, dtype=np.int64)
    for i in inter[1:][::-1]:                                                        # parents have larger indices (post-order)
        depth[i] = depth[bvh[i]["parent"]] + 1
    return int(depth.max())


@pytest.mark.parametrize("n", [1, 2, 3, 7, 100, 5000])
def test_structure_soup(built, n):
    prim = H.scene_soup(n, seed=n, extent=0.1)
    d = check_tree(prim.bvhList, n)
    assert d <= 31
    tri_leaf = prim.bvhList[prim.bvhList["pType"] == L.TRIANGLE]
    assert np.array_equal(np.sort(tri_leaf["pIndex"]), np.arange(n))
    v = prim.triList["v"][prim.idxList.reshape(-1, 3)]                               # leaf box = min/max of the 3 vertices
    order = np.argsort(tri_leaf["pIndex"])
    assert np.array_equal(tri_leaf["mini"][order], v.min(1)) and np.array_equal(tri_leaf["maxi"][order], v.max(1))


def test_structure_mixed_and_meshes(built):
    prim = H.scene_reference_cornell()
    n = (prim.bvhList["pType"] != L.BVH).sum()
    check_tree(prim.bvhList, n)
    # leaf creation order of AAPLRenderer.mm:454-468,546-591: spheres, cubes, squares, triangles
    types = prim.bvhList[1:n + 1]["pType"]
    assert list(types[:12]) == [L.SPHERE] * 12 and list(types[12:14]) == [L.CUBE] * 2 and list(types[14:21]) == [L.SQUARE] * 7
    assert (types[21:] == L.TRIANGLE).all()
    c2 = H.scene_c2()
    assert check_tree(c2.bvhList, c2.nTri) <= 31


def test_identical_centroids_fall_back_to_median_split(built):
    """All centroids equal: relative() is 0/0, every SAH cost is NaN, the partition is empty and the sort+median
    fallback (BVH.hh:187-195) must still produce a valid tree."""
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)
    ms = H.MeshSoup()
    for _ in range(33):
        ms.add(pos, [[0, 1, 2]])
    tri, idx = ms.arrays()
    prim = H.build_primitive(tri, idx)
    assert check_tree(prim.bvhList, 33) <= 31


def test_deterministic(built):
    a = H.scene_soup(3000, seed=5, extent=0.05).bvhList
    b = H.scene_soup(3000, seed=5, extent=0.05).bvhList
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_errors(built):
    import ctypes as C
    n, d = C.c_uint32(0), C.c_uint32(0)
    assert lib.trq_bvh_build_tree(None, 4, C.byref(n), C.byref(d)) == ERR_INVALID
    node = np.zeros(1, dtype=L.bvh_dtype)
    assert lib.trq_bvh_build_tree(node.ctypes.data, 0, C.byref(n), C.byref(d)) == ERR_INVALID
    assert b"empty" in lib.trq_last_error_string()


def test_tree_finds_the_brute_force_closest_hit(built, port):
    """Independent check of builder + traversal: for every ray, t through the BVH == min t over ALL triangles
    tested one by one with the same intersector (ids may differ only on exact ties)."""
    prim = H.scene_soup(400, seed=21, extent=0.3)
    rays = H.random_rays(1500, seed=22)
    got = port.trace(prim, rays)["hits"]
    idx = prim.idxList.reshape(-1, 3)
    for i in range(0, rays.size, 3):
        best = np.float32(L.FLT_MAX)
        for k in range(len(idx)):
            r = np.array([L.FLT_MIN, best], dtype=np.float32)
            h, r2, _, _ = port.triangle_hit(prim.triList, idx[k], rays["o"][i], rays["d"][i], r)
            if h:
                best = r2[1]
        if best < L.FLT_MAX:
            assert (got["flags"][i] & 1) and got["t"][i] == best
        else:
            assert not (got["flags"][i] & 1)


# ------------------------------------------------------------------------------------------------
# parity with the reference's own builder
BVH_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bvh_golden.npz")
BVH_CASES = ("soup2", "soup3", "soup7", "soup100", "soup2000", "cornell_mixed", "lattice", "same_centroid")


def build_with_product(leaves):
    import ctypes as C
    n = leaves.size
    nodes = np.zeros(2 * n - 1, dtype=L.bvh_dtype)
    nodes[:n] = leaves
    nn, md = C.c_uint32(0), C.c_uint32(0)
    assert lib.trq_bvh_build_tree(nodes.ctypes.data, n, C.byref(nn), C.byref(md)) == 0, lib.trq_last_error_string()
    assert nn.value == nodes.size
    return nodes


def assert_same_tree(got, want, where):
    for f in got.dtype.names:
        if f == "pad":
            continue                                             # uninitialised in the reference (no such member in struct BVH)
        a, b = got[f], want[f]
        same = np.array_equal(a.view(np.uint32), b.view(np.uint32)) if a.dtype.kind == "f" else np.array_equal(a, b)
        assert same, f"{where}: field {f} differs at nodes {np.nonzero((a != b).reshape(got.size, -1).any(axis=1))[0][:5]}"


@pytest.mark.parametrize("name", BVH_CASES)
def test_tree_equals_reference_builder_golden(built, name):
    z = np.load(BVH_GOLDEN)
    assert_same_tree(build_with_product(z[f"{name}/leaves"]), z[f"{name}/tree"], name)


def test_build_node_equals_reference_golden(built):
    """BVH::buildNode (BVH.hh:273-314): world box of the 8 transformed corners."""
    z = np.load(BVH_GOLDEN)
    want = z["node/out"]
    for i in range(want.size):
        got = np.zeros(1, dtype=L.bvh_dtype)
        assert lib.trq_bvh_build_node(z["node/lo"][i].ctypes.data, z["node/hi"][i].ctypes.data, z["node/model"][i].ctypes.data,
                                      int(L.CUBE), i, got.ctypes.data) == 0
        assert_same_tree(got, want[i:i + 1], f"node {i}")


def test_tree_equals_reference_builder_live(built, reference_builder):
    """Larger and random inputs against the reference builder itself (only where oracle/_ref could be built)."""
    rng = np.random.default_rng(123)
    prims = [H.scene_soup(n, seed=100 + n, extent=e) for n, e in ((20000, 0.02), (5000, 0.3), (333, 1.0))]
    prims += [H.scene_reference_cornell(), H.scene_c2(), H.scene_c1()]
    for k, prim in enumerate(prims):
        n = int((prim.bvhList["pType"] != L.BVH).sum())
        leaves = prim.bvhList[1:n + 1].copy()
        leaves["parent"] = 0
        assert_same_tree(build_with_product(leaves), reference_builder.build_tree(leaves), f"scene {k}")
        # the harness built prim.bvhList with the product builder: it must be that same tree
        assert_same_tree(prim.bvhList, reference_builder.build_tree(leaves), f"scene {k} (harness)")
    for _ in range(40):                                          # tiny trees, heavy ties: integer lattice boxes
        n = int(rng.integers(2, 40))
        lv = np.zeros(n, dtype=L.bvh_dtype)
        lo = rng.integers(-2, 3, (n, 3)).astype(np.float32)
        lv["mini"], lv["maxi"] = lo, lo + rng.integers(0, 3, (n, 3)).astype(np.float32)
        lv["pType"], lv["pIndex"] = L.TRIANGLE, np.arange(n)
        assert_same_tree(build_with_product(lv), reference_builder.build_tree(lv), f"lattice boxes n={n}")


def test_c3_tree_equals_reference_builder(built, reference_builder):
    """BASELINE C3 at full size: the 1.0 M-triangle scene's tree (built by the product builder inside the harness) is the
    tree the reference's own BVH::buildTree makes of the same leaves."""
    prim = H.scene_c3(2)
    n = prim.nTri
    assert n > 1_000_000 and prim.bvhList.size == 2 * n - 1
    leaves = prim.bvhList[1:n + 1].copy()
    leaves["parent"] = 0
    assert_same_tree(prim.bvhList, reference_builder.build_tree(leaves), "C3")
