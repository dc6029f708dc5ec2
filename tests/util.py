"""Shared helpers for the test suite."""
import os

import numpy as np

from tracer_b200 import layout as L
from tracer_b200.scene import Primitive

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rq_golden.npz")
FIELDS = ("sphereList", "squareList", "cubeList", "triList", "idxList", "bvhList")
CASES = ("cornell_primary", "cornell_bounce", "cornell_shadow_any", "soup_closest", "soup_any_tmax",
         "icosphere_ties", "c1_spheres", "one_leaf", "two_leaves")

_cache = {}


def golden():
    if "z" not in _cache:
        _cache["z"] = np.load(GOLDEN)
    return _cache["z"]


def golden_case(name):
    """-> (Primitive, rays, any, reference records)"""
    z = golden()
    prim = Primitive(**{k: z[f"{name}/{k}"] for k in FIELDS})
    return prim, z[f"{name}/rays"], bool(z[f"{name}/any"][0]), z[f"{name}/records"]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def assert_records_match_reference(recs, want, hits=None, sphere_uv_atol=None, where=""):
    """Bit-exact comparison of HitRecord dumps on the rays the reference reports as hit."""
    assert np.array_equal(recs["hit"], want["hit"]), f"{where}: hit flags differ"
    m = want["hit"] == 1
    for k in ("t", "p", "gn", "sn", "front", "material"):
        assert np.array_equal(bits(recs[k][m]), bits(want[k][m])), f"{where}: {k} differs"
    if sphere_uv_atol is None or hits is None:
        assert np.array_equal(bits(recs["uv"][m]), bits(want["uv"][m])), f"{where}: uv differs"
    else:
        sph = hits["pType"] == L.SPHERE
        assert np.array_equal(bits(recs["uv"][m & ~sph]), bits(want["uv"][m & ~sph])), f"{where}: uv differs"
        assert np.allclose(recs["uv"][m & sph], want["uv"][m & sph], rtol=0, atol=sphere_uv_atol), f"{where}: sphere uv"
