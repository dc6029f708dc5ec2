"""CPU: the C restatement (oracle/oracle_rq.c) against the golden vectors produced by the reference's own
headers compiled verbatim (tests/golden/make_golden.py), and -- where oracle/_ref is present -- against that
build directly. This is what pins the oracle before the CUDA path is compared with it."""
import numpy as np
import pytest

from tracer_b200 import layout as L

from .util import CASES, assert_records_match_reference, bits, golden, golden_case


@pytest.mark.parametrize("name", CASES)
def test_port_matches_golden(port, name):
    prim, rays, any_hit, want = golden_case(name)
    got = port.trace(prim, rays, any=any_hit, records=True, counters=True)
    assert_records_match_reference(got["records"], want, where=name)
    h = got["hits"]
    hit = (h["flags"] & 1) == 1
    assert np.array_equal(hit, want["hit"] == 1)
    # the id the restatement adds must be consistent with the record the reference produced
    assert np.array_equal(bits(h["t"][hit]), bits(want["t"][hit]))
    assert np.array_equal(h["material"][hit], want["material"][hit])
    assert np.array_equal(((h["flags"] >> 1) & 1)[hit], want["front"][hit])
    leaf = prim.bvhList[h["leafNode"][hit]]
    assert np.array_equal(leaf["pType"].astype(np.uint32), h["pType"][hit])
    assert np.array_equal(leaf["pIndex"], h["pIndex"][hit])
    assert (leaf["pType"] != L.BVH).all()


def test_golden_covers_order_dependent_cases(port):
    """The fixture set must exercise what makes ids order-dependent: exact-t ties (Triangle.hh:71 accepts
    t == range.y) and every leaf type of the dispatch switch (Render.hh:213-242)."""
    ties = 0
    types = set()
    for name in CASES:
        prim, rays, any_hit, _ = golden_case(name)
        got = port.trace(prim, rays, any=any_hit, counters=True)
        ties += got["totals"]["n_tie"]
        h = got["hits"]
        types |= set(np.unique(h["pType"][(h["flags"] & 1) == 1]).tolist())
    assert ties > 0, "no exact-t tie in the golden rays"
    assert {L.SPHERE, L.SQUARE, L.CUBE, L.TRIANGLE} <= types


def test_leaf_known_answers(port):
    z = golden()
    for i in range(len(z["leaf/aabb_hit"])):
        h, t = port.aabb_hit_t(z["leaf/aabb_box"][i], z["leaf/aabb_o"][i], z["leaf/aabb_d"][i], z["leaf/aabb_range"][i])
        assert int(h) == int(z["leaf/aabb_hit"][i])
        if h:
            assert bits(np.float32(t)) == bits(z["leaf/aabb_t"][i])
        assert h == port.aabb_hit(z["leaf/aabb_box"][i], z["leaf/aabb_o"][i], z["leaf/aabb_d"][i], z["leaf/aabb_range"][i])
    for i in range(len(z["leaf/off_p"])):
        assert np.array_equal(bits(port.offset_ray(z["leaf/off_p"][i], z["leaf/off_n"][i])), bits(z["leaf/off_out"][i]))


@pytest.mark.parametrize("name", CASES)
def test_port_matches_verbatim_reference_live(port, reference, name):
    """Same inputs through the verbatim build, live (only where oracle/_ref exists): also checks the fixture is fresh."""
    prim, rays, any_hit, want = golden_case(name)
    live = reference.trace(prim, rays, any=any_hit)
    m = want["hit"] == 1
    assert np.array_equal(live["hit"], want["hit"])
    for k in ("t", "p", "gn", "sn", "uv", "front", "material"):
        assert np.array_equal(bits(live[k][m]), bits(want[k][m])), k


def test_port_vs_reference_random_scenes(port, reference):
    """Wider live cross-check: bigger scenes than the fixtures can hold."""
    from tracer_b200 import harness as H
    prim = H.scene_reference_cornell()
    rays = H.cornell_camera_rays(160, 90)
    first = reference.trace(prim, rays, nthreads=4)
    for r, any_hit in ((rays, False), (H.bounce_rays(first)[0], False),
                       (H.shadow_rays(first, prim.squareList[5:6], prim.squareList[6:7])[0], True),
                       (H.random_rays(20000, seed=5, lo=(-245, 0, 0), hi=(800, 555, 555)), False)):
        got = port.trace(prim, r, any=any_hit, records=True, nthreads=4)["records"]
        assert_records_match_reference(got, reference.trace(prim, r, any=any_hit, nthreads=4))
    soup = H.scene_soup(20000, seed=8, extent=0.05)
    r = H.random_rays(20000, seed=6)
    assert_records_match_reference(port.trace(soup, r, records=True, nthreads=4)["records"], reference.trace(soup, r, nthreads=4))
