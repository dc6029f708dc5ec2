"""CPU: leaf-level functions of the C restatement against the verbatim reference headers, on random and
hand-picked inputs (origin inside the box, axis-parallel directions, t == range.y ties, degenerate det)."""
import numpy as np
import pytest

from tracer_b200 import harness as H, layout as L

from .util import bits


def _rec_equal(a, b):
    for k in ("hit", "t", "p", "gn", "sn", "uv", "front", "material"):
        if not np.array_equal(bits(a[k]), bits(b[k])):
            return False
    return True


def test_aabb_random_and_edges(port, reference):
    rng = np.random.default_rng(1)
    box = np.zeros(8, dtype=np.float32)
    for i in range(3000):
        lo = rng.uniform(-2, 2, 3).astype(np.float32)
        box[0:3], box[4:7] = lo, lo + rng.uniform(0, 2, 3).astype(np.float32)
        o = rng.uniform(-4, 4, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        if i % 5 == 0:
            d[rng.integers(3)] = 0.0                      # 1/0 = inf
        if i % 7 == 0:
            o = (box[0:3] + box[4:7]) / 2                 # inside
        if i % 11 == 0:
            o[0] = box[0]                                 # on a slab plane (0 * inf = NaN when d.x == 0)
        d = (d / np.linalg.norm(d)).astype(np.float32)
        r = np.array([L.FLT_MIN, rng.uniform(0.01, 10)], dtype=np.float32)
        a, b = port.aabb_hit_t(box, o, d, r), reference.aabb_hit_t(box, o, d, r)
        assert a[0] == b[0] and (not a[0] or bits(a[1]) == bits(b[1]))
        assert port.aabb_hit(box, o, d, r) == reference.aabb_hit(box, o, d, r)


def test_triangle_random_ties_and_degenerate(port, reference):
    rng = np.random.default_rng(2)
    for i in range(2000):
        pos = rng.uniform(-1, 1, (3, 3)).astype(np.float32)
        if i % 50 == 0:
            pos[2] = pos[0] + (pos[1] - pos[0]) * np.float32(0.5)      # zero-area triangle: det ~ 0
        tv = H.make_vertices(pos, [[0, 1, 2]])
        tv["n"] = rng.normal(size=(3, 3)).astype(np.float32)
        tv["uv"] = rng.uniform(0, 1, (3, 2)).astype(np.float32)
        o = rng.uniform(-3, 3, 3).astype(np.float32)
        target = pos.mean(0) if i % 3 else pos[i % 3]                   # through the centroid or exactly at a vertex
        d = (target - o); d = (d / np.linalg.norm(d)).astype(np.float32)
        r = np.array([L.FLT_MIN, L.FLT_MAX], dtype=np.float32)
        ha, ra, reca, _ = port.triangle_hit(tv, [0, 1, 2], o, d, r)
        hb, rb, recb = reference.triangle_hit(tv, [0, 1, 2], o, d, r)
        assert ha == hb and np.array_equal(bits(ra), bits(rb)) and (not ha or _rec_equal(reca, recb))
        if ha:                                                          # t == range.y must still be accepted (Triangle.hh:71)
            tie = np.array([L.FLT_MIN, reca["t"]], dtype=np.float32)
            assert port.triangle_hit(tv, [0, 1, 2], o, d, tie)[0] and reference.triangle_hit(tv, [0, 1, 2], o, d, tie)[0]
            below = np.array([L.FLT_MIN, np.nextafter(reca["t"], np.float32(0))], dtype=np.float32)
            assert not port.triangle_hit(tv, [0, 1, 2], o, d, below)[0]
            assert not reference.triangle_hit(tv, [0, 1, 2], o, d, below)[0]


def test_sphere_square_cube(port, reference):
    rng = np.random.default_rng(3)
    spheres, squares, cubes = H.cornell_spheres(), H.cornell_squares(), H.cornell_cubes()
    for i in range(3000):
        o = rng.uniform((-245, 0, 0), (800, 555, 555)).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32); d = (d / np.linalg.norm(d)).astype(np.float32)
        r = np.array([L.FLT_MIN, L.FLT_MAX if i % 2 else rng.uniform(1, 900)], dtype=np.float32)
        s = spheres[i % len(spheres)]
        if i % 9 == 0:
            o = s["center"] + rng.uniform(-10, 10, 3).astype(np.float32)     # origin inside the sphere: second root
        a, b = port.sphere_hit(s, o, d, r), reference.sphere_hit(s, o, d, r)
        assert a[0] == b[0] and np.array_equal(bits(a[1]), bits(b[1])) and (not a[0] or _rec_equal(a[2], b[2]))
        q = squares[i % len(squares)]
        a, b = port.square_hit(q, o, d, r), reference.square_hit(q, o, d, r)
        assert a[0] == b[0] and np.array_equal(bits(a[1]), bits(b[1])) and (not a[0] or _rec_equal(a[2], b[2]))
        c = cubes[i % len(cubes)]
        if i % 4 == 0:                                                        # origin inside the cube: internal branch (AABB.hh:124)
            m = np.asarray(c["model"]).reshape(4, 4).T
            o = (m @ np.append(rng.uniform(0.1, 0.9, 3), 1.0).astype(np.float32))[:3].astype(np.float32)
        a, b = port.cube_hit(c, o, d, r), reference.cube_hit(c, o, d, r)
        assert a[0] == b[0] and np.array_equal(bits(a[1]), bits(b[1])) and (not a[0] or _rec_equal(a[2], b[2]))


def test_square_parallel_ray_is_rejected(port, reference):
    q = H.cornell_squares()[2]                                               # plane y = 555
    o = np.array([100, 555, 100], dtype=np.float32); d = np.array([1, 0, 0], dtype=np.float32)
    r = np.array([L.FLT_MIN, L.FLT_MAX], dtype=np.float32)
    assert not port.square_hit(q, o, d, r)[0] and not reference.square_hit(q, o, d, r)[0]    # 0/0 = NaN (Square.hh:84)
    o[1] = 500
    assert not port.square_hit(q, o, d, r)[0] and not reference.square_hit(q, o, d, r)[0]    # x/0 = inf


def test_offset_ray_and_ray_ctor(port, reference):
    rng = np.random.default_rng(4)
    from tracer_b200._lib import lib
    for i in range(2000):
        p = rng.uniform(-600, 600, 3).astype(np.float32)
        if i % 3 == 0:
            p *= np.float32(1e-3)                                             # |p| < 1/32 branch (Math.hh:71-73)
        if i % 17 == 0:
            p[0] = 0.0
        n = rng.normal(size=3).astype(np.float32); n = (n / np.linalg.norm(n)).astype(np.float32)
        want = reference.offset_ray(p, n)
        assert np.array_equal(bits(port.offset_ray(p, n)), bits(want))
        out = np.zeros(3, dtype=np.float32)
        lib.trqh_offset_ray(p.ctypes.data, n.ctypes.data, out.ctypes.data)    # the product's host restatement
        assert np.array_equal(bits(out), bits(want))
        ray = np.zeros(1, dtype=L.ray_dtype); ray["d"][0] = rng.uniform(-5, 5, 3)
        want_d = reference.ray_ctor(np.zeros(3, np.float32), ray["d"][0])[1]
        lib.trqh_normalize_rays(ray.ctypes.data, 1)
        assert np.array_equal(bits(ray["d"][0]), bits(want_d))


def test_next_float(reference):
    for v in (0.0, -0.0, 1.0, -1.0, 1e-30, 3.4e38, -3.4e38):
        up, dn = reference.lib.ref_next_float_up(v), reference.lib.ref_next_float_down(v)
        assert up == np.nextafter(np.float32(v), np.float32(np.inf)) and dn == np.nextafter(np.float32(v), np.float32(-np.inf))
