"""CPU: the workload generators (csrc/host/harness.cpp) against the reference functions they restate."""
import numpy as np

from tracer_b200 import harness as H, layout as L
from tracer_b200._lib import lib

from .util import bits


def test_pcg32_known_answer():
    """pcg32_srandom(42, 54): the upstream pcg-c-basic demo sequence (Random.metal:3-19 == pcg_basic.c:44-67)."""
    out = np.zeros(6, dtype=np.uint32)
    lib.trqh_pcg32_fill_u32(42, 54, 6, out.ctypes.data)
    assert [hex(x) for x in out] == ["0xa15c02b7", "0x7b47f409", "0xba1d3330", "0x83d2f293", "0xbfa4784b", "0xcbed606e"]
    f = H.pcg32_floats(42, 54, 6)
    assert np.array_equal(f, np.ldexp(out.astype(np.float32), -32).astype(np.float32)) and (f >= 0).all() and (f <= 1).all()


def test_pcg32_is_the_references(reference):
    """The harness / producer PCG32 against the reference's own Random.metal (compiled verbatim): raw draws, randomF's
    ldexp(float(i), -32), and RandomSampler::sample2D's draw order."""
    rng = np.random.default_rng(4)
    for _ in range(25):
        seed, seq = int(rng.integers(0, 1 << 62)), int(rng.integers(0, 1 << 62))
        want_u, want_f = reference.pcg32(seed, seq, 64)
        got_u = np.zeros(64, dtype=np.uint32)
        lib.trqh_pcg32_fill_u32(seed, seq, 64, got_u.ctypes.data)
        assert np.array_equal(got_u, want_u)
        assert np.array_equal(bits(H.pcg32_floats(seed, seq, 64)), bits(want_f))
        assert np.array_equal(bits(H.pcg32_floats(seed, seq, 2)), bits(reference.sample2d(seed, seq)))


def test_camera_rays_are_reference_rays(reference):
    rays = H.cornell_camera_rays(32, 18)
    assert np.allclose(np.linalg.norm(rays["d"], axis=1), 1.0, atol=1e-6)
    assert (rays["o"] == np.array([278, 278, -800], dtype=np.float32)).all() and (rays["tmax"] == np.float32(L.FLT_MAX)).all()
    # the centre pixel looks down +z; normalisation is the Ray ctor's (three IEEE divides)
    d = rays["d"][9 * 32 + 16]
    assert abs(d[0]) < 1e-6 and abs(d[1]) < 1e-6 and d[2] > 0.999
    o2, d2 = reference.ray_ctor(rays["o"][5], rays["d"][5] * np.float32(3.0))
    assert np.allclose(d2, rays["d"][5], atol=1e-7)


def test_camera_rays_are_castray_bit_for_bit(reference, reference_setup):
    """trqh_gen_camera_rays == the reference's castRay (Camera.hh:59-69) over the camera its prepareCamera builds
    (Tracer.mm:371-411), with s = x / W, t = y / H (Render.metal:523-524): origins and directions bit for bit."""
    W, Hh = 64, 36
    rays = H.cornell_camera_rays(W, Hh)
    cam = reference_setup.prepare_camera(W, Hh)
    for y in range(0, Hh, 5):
        for x in range(0, W, 7):
            o, d = reference.cast_ray(cam, np.float32(x) / np.float32(W), np.float32(y) / np.float32(Hh), seed=x + 1, seq=y + 1)
            k = y * W + x
            assert np.array_equal(bits(rays["o"][k]), bits(o)) and np.array_equal(bits(rays["d"][k]), bits(d)), (x, y)


def test_bounce_rays_follow_the_reference_spawn(reference, port):
    prim = H.scene_reference_cornell()
    first = reference.trace(prim, H.cornell_camera_rays(48, 27))
    rays, src = H.bounce_rays(first, seed_base=7)
    assert rays.size == int(first["hit"].sum()) and np.array_equal(src, np.nonzero(first["hit"])[0].astype(np.uint32))
    for k in range(0, rays.size, 37):
        rec = first[src[k]]
        assert np.array_equal(bits(rays["o"][k]), bits(reference.offset_ray(rec["p"], rec["sn"])))          # Render.metal:450
        uu = H.pcg32_floats(7 + int(src[k]), 1, 2)
        nx, ny = reference.coordinate_system(rec["sn"])
        wi = reference.cosine_sample_hemisphere(uu)
        want = ((nx * wi[0]).astype(np.float32) + (ny * wi[1]).astype(np.float32)).astype(np.float32)
        want = (want + (rec["sn"] * wi[2]).astype(np.float32)).astype(np.float32)                             # stw * wi  :475
        _, want = reference.ray_ctor(rays["o"][k], want)                                                      # ray.update normalises
        assert np.array_equal(bits(rays["d"][k]), bits(want)), k                                              # host producer: bit for bit
        assert np.dot(rays["d"][k], rec["sn"]) >= -1e-6                                                        # into the hemisphere of sn


def test_shadow_rays_point_at_the_lights(reference):
    prim = H.scene_reference_cornell()
    first = reference.trace(prim, H.cornell_camera_rays(48, 27))
    la, lb = prim.squareList[5:6], prim.squareList[6:7]
    rays, src = H.shadow_rays(first, la, lb, seed_base=3)
    assert rays.size == int(first["hit"].sum())
    end = rays["o"] + rays["d"] * rays["tmax"][:, None]
    on_a = np.abs(end[:, 1] - 554.9) < 0.5
    on_b = np.abs(end[:, 0] + 300) < 0.5
    assert (on_a | on_b).all() and on_a.any() and on_b.any()
    assert np.allclose(np.linalg.norm(rays["d"], axis=1), 1.0, atol=1e-6)


def test_shadow_rays_follow_the_reference_spawn(reference):
    """Render.metal:313-337 composed from the reference's own pieces (sample2D, offset_ray, Square::sample, the Ray ctor):
    the producer's shadow rays are those rays bit for bit."""
    f32 = np.float32
    prim = H.scene_reference_cornell()
    first = reference.trace(prim, H.cornell_camera_rays(48, 27))
    la, lb = prim.squareList[5], prim.squareList[6]
    rays, src = H.shadow_rays(first, prim.squareList[5:6], prim.squareList[6:7], seed_base=3)
    picked = [0, 0]
    for k in range(0, rays.size, 23):
        rec = first[src[k]]
        u, f = reference.pcg32(3 + int(src[k]), 1, 3)
        origin = reference.offset_ray(rec["p"], rec["sn"])                       # :316
        which = 0 if f[2] < f32(0.5) else 1                                      # :319-323
        picked[which] += 1
        lp, _ = reference.square_sample(la if which == 0 else lb, f[:2], origin)
        dirv = (lp - origin).astype(np.float32)                                  # :325
        _, nor = reference.ray_ctor(origin, dirv)                                # normalize(_dir)  :326
        _, d = reference.ray_ctor(origin, nor)                                   # Ray(_origin, _nor) normalises again  :335
        dis = np.sqrt(f32(f32(dirv[0] * dirv[0] + dirv[1] * dirv[1]) + dirv[2] * dirv[2]), dtype=np.float32)   # length(_dir)  :334
        assert np.array_equal(bits(rays["o"][k]), bits(origin)), k
        assert np.array_equal(bits(rays["d"][k]), bits(d)), k
        assert bits(rays["tmax"][k:k + 1])[0] == bits(np.array([dis], dtype=np.float32))[0], k
    assert picked[0] > 0 and picked[1] > 0


def test_random_rays_and_soup_are_deterministic_and_shardable():
    a = H.random_rays(1000, seed=2)
    b = np.concatenate([H.random_rays(400, seed=2, first=0), H.random_rays(600, seed=2, first=400)])
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))                     # stream-per-ray => any sharding gives the same rays
    assert (a["d"] != 0).all() and np.allclose(np.linalg.norm(a["d"], axis=1), 1.0, atol=1e-6)
    s1, s2 = H.scene_soup(100, seed=1), H.scene_soup(100, seed=1)
    assert np.array_equal(s1.triList.view(np.uint8), s2.triList.view(np.uint8))


def test_subdivide_and_scene_sizes():
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], dtype=np.float32)
    p2, t2 = H.subdivide(pos, [[0, 1, 2], [1, 3, 2]], 2)
    assert len(t2) == 32 and len(p2) == 25                                       # shared edge midpoints
    c2 = H.scene_c2()
    assert c2.nTri == 14 + 24 + 15704


_M64 = (1 << 64) - 1
_MULT = 6364136223846793005


def _pcg_step(state, inc):
    return (state * _MULT + inc) & _M64


def _pcg_seeded(seed, seq):
    """pcg32_srandom_r (pcg_basic.c:44-51): -> (state, inc)."""
    inc = ((seq << 1) | 1) & _M64
    state = _pcg_step(0, inc)
    state = (state + seed) & _M64
    return _pcg_step(state, inc), inc


def test_rng_state_texture_is_the_references_wire_format(reference):
    """Row f-4: the spawns draw from / write back the per-pixel RNG texture exactly as Render.metal:511-557 does
    (toRNG / exRNG, Render.hh:96-120, executed from the verbatim build) -- including toRNG's quirk: it fills
    pcg32_t {state, inc} positionally with (inc, state), so the halves of a texel trade places at every kernel entry."""
    probe = np.array([0x11111111, 0x22222222, 0x33333333, 0x44444444], dtype=np.uint32)
    inc, state = reference.to_rng(probe)
    assert (inc, state) == (0x1111111122222222, 0x3333333344444444)       # (r, g) -> .inc, (b, a) -> .state
    assert list(reference.ex_rng(inc, state)) == [0x33333333, 0x44444444, 0x11111111, 0x22222222]
    prim = H.scene_reference_cornell()
    first = reference.trace(prim, H.cornell_camera_rays(48, 27))
    n = first.size
    hit = first["hit"] == 1
    rng = np.random.default_rng(5)
    tex0 = rng.integers(0, 1 << 32, size=(n, 4), dtype=np.uint64).astype(np.uint32)
    tex = tex0.copy()
    H.rng_frame_begin(tex)                                             # the reference's kernel entry
    rays, src = H.bounce_rays(first, rng_state=tex)
    for i in np.nonzero(~hit)[0][::17]:                                # a pixel that spawns nothing draws nothing:
        assert np.array_equal(tex[i], reference.ex_rng(*reference.to_rng(tex0[i])))   # entry + exit only
    for i in np.nonzero(hit)[0][::29]:
        inc, state = reference.to_rng(tex0[i])
        state = _pcg_step(_pcg_step(state, inc), inc)                  # sample2D = two draws
        assert np.array_equal(tex[i], reference.ex_rng(inc, state)), i
    # a texture that holds the streams PCG32(seedBase + i, 1) reproduces the seeded producer bit for bit
    seeded = np.zeros((n, 4), dtype=np.uint32)
    for i in range(n):
        state, inc = _pcg_seeded(11 + i, 1)
        seeded[i] = reference.ex_rng(inc, state)
    a, sa = H.bounce_rays(first, rng_state=seeded)
    b, sb = H.bounce_rays(first, seed_base=11)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8)) and np.array_equal(sa, sb)
    # three draws for a shadow ray (sample2D + the light pick), and pixel_of routes a compacted wave to its pixels
    la, lb = prim.squareList[5:6], prim.squareList[6:7]
    tex = tex0.copy()
    H.rng_frame_begin(tex)
    begun = tex.copy()
    sub = np.nonzero(hit)[0].astype(np.uint32)
    rays2, src2 = H.shadow_rays(first[sub], la, lb, pixel_of=sub, rng_state=tex)
    assert np.array_equal(src2, sub)                                   # srcIndex carries the pixel on
    i = int(sub[3])
    inc, state = reference.to_rng(tex0[i])
    for _ in range(3):
        state = _pcg_step(state, inc)
    assert np.array_equal(tex[i], reference.ex_rng(inc, state))
    assert np.array_equal(tex[~hit], begun[~hit])


def test_scene_constants_are_the_references(reference_setup):
    """The Cornell squares, the two BVH cubes, the 12 spheres and the camera the workloads restate, against the reference's
    own set-up code (Tracer.mm prepareCornellBox / prepareCubeList / prepareSphereList / prepareCamera, compiled from its
    source with the lists built in the application's order so that material indices agree)."""
    def same(a, b, fields, where):
        for f in fields:
            x, y = a[f], b[f]
            ok = np.array_equal(bits(x), bits(y)) if x.dtype.kind == "f" else np.array_equal(x, y)
            assert ok, f"{where}.{f}: {x} vs {y}"

    sq, ref_sq = H.cornell_squares(), reference_setup.squares()
    assert sq.size == ref_sq.size == 7
    # MakeSquare leaves normal / inverse matrices unset (Tracer.mm:127-153); nothing on the path reads them
    same(sq, ref_sq, ("axis_i", "axis_j", "axis_k", "range_i", "range_j", "value_k", "model", "material", "box_mini", "box_maxi"), "square")
    cu, ref_cu = H.cornell_cubes(), reference_setup.cubes()
    assert ref_cu.size == 3 and cu.size == 2                          # the third (density volume) cube never enters the BVH
    same(cu, ref_cu[:2], ("model", "box_mini", "box_maxi", "material"), "cube")
    for f in ("inverse", "normal"):                                   # Apple's simd_inverse is closed source: tolerance
        assert np.allclose(cu[f], ref_cu[:2][f], rtol=1e-6, atol=1e-9), f
    sp, ref_sp = H.cornell_spheres(), reference_setup.spheres()
    assert sp.size == ref_sp.size == 12
    same(sp, ref_sp, ("radius", "center", "model", "material", "box_mini", "box_maxi"), "sphere")
    # camera: prepareCamera(view 1920x1080, no rotation / offset) == MakeCamera with the constants the harness passes
    want = reference_setup.prepare_camera(1920, 1080)
    got = np.zeros(18, dtype=np.float32)
    frm, at, up = (np.array(v, dtype=np.float32) for v in ((278, 278, -800), (278, 278, 278), (0, 1, 0)))
    lib.trqh_make_camera(frm.ctypes.data, at.ctypes.data, up.ctypes.data, np.float32(45 * (np.pi / 180)),
                         np.float32(1920) / np.float32(1080), np.float32(10.0), got.ctypes.data)
    assert np.array_equal(bits(got), bits(want)), (got, want)
    rng = np.random.default_rng(8)
    for _ in range(20):                                               # MakeCamera for arbitrary cameras
        frm, at = rng.uniform(-5, 5, 3).astype(np.float32), rng.uniform(-5, 5, 3).astype(np.float32)
        vfov, aspect, focus = np.float32(rng.uniform(0.3, 1.5)), np.float32(rng.uniform(0.5, 2.5)), np.float32(rng.uniform(1, 20))
        want = reference_setup.make_camera(frm, at, up, 0.0, aspect, vfov, focus)
        lib.trqh_make_camera(frm.ctypes.data, at.ctypes.data, up.ctypes.data, vfov, aspect, focus, got.ctypes.data)
        assert np.allclose(got, want, rtol=2e-6, atol=1e-6)           # simd_normalize multiplies by 1/sqrt, the restatement divides


def test_scene_constants_frozen():
    """The same constants without the reference build (runs on the GPU box too): material numbering in application order."""
    assert list(H.cornell_squares()["material"]) == [5, 4, 6, 6, 6, 3, 3]
    assert list(H.cornell_cubes()["material"]) == [0, 19]
    assert list(H.cornell_spheres()["material"]) == list(range(7, 19))
