"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on identical rays.

Bar (BASELINE.json north_star): hit flags and primitive ids bit-exact; t / barycentrics within 1e-5
relative -- we assert BIT-EXACT t/u/v for triangle, square and cube hits (same arithmetic order, no
FMA), and 1e-5 absolute on sphere uv only (atan2f/asinf differ between libm and CUDA by ulps).
"""
import numpy as np
import pytest

from tracer_b200 import layout as L

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def assert_hits_equal(got, want, where=""):
    assert got.size == want.size
    for k in ("flags", "pType", "pIndex", "leafNode", "material"):
        bad = np.nonzero(got[k] != want[k])[0]
        assert bad.size == 0, f"{where}: {k} differs on {bad.size} rays, first {bad[:5]}: got {got[k][bad[:5]]} want {want[k][bad[:5]]}"
    bad = np.nonzero(got["t"].view(np.uint32) != want["t"].view(np.uint32))[0]
    assert bad.size == 0, f"{where}: t differs on {bad.size} rays"
    sph = want["pType"] == L.SPHERE
    hit = (want["flags"] & 1) == 1
    exact = hit & ~sph
    for k in ("u", "v"):
        assert np.array_equal(got[k][exact].view(np.uint32), want[k][exact].view(np.uint32)), f"{where}: {k} not bit-exact"
        if sph.any():
            assert np.allclose(got[k][hit & sph], want[k][hit & sph], rtol=0, atol=1e-5), f"{where}: sphere {k}"


def gpu_trace(scene, rays, any=False, reflayout=False, host=False):
    torch = _torch()
    from tracer_b200 import hits_to_numpy, rays_to_torch
    if host:
        return scene.hit(rays, any=any, reflayout=reflayout)
    d = rays_to_torch(rays, f"cuda:{scene.device}")
    h = scene.hit(d, any=any, reflayout=reflayout)
    torch.cuda.synchronize()
    return hits_to_numpy(h)


@pytest.fixture(scope="module")
def c2(built):
    _torch()
    from tracer_b200 import Scene, harness as H
    prim = H.scene_c2()
    return prim, Scene(prim, 0)


@pytest.fixture(scope="module")
def refcornell(built):
    _torch()
    from tracer_b200 import Scene, harness as H
    prim = H.scene_reference_cornell()
    return prim, Scene(prim, 0)


@pytest.mark.parametrize("reflayout", [True, False])
@pytest.mark.parametrize("any_hit", [False, True])
def test_c2_primary_and_bounce(c2, port, reflayout, any_hit):
    from tracer_b200 import harness as H
    prim, scene = c2
    rays = H.cornell_camera_rays(384, 216)
    want = port.trace(prim, rays, any=any_hit, records=True, nthreads=8)
    assert_hits_equal(gpu_trace(scene, rays, any_hit, reflayout), want["hits"], "primary")
    bounce, _ = H.bounce_rays(port.trace(prim, rays, records=True, nthreads=8)["records"])
    want = port.trace(prim, bounce, any=any_hit, nthreads=8)
    assert_hits_equal(gpu_trace(scene, bounce, any_hit, reflayout), want["hits"], "bounce")


@pytest.mark.parametrize("reflayout", [True, False])
def test_mixed_primitives(refcornell, port, reflayout):
    """Cube + Square + Sphere + Triangle leaves (the reference's actual Cornell leaf mix)."""
    from tracer_b200 import harness as H
    prim, scene = refcornell
    rays = H.cornell_camera_rays(320, 180)
    first = port.trace(prim, rays, records=True, nthreads=8)
    assert_hits_equal(gpu_trace(scene, rays, False, reflayout), first["hits"], "primary")
    bounce, _ = H.bounce_rays(first["records"])
    assert_hits_equal(gpu_trace(scene, bounce, False, reflayout), port.trace(prim, bounce, nthreads=8)["hits"], "bounce")
    shadow, _ = H.shadow_rays(first["records"], prim.squareList[5:6], prim.squareList[6:7])
    assert_hits_equal(gpu_trace(scene, shadow, True, reflayout), port.trace(prim, shadow, any=True, nthreads=8)["hits"], "shadow")
    rnd = H.random_rays(100000, seed=5, lo=(-245, 0, 0), hi=(800, 555, 555))
    assert_hits_equal(gpu_trace(scene, rnd, False, reflayout), port.trace(prim, rnd, nthreads=8)["hits"], "random")


def test_expand_hits_matches_oracle_records(refcornell, port):
    torch = _torch()
    from tracer_b200 import harness as H, rays_to_torch
    prim, scene = refcornell
    rays = H.cornell_camera_rays(320, 180)
    want = port.trace(prim, rays, records=True, nthreads=8)["records"]
    d = rays_to_torch(rays, "cuda:0")
    recs = scene.expand(d, scene.hit(d)).cpu().numpy().view(L.record_dtype).reshape(-1)
    assert np.array_equal(recs["hit"], want["hit"])
    m = want["hit"] == 1
    hits = port.trace(prim, rays, nthreads=8)["hits"]
    sph = hits["pType"] == L.SPHERE
    for k in ("t", "p", "gn", "sn", "front", "material"):
        assert np.array_equal(recs[k][m].view(np.uint32), want[k][m].view(np.uint32)), k
    assert np.array_equal(recs["uv"][m & ~sph].view(np.uint32), want["uv"][m & ~sph].view(np.uint32))
    assert np.allclose(recs["uv"][m & sph], want["uv"][m & sph], rtol=0, atol=1e-5)


def test_host_pointer_path_equals_device_path(c2, port):
    from tracer_b200 import harness as H
    prim, scene = c2
    rays = H.random_rays(300000, seed=11, lo=(-245, 0, 0), hi=(800, 555, 555))
    a = gpu_trace(scene, rays, host=False)
    b = gpu_trace(scene, rays, host=True)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert_hits_equal(a, port.trace(prim, rays, nthreads=8)["hits"], "random")


def test_soup_incoherent(built, port):
    """C5-shaped workload scaled down: 200k-triangle soup, 400k uniform random rays."""
    _torch()
    from tracer_b200 import Scene, harness as H
    prim = H.scene_soup(200000, seed=1, extent=0.01)
    scene = Scene(prim, 0)
    rays = H.random_rays(400000, seed=2)
    want = port.trace(prim, rays, nthreads=8)
    for reflayout in (True, False):
        assert_hits_equal(gpu_trace(scene, rays, False, reflayout), want["hits"], f"soup reflayout={reflayout}")
    short = rays.copy(); short["tmax"] = 0.05
    want = port.trace(prim, short, any=True, nthreads=8)
    assert_hits_equal(gpu_trace(scene, short, True, False), want["hits"], "soup any, finite tmax")


def test_edge_cases(built, port):
    """Empty batch, single-primitive scene (root is a leaf, SURVEY appendix A note 4), two-leaf scene,
    axis-parallel directions, origin inside boxes, ragged batch sizes around the warp/block size."""
    torch = _torch()
    from tracer_b200 import Scene, harness as H
    tri = H.make_vertices(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float32), [[0, 1, 2], [0, 1, 3]])
    one = H.build_primitive(tri, np.array([0, 1, 2], dtype=np.uint32))
    two = H.build_primitive(tri, np.array([0, 1, 2, 0, 1, 3], dtype=np.uint32))
    assert one.bvhList.size == 1 and two.bvhList.size == 3
    for prim in (one, two):
        scene = Scene(prim, 0)
        assert scene.hit(torch.empty((0, 8), dtype=torch.float32, device="cuda:0")).shape[0] == 0
        for n in (1, 31, 32, 33, 255, 257, 1000):
            rays = H.random_rays(n, seed=n, lo=(-0.5, -0.5, -0.5), hi=(1, 1, 1))
            axis = rays.copy()
            axis["d"] = 0; axis["d"][:, n % 3] = 1.0          # axis-parallel: 1/0 = inf slabs
            for r in (rays, axis):
                want = port.trace(prim, r)["hits"]
                for reflayout in (True, False):
                    assert_hits_equal(gpu_trace(scene, r, False, reflayout), want, f"n={n}")
        scene.close()
