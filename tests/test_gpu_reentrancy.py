"""GPU: trq_trace is re-entrant -- many launches in flight on many streams, and more launches than the queue-head ring
holds. Every launch draws its own work-queue head and the kernel's last CTA re-arms it (no memset, no follow-up pass),
so a head is zero whenever a launch picks it up -- whatever kind of launch used it before (round-1 regression: heads left
dirty by one kernel variant made the next variant skip rays)."""
import threading

import numpy as np
import pytest

from tracer_b200 import layout as L

from .test_gpu_parity import _torch

pytestmark = pytest.mark.gpu


def test_512_launches_in_flight_on_8_streams(built):
    torch = _torch()
    from tracer_b200 import Scene, harness as H, rays_to_torch
    prim = H.scene_soup(100000, seed=1, extent=0.02)
    scene = Scene(prim, 0)
    sizes = [50000, 1, 333, 20000, 4097, 65536, 31, 9999]
    batches = [rays_to_torch(H.random_rays(n, seed=100 + k), "cuda:0") for k, n in enumerate(sizes)]
    want = [scene.hit(b).clone() for b in batches]
    want_any = [scene.hit(b, any=True).clone() for b in batches]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(8)]
    outs = []
    for rep in range(64):                                      # 8 x 64 = 512 launches queued without a single synchronise
        for k, st in enumerate(streams):
            j = (k + rep) % len(batches)
            any_hit = (rep + k) % 3 == 0
            o = torch.empty_like(batches[j])
            scene.hit(batches[j], any=any_hit, out=o, stream=st.cuda_stream)
            outs.append((j, any_hit, o))
    torch.cuda.synchronize()
    for j, any_hit, o in outs:
        assert torch.equal(o.view(torch.int32), (want_any if any_hit else want)[j].view(torch.int32))
    scene.close()


def test_more_launches_than_queue_heads_and_mixed_kernels(built):
    """> 4096 launches (the ring wraps), alternating closest / any / 16-byte records / the reference-layout kernel."""
    torch = _torch()
    from tracer_b200 import Scene, harness as H, rays_to_torch
    prim = H.scene_soup(20000, seed=2, extent=0.05)
    scene = Scene(prim, 0)
    d = rays_to_torch(H.random_rays(3000, seed=5), "cuda:0")
    want = {(a, h): scene.hit(d, any=a, hit16=h).clone() for a in (False, True) for h in (False, True)}
    ref = scene.hit(d, reflayout=True).clone()
    assert torch.equal(ref.view(torch.int32), want[(False, False)].view(torch.int32))
    out32, out16 = torch.empty_like(want[(False, False)]), torch.empty_like(want[(False, True)])
    for k in range(4500):
        a, h = bool(k & 1), bool(k & 2)
        o = scene.hit(d, any=a, hit16=h, out=out16 if h else out32)
        if k % 500 == 499 or k > 4090:
            assert torch.equal(o.view(torch.int32), want[(a, h)].view(torch.int32)), k
        if k % 1000 == 0:
            scene.hit(d, reflayout=True, out=out32)
    scene.close()


def test_host_threads_share_one_scene(built):
    torch = _torch()
    from tracer_b200 import Scene, harness as H, rays_to_torch
    prim = H.scene_soup(50000, seed=3, extent=0.03)
    scene = Scene(prim, 0)
    scene.profile(True)                                        # the profiling ring is shared state too
    batches = [rays_to_torch(H.random_rays(20000 + 1000 * k, seed=k), "cuda:0") for k in range(4)]
    want = [scene.hit(b).clone() for b in batches]
    torch.cuda.synchronize()
    errors = []

    def worker(k):
        try:
            torch.cuda.set_device(0)
            st = torch.cuda.Stream()
            o = torch.empty_like(batches[k])
            for _ in range(100):
                scene.hit(batches[k], out=o, stream=st.cuda_stream)
            st.synchronize()
            if not torch.equal(o.view(torch.int32), want[k].view(torch.int32)):
                errors.append(k)
        except Exception as e:                                 # pragma: no cover
            errors.append(repr(e))

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors
    n, a, b = scene.profile_read()
    assert 0 < n <= 64 and a > 0
    scene.profile(False)
    scene.close()
